"""Randomized soak of the fused path against the oracle (more seeds than the test suite runs).

    python tools/soak.py [first_seed] [count]        # needs a CUDA device
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_gpu_parity as t  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
t0 = time.time()
bad = []
for s in range(first, first + count):
    try:
        t.test_fused_randomized(s)
    except AssertionError as e:                # keep going: report every failing seed
        bad.append((s, str(e)[:200]))
print("seeds %d..%d: %d failures in %.0f s" % (first, first + count - 1, len(bad), time.time() - t0))
for b in bad:
    print(b)
sys.exit(1 if bad else 0)
