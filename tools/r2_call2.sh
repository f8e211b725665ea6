#!/bin/bash
# parity suite + default bench (+ ground scene, short)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
timeout 600 python bench.py --no-rows > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --scene ground --steps 30 --no-rows --no-e2e --no-cpu-baseline > gpurun_out/bench_ground.json 2> gpurun_out/bench_ground.err; echo "ground rc=$?"
python - <<'PY'
import json
for f in ("bench_default","bench_ground"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, d["ms_per_step"], d["path_roofline"]["frac"], d["path_roofline"]["stage_ms_per_step_single_stream"])
        print("  masks", d.get("with_masks",{}).get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("frames_per_sec"))
    except Exception as e: print(f, "ERR", e)
PY
