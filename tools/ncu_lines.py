"""Per-source-line cost of one kernel: joins the SASS page of an .ncu-rep (instructions executed, stall
samples per instruction) with the line table of the object it was built from (nvdisasm --print-line-info).
The object must be the build that was profiled (same SASS).

    python tools/ncu_lines.py rep.ncu-rep 3d-reconstruction-detection_b200/csrc/depth.o hv_pass_kernelINS_11DepthSourceELi1 [min_instr] [launch index]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kern = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
min_instr = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
skip = sys.argv[5] if len(sys.argv) > 5 else "0"            # which launch of a multi-launch report

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
again = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(again) > 1:                       # newer ncu prints the page once per view: keep the first
    rows = rows[:again[1]]
hdr, data = rows[1], rows[2:]
iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
warps = int(data[0][iE])

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cubin)], capture_output=True, text=True).stdout

lines, on, cur = [], False, ("?", 0)
for ln in dis.splitlines():
    if ln.startswith(".text."):
        on = kern in ln
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
if len(lines) != len(data):
    sys.exit("instruction count mismatch: object %d vs report %d (not the profiled build?)" % (len(lines), len(data)))

agg = {}
for (f, l), r in zip(lines, data):
    a = agg.setdefault((f, l), [0.0, 0, {}])
    a[0] += int(r[iE]) / warps
    a[1] += int(r[iS])
    for i, name in stalls:
        v = int(r[i] or 0)
        if v:
            a[2][name] = a[2].get(name, 0) + v
src = {}
tot_i = sum(a[0] for a in agg.values())
tot_s = sum(a[1] for a in agg.values())
print("warps %d, %.1f instructions per warp, %d samples" % (warps, tot_i, tot_s))
print("%-22s %8s %6s %8s %6s  %s" % ("file:line", "instr/w", "%", "samples", "%", "top stalls | source"))
for (f, l), a in sorted(agg.items()):
    if a[0] < min_instr and a[1] < 0.005 * tot_s:
        continue
    if f not in src:
        p = os.path.join(os.path.dirname(obj), f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:70] if 0 < l <= len(src[f]) else ""
    top = " ".join("%s=%d" % kv for kv in sorted(a[2].items(), key=lambda kv: -kv[1])[:2])
    print("%-22s %8.1f %5.1f%% %8d %5.1f%%  %-28s | %s" % ("%s:%d" % (f, l), a[0], 100 * a[0] / tot_i, a[1],
                                                          100.0 * a[1] / max(tot_s, 1), top, text))
