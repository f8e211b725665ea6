#!/bin/bash
# rounds sweep (RD3_ROUNDS): parity suite at 9 rounds, then short benches on both scenes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
RD3_ROUNDS=9 timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r9.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_r9.log
B="python bench.py --steps 40 --no-rows --no-e2e --no-cpu-baseline --no-masks"
for r in 9 10; do
  RD3_ROUNDS=$r timeout 100 $B > gpurun_out/r_mix_$r.json 2>/dev/null
  RD3_ROUNDS=$r timeout 100 $B --scene ground > gpurun_out/r_gnd_$r.json 2>/dev/null
done
python - <<'PY'
import json
for f in ("r_mix_9","r_gnd_9","r_mix_10","r_gnd_10"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],4), round(d["path_roofline"]["frac"],4), {k:round(v,3) for k,v in d["path_roofline"]["stage_ms_per_step_single_stream"].items()})
    except Exception as e: print(f,"ERR",e)
PY
