#!/bin/bash
# parity + env sweeps of the fused path (short benches)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
B="python bench.py --steps 50 --no-rows --no-e2e --no-cpu-baseline --no-masks"
run() { name=$1; shift; env "$@" timeout 300 $B $EXTRA > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err; }
run base X=1
EXTRA="--scene ground" run gnd X=1
python - <<'PY'
import json,glob
for f in ["base","gnd"]:
    try:
        d=json.load(open("gpurun_out/sw_%s.json"%f))
        print("%-10s"%f, "masks", (d.get("with_masks") or {}).get("ms_per_step"), round(d["ms_per_step"],4), round(d["path_roofline"]["frac"],4), {k:round(v,3) for k,v in d["path_roofline"]["stage_ms_per_step_single_stream"].items()})
    except Exception as e: print(f, "ERR", e)
PY
