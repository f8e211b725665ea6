#!/bin/bash
# parity + env sweeps of the fused path (short benches)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 50 --no-rows --no-e2e --no-cpu-baseline --no-masks"
run() { name=$1; shift; env "$@" timeout 300 $B $EXTRA > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err; }
run base X=1
run st1 RD3_STREAMS=1
run st3 RD3_STREAMS=3
run st4 RD3_STREAMS=4
EXTRA="--scene ground" run gnd X=1
EXTRA="--scene ground" run gnd_st1 RD3_STREAMS=1
EXTRA="--scene ground" run gnd_st3 RD3_STREAMS=3
python - <<'PY'
import json,glob
for f in ["base","st1","st3","st4","gnd","gnd_st1","gnd_st3"]:
    try:
        d=json.load(open("gpurun_out/sw_%s.json"%f))
        print("%-10s"%f, "masks", (d.get("with_masks") or {}).get("ms_per_step"), round(d["ms_per_step"],4), round(d["path_roofline"]["frac"],4), {k:round(v,3) for k,v in d["path_roofline"]["stage_ms_per_step_single_stream"].items()})
    except Exception as e: print(f, "ERR", e)
PY
