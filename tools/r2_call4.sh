#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -12 gpurun_out/pytest.log
B="python bench.py --steps 50 --no-rows --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/b_mix.json 2> gpurun_out/b_mix.err
timeout 300 $B --scene ground > gpurun_out/b_gnd.json 2> gpurun_out/b_gnd.err
python - <<'PY'
import json
for f in ("b_mix","b_gnd"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "masks", (d.get("with_masks") or {}).get("ms_per_step"), round(d["ms_per_step"],4), round(d["path_roofline"]["frac"],4), {k:round(v,3) for k,v in d["path_roofline"]["stage_ms_per_step_single_stream"].items()})
    except Exception as e: print(f, "ERR", e)
PY
