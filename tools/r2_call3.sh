#!/bin/bash
# parity suite + short benches (filter on/off, both scenes) + launch list and full capture of every pipeline kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
B="python bench.py --steps 50 --no-rows --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/b_mix.json 2> gpurun_out/b_mix.err
RD3_FILTER=0 timeout 300 $B --no-masks > gpurun_out/b_mix_nofilter.json 2> gpurun_out/b_mix_nofilter.err
timeout 300 $B --scene ground > gpurun_out/b_gnd.json 2> gpurun_out/b_gnd.err
RD3_FILTER=0 timeout 300 $B --scene ground --no-masks > gpurun_out/b_gnd_nofilter.json 2> gpurun_out/b_gnd_nofilter.err
python - <<'PY'
import json
for f in ("b_mix","b_mix_nofilter","b_gnd","b_gnd_nofilter"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, round(d["ms_per_step"],4), round(d["path_roofline"]["frac"],4), {k:round(v,3) for k,v in d["path_roofline"]["stage_ms_per_step_single_stream"].items()})
        print("  masks", d.get("with_masks",{}).get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
if [ -n "$PROF" ]; then
CMD="python bench.py --profile-only --steps 2 --warmup 1"
export RD3_STREAMS=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu0.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hv_|calib" -s 14 -c 14 -f -o gpurun_out/r2_all $CMD > gpurun_out/r2_ncu1.log 2>&1
fi
