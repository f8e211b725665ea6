#!/bin/bash
# round-2 evidence: GPU parity suite, the default bench line, a ground-scene line, launch list and full ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -3 gpurun_out/r2_pytest.log
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --scene ground --steps 50 --no-rows --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_ground.json 2> gpurun_out/r2_bench_ground.err; echo "ground rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
bash tools/r2_prof.sh > gpurun_out/r2_prof.log 2>&1
timeout 300 python tools/bench_rows.py --iters 20 > gpurun_out/r2_rows_mixture.txt 2> gpurun_out/r2_rows.err
python tools/prof_unproject.py > gpurun_out/up_plain.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:up_single -s 2 -c 1 -f -o gpurun_out/r2_unproject python tools/prof_unproject.py > gpurun_out/up_ncu.log 2>&1
python - <<'PY'
import json
for f in ("r2_bench_default","r2_bench_ground"):
    d=json.load(open("gpurun_out/%s.json"%f))
    print(f, round(d["ms_per_step"],4), round(d["path_roofline"]["frac"],4), {k:round(v,3) for k,v in d["path_roofline"]["stage_ms_per_step_single_stream"].items()})
    print("  masks", (d.get("with_masks") or {}).get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("frames_per_sec"), "cpu", (d.get("cpu_baseline") or {}).get("frames_per_sec"))
PY
