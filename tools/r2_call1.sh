#!/bin/bash
# round-2 evidence call: GPU parity suite, default bench line, launch list, full captures of the two pass kernels + emit
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -3 gpurun_out/r2_pytest.log
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
CMD="python bench.py --profile-only --steps 2 --warmup 1"
export RD3_STREAMS=1
$CMD > gpurun_out/r2_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu0.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hv_pass_kernel|hv_emit_kernel" -s 10 -c 10 -f -o gpurun_out/r2_pass $CMD > gpurun_out/r2_ncu1.log 2>&1
head -c 3000 gpurun_out/r2_bench_default.json
