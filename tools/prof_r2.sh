#!/bin/bash
# round-2 profile call: ncu --set full on the pass kernels and the rank / cull kernels (one stream lane = 64 frames per launch)
cd "$(dirname "$0")/.."
export RD3_STREAMS=1
CMD="python bench.py --profile-only --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hv_pass_kernel -s 9 -c 9 -o gpurun_out/r2_pass $CMD > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hv_rank|hv_cull" -s 2 -c 2 -o gpurun_out/r2_misc $CMD > gpurun_out/r2_ncu2.log 2>&1
