#!/bin/bash
# round-2 profile call: stage times, then ncu --set full on the pass kernels (one stream lane = 64 frames per launch)
cd "$(dirname "$0")/.."
python tools/ab.py cur=ab_libs/cur.so stcs=ab_libs/stcs.so > gpurun_out/r2_c3_ab.log 2>&1
export RD3_STREAMS=1
CMD="python bench.py --profile-only --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r2_c3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hv_pass_kernel -s 9 -c 9 -o gpurun_out/r2_pass $CMD > gpurun_out/r2_c3_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hv_rank|hv_emit|hv_cull|hv_flagscan" -s 4 -c 4 -o gpurun_out/r2_misc $CMD > gpurun_out/r2_c3_ncu2.log 2>&1
