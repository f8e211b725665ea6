#!/bin/bash
# round-2 profile call (one GPU): launch list of one bench step, then ncu --set full on the pass / emit kernels
# (one stream lane = 64 frames per launch, like the bench's own per-stage timing)
cd "$(dirname "$0")/.."
export RD3_STREAMS=1
CMD="python bench.py --profile-only --steps 2 --warmup 1"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hv_pass_kernel|hv_emit_kernel" -s 10 -c 10 -o gpurun_out/r2_pass $CMD > gpurun_out/r2_ncu1.log 2>&1
