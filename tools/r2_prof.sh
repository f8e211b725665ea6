#!/bin/bash
# launch list of two bench steps + full captures of the pipeline kernels of one step (one stream lane)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export RD3_STREAMS=1
CMD="python bench.py --profile-only --steps 2 --warmup 1 ${SCENE:+--scene $SCENE}"
$CMD > gpurun_out/r2_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches${SCENE:+_$SCENE}.csv $CMD > gpurun_out/r2_ncu0.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hv_|calib" -s 14 -c 14 -f -o gpurun_out/r2_all${SCENE:+_$SCENE} $CMD > gpurun_out/r2_ncu1.log 2>&1
ls -la gpurun_out/*.ncu-rep
