#!/bin/bash
# per-launch ncu metrics of the insert kernel for each library given (names under ab_libs/)
for v in "$@"; do
  RD3_LIB_PATH=$PWD/ab_libs/$v.so RD3_STREAMS=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum --clock-control none -k regex:hv_insert -s 8 -c 8 --csv --log-file gpurun_out/rounds_$v.csv python bench.py --frames 64 --steps 1 --warmup 1 --profile-only ${SCENE:+--scene $SCENE} > /dev/null 2>&1
done
