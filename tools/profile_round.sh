#!/bin/bash
# One GPU call that regenerates the evidence under profiles/ (run through gpurun; outputs land in gpurun_out/).
# ROWS=1 also refreshes the per-row table (tools/bench_rows.py, both scenes).
set -x
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --frames 64 --steps 2 --warmup 1 --profile-only > gpurun_out/ncu_l.log 2>&1
RD3_STREAMS=1 ncu --set full --clock-control none --import-source on -k regex:hv_insert -s 12 -c 1 -f \
    -o gpurun_out/prof_insert_full python bench.py --frames 64 --steps 2 --warmup 1 --profile-only > gpurun_out/ncu_f.log 2>&1
RD3_STREAMS=1 ncu --set full --clock-control none --import-source on -k regex:hv_emit -s 1 -c 1 -f \
    -o gpurun_out/prof_emit_full python bench.py --frames 64 --steps 2 --warmup 1 --profile-only > gpurun_out/ncu_e.log 2>&1
if [ -n "$ROWS" ]; then
  python tools/bench_rows.py --scene mixture > gpurun_out/rows_mixture.txt 2> gpurun_out/rows.err
  python tools/bench_rows.py --scene ground > gpurun_out/rows_ground.txt 2>> gpurun_out/rows.err
fi
tail -c 400 gpurun_out/bench_default.json
