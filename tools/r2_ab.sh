#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python tools/ab.py cur=ab_libs/cur.so ins6=ab_libs/ins6.so ins4=ab_libs/ins4.so lkp6=ab_libs/lkp6.so lkp4=ab_libs/lkp4.so emit5=ab_libs/emit5.so emit8=ab_libs/emit8.so cur3=ab_libs/cur.so,RD3_STREAMS:3 2>&1 | tee gpurun_out/ab.txt
