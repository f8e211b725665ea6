#!/bin/bash
# round-2 A/B of the strip-pass pipeline: stream lanes, rounds / table size
cd "$(dirname "$0")/.."
python tools/ab.py m5=ab_libs/m5.so m5_s1=ab_libs/m5.so,RD3_STREAMS:1 m5_s3=ab_libs/m5.so,RD3_STREAMS:3 \
  m5_r16_l70=ab_libs/m5.so,RD3_ROUNDS:16,RD3_TABLE_LOAD_PCT:70
