#!/bin/bash
# round-2 A/B of the strip-pass pipeline: stream lanes, strip length, rounds / table size
cd "$(dirname "$0")/.."
python tools/ab.py m5=ab_libs/m5.so m5_s2=ab_libs/m5.so,RD3_STREAMS:2 m5_s1=ab_libs/m5.so,RD3_STREAMS:1 m5_s4=ab_libs/m5.so,RD3_STREAMS:4 \
  m5_l4=ab_libs/m5.so,RD3_LKP_ITERS:4 m5_l4_s2=ab_libs/m5.so,RD3_LKP_ITERS:4,RD3_STREAMS:2 \
  m5_r16_l70=ab_libs/m5.so,RD3_ROUNDS:16,RD3_TABLE_LOAD_PCT:70 m5_r16_l70_s2=ab_libs/m5.so,RD3_ROUNDS:16,RD3_TABLE_LOAD_PCT:70,RD3_STREAMS:2
