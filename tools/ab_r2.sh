#!/bin/bash
# round-2 A/B of the strip-pass pipeline: register budget (CTAs/SM), strip length, rounds / table size, culling
cd "$(dirname "$0")/.."
python tools/ab.py m5=ab_libs/m5.so m4=ab_libs/m4.so \
  m5_nocull=ab_libs/m5.so,RD3_CULL:0 \
  m5_l4=ab_libs/m5.so,RD3_LKP_ITERS:4 m5_i8=ab_libs/m5.so,RD3_INS_ITERS:8 m5_i2=ab_libs/m5.so,RD3_INS_ITERS:2 \
  m5_r16_l70=ab_libs/m5.so,RD3_ROUNDS:16,RD3_TABLE_LOAD_PCT:70 \
  m5_s2=ab_libs/m5.so,RD3_STREAMS:2
