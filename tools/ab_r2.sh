#!/bin/bash
# round-2 A/B: item queue slack + L2 prefetch of the table sectors, emit occupancy
cd "$(dirname "$0")/.."
python tools/ab.py cur=ab_libs/cur.so nopf=ab_libs/nopf.so s32=ab_libs/s32.so e8=ab_libs/e8.so cur_r16_l70=ab_libs/cur.so,RD3_ROUNDS:16,RD3_TABLE_LOAD_PCT:70
