#!/bin/bash
cd "$(dirname "$0")/.."
python tools/ab.py cur=ab_libs/cur.so cur_s1=ab_libs/cur.so,RD3_STREAMS:1 cur_s3=ab_libs/cur.so,RD3_STREAMS:3
