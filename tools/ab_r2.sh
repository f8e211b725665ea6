#!/bin/bash
cd "$(dirname "$0")/.."
python tools/ab.py cur=ab_libs/cur.so e8=ab_libs/e8.so e4=ab_libs/e4.so
