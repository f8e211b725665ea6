#!/bin/bash
cd "$(dirname "$0")/.."
python tools/ab.py cur=ab_libs/cur.so
