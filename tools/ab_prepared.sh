#!/bin/bash
# The prepared (default-off, not yet measured) kernel experiments of profiles/r1_analysis.md as one A/B.
#   here:      bash tools/ab_prepared.sh build    -> ab_libs/{cur,late,etma,prefetch,rankp,all4}.so, default build restored
#   on a GPU:  bash tools/ab_prepared.sh run      -> parity tests on the all-four build, then per-stage times of each
set -e
cd "$(dirname "$0")/.."
B=3d-reconstruction-detection_b200
if [ "$1" = build ]; then
  mkdir -p ab_libs
  mk() { RD3_NVCC_EXTRA="$2" python $B/build.py --force > /dev/null; cp $B/librd3_b200.so ab_libs/$1.so; }
  mk late "-DRD3_LATE_CLAIMS=1"
  mk etma "-DRD3_EMIT_TMA=1"
  mk prefetch "-DRD3_PREFETCH=1"
  mk rankp "-DRD3_RANK_PACKED=1"
  mk all4 "-DRD3_LATE_CLAIMS=1 -DRD3_EMIT_TMA=1 -DRD3_PREFETCH=1 -DRD3_RANK_PACKED=1"
  mk cur ""
else
  RD3_LIB_PATH=ab_libs/all4.so python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  python tools/ab.py cur=ab_libs/cur.so late=ab_libs/late.so etma=ab_libs/etma.so prefetch=ab_libs/prefetch.so rankp=ab_libs/rankp.so all4=ab_libs/all4.so
fi
