#!/bin/bash
# Cost breakdown of the lookup-only insert rounds by knock-out builds (profiles/r1_analysis.md, third pass).
# The RD3_KNOCK builds produce WRONG results and are for timing only; they never replace the in-tree library.
#   here:      bash tools/knockout.sh build        (writes ab_libs/k{0,1,2,3}.so, restores the default build)
#   on a GPU:  bash tools/knockout.sh run          (prints the per-stage times of each build, mixture scene)
set -e
cd "$(dirname "$0")/.."
B=3d-reconstruction-detection_b200
if [ "$1" = build ]; then
  mkdir -p ab_libs
  for k in 1 2 3; do
    RD3_NVCC_EXTRA="-DRD3_KNOCK=$k" python $B/build.py --force > /dev/null
    cp $B/librd3_b200.so ab_libs/k$k.so
  done
  python $B/build.py --force > /dev/null
  cp $B/librd3_b200.so ab_libs/k0.so
else
  python tools/ab.py mixture k0=ab_libs/k0.so k1=ab_libs/k1.so k2=ab_libs/k2.so k3=ab_libs/k3.so
fi
