#!/bin/bash
# in-situ kernel times of the in-tree library and of variant builds:  VARIANTS="tools/bin/a.so ..." bash tools/_ab2.sh <tag>
cd "$(dirname "$0")/.."
for lib in "" $VARIANTS; do
  echo "== library: ${lib:-in-tree}"
  DPN_LIB_OVERRIDE=${lib:+$PWD/$lib} timeout 200 python tools/insitu_kernels.py 2>&1 | grep -v -i Warn | head -4
done 2>&1 | tee gpurun_out/${1:-ab2}.txt
