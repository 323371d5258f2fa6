#!/bin/bash
# A/B the experimental builds of the library: usage tools/ab_libs.sh "<lib names>" "<T cases>"
for lib in $1; do
  echo "== $lib"
  BGD_LIB_PATH=$PWD/background-debiased-video-cil_b200/$lib COLTHR=128 python tools/perf_sweep.py $2 2>&1 | grep -E "variant|ERROR|Error"
done
