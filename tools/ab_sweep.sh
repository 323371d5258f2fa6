#!/bin/bash
# A/B of library builds in ONE run on one box: usage tools/ab_sweep.sh "<lib names>" "<T cases>" [rounds]
# (alternates the libraries round by round so clock / power drift hits all of them alike)
libs=$1; cases=$2; rounds=${3:-2}
for r in $(seq 1 $rounds); do
  for lib in $libs; do
    echo "== round $r $lib"
    BGD_LIB_PATH=$PWD/background-debiased-video-cil_b200/$lib COMBOS=0:0:0 ITERS=10 python tools/perf_sweep_ldsm.py $cases 2>&1 | grep -E "^T=|ERROR"
  done
done
