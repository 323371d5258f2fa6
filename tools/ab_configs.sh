#!/bin/bash
# same-box A/B of library builds on the resident-shard configs: usage tools/ab_configs.sh "<lib names>" "<configs>" [rounds]
libs=$1; cfgs=$2; rounds=${3:-2}
for r in $(seq 1 $rounds); do
  for lib in $libs; do
    echo "== round $r $lib"
    BGD_LIB_PATH=$PWD/background-debiased-video-cil_b200/$lib python tools/perf_configs.py $cfgs 2>&1 | python -c "import sys,json; [print(' ', json.loads(l)['config'], round(json.loads(l)['GB/s'])) for l in sys.stdin if l.startswith('{')]"
  done
done
