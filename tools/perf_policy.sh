#!/bin/bash
# L2 policy / promotion sweep of the tensor copies for misaligned frame rows (N % 128 != 0)
for pol in 0 1 2; do for promo in 0 64 128 256; do
  echo -n "policy=$pol promo=$promo  "
  BGD_TMA_POLICY=$pol BGD_TMA_L2PROMO=$promo python tools/perf_configs.py ${1:-sthv2} 2>&1 | python -c "import sys,json; [print(round(json.loads(l)['GB/s']), json.loads(l)['parity_spotcheck']) for l in sys.stdin if l.startswith('{')]"
done; done
