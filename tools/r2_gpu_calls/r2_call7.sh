#!/bin/bash
# fused pass (CTA-granular items, impl 0): software prefetch into L2, distance in batches of 5 levels
mkdir -p gpurun_out
for cfg in "C3 128 f32" "C3 128 f64" "C4 24 f32" "C4 24 f64" "C5 4 f32" "C5 4 f64"; do
  set -- $cfg
  for pf in 0 1 2 4 0 2; do
    timeout 300 python tools/run_once.py --workload $1 --nt $2 --dtype $3 --passes 5 --opt 12=0 --opt 4=2 --opt 14=$pf > gpurun_out/c7_$1_$3_pf$pf.json 2> gpurun_out/c7_err.log || echo "FAILED $cfg pf $pf"
    python - <<PY
import json
d=json.load(open('gpurun_out/c7_$1_$3_pf$pf.json'))
print('$1 $3 nt=$2 impl=0 prefetch=$pf', 'ms', ['%.3f'%x for x in d['ms']], 'GB/s', max(round(x) for x in d['gbs']), 'finite', d['finite'], 'status', d['status'])
PY
  done
done
