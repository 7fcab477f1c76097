#!/bin/bash
# fused pass (k23_fused): ring slot size on multi-panel grids
mkdir -p gpurun_out
for cfg in "C5 8 f32" "C4 40 f32" "C5 8 f64" "C4 40 f64"; do
  for mb in 8 4 2 16; do
    set -- $cfg $mb
    timeout 300 python tools/run_once.py --workload $1 --nt $2 --dtype $3 --passes 4 --opt 12=0 --opt 4=2 --opt 15=1 --opt 5=$4 > gpurun_out/c13_tmp.json 2> gpurun_out/c13_err.log || echo "FAILED $cfg $mb"
    python - <<PY
import json
d=json.load(open('gpurun_out/c13_tmp.json'))
print('$1 $3 nt=$2 slot=$4 MB panels', d['panels'], 'ms', ['%.3f'%x for x in d['ms']], 'GB/s', max(round(x) for x in d['gbs']), 'finite', d['finite'])
PY
  done
done
