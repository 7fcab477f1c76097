#!/bin/bash
# fused pass, float32 storage: kernel shapes (vector width x levels in flight x CTAs per SM)
mkdir -p gpurun_out
for cfg in "C3 128 f32" "C4 24 f32"; do
  set -- $cfg
  for o in "7=45" "10=3" "7=43" "7=85" "7=83" "7=45"; do
    timeout 300 python tools/run_once.py --workload $1 --nt $2 --dtype $3 --passes 5 --opt 12=0 --opt 4=2 --opt $o > gpurun_out/c8_tmp.json 2> gpurun_out/c8_err.log || echo "FAILED $cfg $o"
    python - <<PY
import json
d=json.load(open('gpurun_out/c8_tmp.json'))
print('$1 $3 nt=$2 impl=0 opt $o', 'ms', ['%.3f'%x for x in d['ms']], 'GB/s', max(round(x) for x in d['gbs']), 'finite', d['finite'], 'status', d['status'])
PY
  done
done
