#!/bin/bash
# fused pass: impl 0 (k23_fused) vs impl 2 (team kernel, unified queue: prefetched item id, per-warp slot check, one barrier), lag 1 / auto
mkdir -p gpurun_out
for cfg in "C3 365 f32" "C3 365 f64" "C4 40 f32" "C4 40 f64" "C5 8 f32" "C5 8 f64"; do
  set -- $cfg
  for o in "0 1" "0 0" "2 1" "2 0"; do
    set -- $cfg $o
    timeout 300 python tools/run_once.py --workload $1 --nt $2 --dtype $3 --passes 5 --opt 12=$4 --opt 4=2 --opt 15=$5 > gpurun_out/c12_$1_$3_impl$4_lag$5.json 2> gpurun_out/c12_err.log || echo "FAILED $cfg $o"
    python - <<PY
import json
d=json.load(open('gpurun_out/c12_$1_$3_impl$4_lag$5.json'))
print('$1 $3 nt=$2 impl=$4 lag=$5', 'ms', ['%.3f'%x for x in d['ms']], 'GB/s', max(round(x) for x in d['gbs']), 'finite', d['finite'], 'status', d['status'])
PY
  done
done
