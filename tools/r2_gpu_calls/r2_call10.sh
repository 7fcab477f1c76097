#!/bin/bash
mkdir -p gpurun_out
for cfg in "f32 64" "f32 1" "f64 64" "f64 1"; do
  set -- $cfg
  timeout 300 python tools/run_once.py --workload C3 --nt 365 --dtype $1 --passes 5 --opt 12=0 --transects $2 > gpurun_out/c10_tmp.json 2> gpurun_out/c10_err.log || echo "FAILED $cfg"
  python - <<PY
import json
d=json.load(open('gpurun_out/c10_tmp.json'))
print('C3 nt=365 $1 transects=$2', 'ms', ['%.3f'%x for x in d['ms']], 'GB/s', max(round(x) for x in d['gbs']))
PY
done
