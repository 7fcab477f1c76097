#!/bin/bash
# fused pass (CTA-granular items): lag of the K3 items behind the K2 tiles of their batch (1 = round-1 behaviour, 0 = automatic)
mkdir -p gpurun_out
for cfg in "C3 365 f32" "C3 365 f64" "C4 40 f32" "C4 40 f64" "C5 8 f32" "C5 8 f64"; do
  set -- $cfg
  for lag in 1 0 1 0; do
    timeout 300 python tools/run_once.py --workload $1 --nt $2 --dtype $3 --passes 5 --opt 12=0 --opt 4=2 --opt 15=$lag > gpurun_out/c11_$1_$3_lag$lag.json 2> gpurun_out/c11_err.log || echo "FAILED $cfg lag $lag"
    python - <<PY
import json
d=json.load(open('gpurun_out/c11_$1_$3_lag$lag.json'))
print('$1 $3 nt=$2 lag=$lag', 'ms', ['%.3f'%x for x in d['ms']], 'GB/s', max(round(x) for x in d['gbs']), 'finite', d['finite'], 'status', d['status'])
PY
  done
done
python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1; tail -3 gpurun_out/c11_pytest.log
