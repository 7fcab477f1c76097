#!/bin/bash
# A/B of the fused pass: CTA-granular items (impl 0) vs warp-specialised (impl 1); then the GPU test-suite with the new default
mkdir -p gpurun_out
for cfg in "C3 128 f64" "C3 128 f32" "C4 24 f64" "C4 24 f32" "C5 4 f64" "C5 4 f32"; do
  set -- $cfg
  for impl in 0 1; do
    timeout 300 python tools/run_once.py --workload $1 --nt $2 --dtype $3 --passes 5 --opt 12=$impl --opt 4=2 > gpurun_out/c6_$1_$3_impl$impl.json 2> gpurun_out/c6_err.log || echo "FAILED $cfg impl $impl"
    python - <<PY
import json
d=json.load(open('gpurun_out/c6_$1_$3_impl$impl.json'))
print('$1 $3 nt=$2 impl=$impl', 'ms', ['%.3f'%x for x in d['ms']], 'GB/s', max(round(x) for x in d['gbs']), 'finite', d['finite'], 'status', d['status'])
PY
  done
done
python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1; tail -5 gpurun_out/c6_pytest.log
