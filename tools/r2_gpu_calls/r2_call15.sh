#!/bin/bash
# fused pass: upper bound of the whole ring (evict-last lines in L2) on multi-panel grids; slot size per dtype
mkdir -p gpurun_out
for cfg in "C5 8 f32 8" "C5 8 f32 4" "C4 40 f32 8" "C5 8 f64 4" "C4 40 f64 4" "C3 365 f32 0" "C3 365 f64 0"; do
  for mx in 64 32 24 16 12; do
    set -- $cfg $mx
    timeout 300 python tools/run_once.py --workload $1 --nt $2 --dtype $3 --passes 4 --opt 5=$4 --opt 16=$5 > gpurun_out/c15_tmp.json 2> gpurun_out/c15_err.log || echo "FAILED $cfg $mx"
    python - <<PY
import json
d=json.load(open('gpurun_out/c15_tmp.json'))
print('$1 $3 nt=$2 slot=$4 MB ring<=$5 MB panels', d['panels'], 'ms', ['%.3f'%x for x in d['ms']], 'GB/s', max(round(x) for x in d['gbs']), 'finite', d['finite'])
PY
  done
done
NFX_DEBUG_GUARDS=1 python -m pytest tests -m gpu -x -q > gpurun_out/c15_pytest_guards.log 2>&1; echo "pytest with guard bands rc=$?"; tail -3 gpurun_out/c15_pytest_guards.log
