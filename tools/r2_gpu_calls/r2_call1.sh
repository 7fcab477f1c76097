#!/bin/bash
# round-2 GPU call 1: f32 conversion A/B, read ceiling, compute-sanitizer on small shapes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_gpu.txt
python tools/f32_conv_ab.py --workload C3 --nt 128 --out gpurun_out/c1_f32_conv_C3.json > gpurun_out/c1_f32_conv_C3.log 2>&1
python tools/f32_conv_ab.py --workload C4 --nt 24 --out gpurun_out/c1_f32_conv_C4.json > gpurun_out/c1_f32_conv_C4.log 2>&1
tools/readbw > gpurun_out/c1_readbw.txt 2>&1
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 $SAN --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_san_smoke_$tool.log 2>&1
  echo "smoke $tool rc=$?" >> gpurun_out/c1_san_rc.txt
done
timeout 900 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_fast_series_path_matches_classic or test_batch_range_partials_add_up or test_k1_random_transects_bit_exact" > gpurun_out/c1_san_tests_memcheck.log 2>&1
echo "tests memcheck rc=$?" >> gpurun_out/c1_san_rc.txt
timeout 600 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_fast_series_path_matches_classic or test_batch_range_partials_add_up" > gpurun_out/c1_san_tests_racecheck.log 2>&1
echo "tests racecheck rc=$?" >> gpurun_out/c1_san_rc.txt
cat gpurun_out/c1_san_rc.txt
