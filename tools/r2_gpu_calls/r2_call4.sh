#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/c4_bench_n1.json 2> gpurun_out/c4_bench_n1.err; echo "bench n1 rc=$?"
python bench.py --steps 3 --warmup 3 --max-chunk-steps 100 --no-cpu --no-e2e > gpurun_out/c4_bench_n1_chunked.json 2> gpurun_out/c4_bench_n1_chunked.err; echo "bench chunked rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; tail -3 gpurun_out/c4_pytest.log
