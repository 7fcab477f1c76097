#!/bin/bash
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c21_bench_n8.json 2> gpurun_out/c21_bench_n8.err ) 2> gpurun_out/c21_time_n8.txt; echo "bench n8 rc=$?"
tail -n 4 gpurun_out/c21_time_n8.txt
tail -c 600 gpurun_out/c21_bench_n8.err
