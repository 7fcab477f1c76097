#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c16_topo.txt 2>&1
free -g > gpurun_out/c16_mem.txt; nproc >> gpurun_out/c16_mem.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c16_bench_n8.json 2> gpurun_out/c16_bench_n8.err ) 2> gpurun_out/c16_time_n8.txt; echo "bench n8 rc=$?"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/c16_bench_n4.json 2> gpurun_out/c16_bench_n4.err ) 2> gpurun_out/c16_time_n4.txt; echo "bench n4 rc=$?"
tail -3 gpurun_out/c16_time_n8.txt gpurun_out/c16_time_n4.txt
tail -c 400 gpurun_out/c16_bench_n8.err
