#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --workload C5 --nt-total 9 --balance --no-e2e > gpurun_out/c22_bench_c5_n2.json 2> gpurun_out/c22_bench_c5_n2.err; echo "bench C5 balanced n2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c22_bench_n2.json 2> gpurun_out/c22_bench_n2.err; echo "bench n2 rc=$?"
tail -c 300 gpurun_out/c22_bench_c5_n2.err
