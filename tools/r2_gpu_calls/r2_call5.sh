#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c5_bench_n2.json 2> gpurun_out/c5_bench_n2.err; echo "bench n2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --impl reference > gpurun_out/c5_ref_n2.json 2> gpurun_out/c5_ref_n2.err; echo "ref n2 rc=$?"
tail -c 600 gpurun_out/c5_bench_n2.err
