"""Concurrent pinned-host -> device copy bandwidth of every rank, before and after binding the rank to the CPUs
NVML reports as local to its GPU (nemoflux_b200.dist.bind_to_gpu_cpus).  Explains the e2e figure of bench.py at
8 GPUs: python -m torch.distributed.run --nproc-per-node 8 tools/numa_probe.py"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemoflux_b200 import dist as nfx_dist  # noqa: E402

rank = int(os.environ.get('RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
lr = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)


def measure(tag, nbytes=2 << 30, reps=6):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)                      # touch every page from this thread
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    gbs = nbytes * reps / (time.perf_counter() - t) / 1e9
    out = [None] * world
    if world > 1:
        dist.all_gather_object(out, gbs)
    else:
        out = [gbs]
    if rank == 0:
        print(json.dumps({'phase': tag, 'h2d_gbs_per_rank': [round(x, 1) for x in out], 'sum': round(sum(out), 1)}), flush=True)
    del h, d


info = {'rank': rank, 'cpus_allowed': len(os.sched_getaffinity(0))}
measure('unbound')
info['bound_to'] = nfx_dist.bind_to_gpu_cpus(lr)
info['cpus_after'] = len(os.sched_getaffinity(0))
allinfo = [None] * world
if world > 1:
    dist.all_gather_object(allinfo, info)
else:
    allinfo = [info]
if rank == 0:
    print(json.dumps(allinfo), flush=True)
    try:
        nodes = sorted(n for n in os.listdir('/sys/devices/system/node') if n.startswith('node'))
        print(json.dumps({n: open(f'/sys/devices/system/node/{n}/cpulist').read().strip() for n in nodes}), flush=True)
    except Exception as e:
        print('no NUMA info:', e, flush=True)
measure('bound to the GPU-local CPUs')
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
