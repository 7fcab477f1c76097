"""probe: can a read-only memory-mapped file be page-locked in place (cudaHostRegister ... ReadOnly) and copied to the
device at PCIe rate without a host-side memcpy?  Prints GB/s for: registered mmap, plain mmap (pageable), pinned copy."""
import mmap
import os
import sys
import time

import numpy
import torch

path = sys.argv[1] if len(sys.argv) > 1 else '/tmp/nfx_probe.bin'
size = 1 << 30
if not os.path.exists(path) or os.path.getsize(path) != size:
    with open(path, 'wb') as f:
        f.write(numpy.random.default_rng(0).integers(0, 255, size, dtype=numpy.uint8).tobytes())
fd = os.open(path, os.O_RDONLY)
mm = mmap.mmap(fd, size, access=mmap.ACCESS_READ)
arr = numpy.frombuffer(mm, dtype=numpy.uint8)
_ = int(arr[::4096].sum())                                  # fault the pages in (page cache hot)
ptr = arr.ctypes.data
dev = torch.empty(size, dtype=torch.uint8, device='cuda')
rt = torch.cuda.cudart()


import warnings
warnings.filterwarnings('ignore')
src = torch.from_numpy(arr)


def h2d(label, n=3):
    best = 0.0
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, size / (time.perf_counter() - t0) / 1e9)
    print(f'{label}: {best:.1f} GB/s (torch sees pinned: {src.is_pinned()})', flush=True)


h2d('plain mmap (pageable) -> device')
print('cudart functions:', [n for n in dir(rt) if 'Host' in n or 'Memcpy' in n])
for flags, name in ((0x08, 'ReadOnly'), (0x00, 'Default'), (0x01 | 0x08, 'Portable|ReadOnly')):
    t0 = time.perf_counter()
    err = rt.cudaHostRegister(ptr, size, flags)
    print(f'cudaHostRegister flags={name}: err={err} in {time.perf_counter() - t0:.3f} s', flush=True)
    if int(err) == 0:
        try:
            h2d('registered mmap -> device')
        finally:
            print('unregister:', rt.cudaHostUnregister(ptr))
        break
pin = torch.empty(size, dtype=torch.uint8, pin_memory=True)
t0 = time.perf_counter()
pin.numpy()[:] = arr
print(f'memcpy mmap -> pinned, 1 thread: {size / (time.perf_counter() - t0) / 1e9:.1f} GB/s')
from concurrent.futures import ThreadPoolExecutor
for nth in (4, 8, 16):
    pool = ThreadPoolExecutor(nth)
    parts = numpy.array_split(numpy.arange(size // (1 << 20)), nth)
    pn = pin.numpy()

    def job(p):
        a, b = int(p[0]) << 20, (int(p[-1]) + 1) << 20
        pn[a:b] = arr[a:b]
    t0 = time.perf_counter()
    list(pool.map(job, parts))
    print(f'memcpy mmap -> pinned, {nth} threads: {size / (time.perf_counter() - t0) / 1e9:.1f} GB/s', flush=True)
    pool.shutdown()
t0 = time.perf_counter()
dev.copy_(pin, non_blocking=True)
torch.cuda.synchronize()
print(f'pinned -> device: {size / (time.perf_counter() - t0) / 1e9:.1f} GB/s')
print('cpu count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
