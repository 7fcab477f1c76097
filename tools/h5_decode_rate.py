"""Host-side decode rate of chunked + shuffled + deflated NetCDF-4 variables (what XIOS writes for uo/vo), read the way
Field.fluxSeries reads them: one task per (time step, slab of levels) on a thread pool (zlib releases the GIL).
No GPU needed.   python tools/h5_decode_rate.py [--workload C3] [--nt 4] [--threads 8]"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import h5build  # noqa: E402
from nemoflux_b200 import ncio, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='C3')
ap.add_argument('--nt', type=int, default=4)
ap.add_argument('--threads', type=int, default=max(1, min(8, (os.cpu_count() or 2) // 2)))
ap.add_argument('--out', default='')
a = ap.parse_args()
syn = synth.make(a.workload)
u = numpy.stack([syn.uv_host(t)[0].astype(numpy.float32) for t in range(a.nt)])
u[numpy.isnan(u)] = 1.e20
path = '/tmp/nfx_h5_decode.nc'
h5build.write(path, {'uo': dict(data=u, chunks=(1, 1, syn.ny, syn.nx), deflate=True, shuffle=True,
                                attrs={'_FillValue': numpy.float32(1.e20)})})
fsize = os.path.getsize(path)
rows = []
with ncio.open_dataset(path) as nc:
    var = nc['uo']
    dst = numpy.zeros(u.shape, numpy.float32)
    for threads in sorted({1, a.threads}):
        pool = ThreadPoolExecutor(max_workers=threads)
        nsplit = max(1, min(syn.nz, 2 * threads))
        zs = [(k * syn.nz // nsplit, (k + 1) * syn.nz // nsplit) for k in range(nsplit)]
        t0 = time.perf_counter()
        jobs = [pool.submit(var.read_into, dst[t, z0:z1], (t, slice(z0, z1))) for t in range(a.nt) for z0, z1 in zs]
        for j in jobs:
            j.result()
        dt = time.perf_counter() - t0
        pool.shutdown()
        assert numpy.array_equal(dst, u)
        rows.append(dict(threads=threads, seconds=round(dt, 4), decoded_GBps=round(u.nbytes / dt / 1e9, 3),
                         file_GBps=round(fsize / dt / 1e9, 3)))
        print(json.dumps(rows[-1]), flush=True)
res = dict(workload=a.workload, nt=a.nt, decoded_bytes=int(u.nbytes), file_bytes=int(fsize),
           compression_ratio=round(u.nbytes / fsize, 2), cores=os.cpu_count(), rows=rows)
print(json.dumps(res))
if a.out:
    json.dump(res, open(a.out, 'w'), indent=1)
os.unlink(path)
