"""mint.VectorInterp on the GPU: findPoints + getFaceVectors for many points, timed against the C
oracle on a sample of the same points.     python tools/vinterp_bench.py --workload C3 --npts 1000000"""
import argparse
import json
import os
import sys
import time

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemoflux_b200 import nemoflux_gpu, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='C3')
ap.add_argument('--npts', type=int, default=1000000)
ap.add_argument('--cpu-pts', type=int, default=20000)
ap.add_argument('--out', default='gpurun_out/vinterp_bench.json')
a = ap.parse_args()
syn = synth.make(a.workload)
rng = numpy.random.default_rng(11)
lo = syn.points[..., :2].reshape(-1, 2).min(0)
hi = syn.points[..., :2].reshape(-1, 2).max(0)
pts = numpy.zeros((a.npts, 3))
pts[:, 0] = rng.uniform(lo[0], hi[0], a.npts)
pts[:, 1] = rng.uniform(lo[1] + 1., hi[1] - 1., a.npts)
data = rng.standard_normal((syn.ncell, 4))
g = nemoflux_gpu.Grid()
g.setPoints(syn.points)
vi = nemoflux_gpu.VectorInterp()
vi.setGrid(g)
vi.buildLocator()
res = dict(workload=a.workload, ncell=int(syn.ncell), npts=a.npts)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nmiss = vi.findPoints(pts)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    vec = vi.getFaceVectors(data)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
res.update(find_points_s=t1 - t0, face_vectors_s=t2 - t1, not_found=int(nmiss),
           gpu_points_per_s=a.npts / (t2 - t0))
og = O.Grid(syn.points)
ov = O.VectorInterp(og)
sub = pts[:a.cpu_pts]
t0 = time.perf_counter()
ov.findPoints(sub)
ref = ov.getFaceVectors(data)
t_cpu = time.perf_counter() - t0
cells, _ = vi.getCells()
found = numpy.asarray(cells[:a.cpu_pts]) >= 0
scale = numpy.abs(ref[found]).max()
res.update(cpu_points=a.cpu_pts, cpu_s=t_cpu, cpu_points_per_s=a.cpu_pts / t_cpu,
           max_rel_diff=float(numpy.abs(numpy.asarray(vec)[:a.cpu_pts][found] - ref[found]).max() / scale),
           speedup=(a.npts / (t2 - t0 if False else res['find_points_s'] + res['face_vectors_s'])) / (a.cpu_pts / t_cpu))
os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
json.dump(res, open(a.out, 'w'), indent=1)
print(json.dumps(res, indent=1))
