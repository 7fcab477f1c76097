"""run the flux-series pass a few times (for ncu captures): --dtype f32|f64 --workload C3 --nt 32 [--fused|--classic]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemoflux_b200 import _lib, nemoflux_gpu, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--dtype', default='f32')
ap.add_argument('--workload', default='C3')
ap.add_argument('--nt', type=int, default=32)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--pad', type=int, default=0)
ap.add_argument('--fused', action='store_true')
ap.add_argument('--classic', action='store_true')
a = ap.parse_args()
dev = torch.device('cuda', 0)
syn = synth.make(a.workload)
g = nemoflux_gpu.Grid()
g.setPoints(syn.points)
g.setCGridShape(syn.ny, syn.nx)
p = nemoflux_gpu.PolylineIntegral()
p.build(g)
p.computeWeights(syn.transects)
u, v = syn.fill_device(0, a.nt, dev, dtype=torch.float32 if a.dtype == 'f32' else torch.float64, pad=a.pad)
th, a1, a2 = (torch.from_numpy(x).to(dev) for x in (syn.thickness, syn.arc1, syn.arc2))
if a.fused:
    _lib.set_option(_lib.NFX_OPT_FAST_SERIES, 2)
if a.classic:
    _lib.set_option(_lib.NFX_OPT_FAST_SERIES, 0)
out = torch.empty((a.nt, syn.ntransects), dtype=torch.float64, device=dev)
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    p.fluxSeries(u, v, th, a1, a2, out=out)
    e1.record()
    torch.cuda.synchronize()
    print('ms', e0.elapsed_time(e1))
print('ok', float(out[0, 0]))
