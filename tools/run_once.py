"""One flux-series pass on a chosen shape, for ncu captures and quick timings.

    python tools/run_once.py --workload C4 --nt 8 --dtype f64 [--pad 4] [--passes 3] [--opt 4=2 --opt 11=1]

Prints the CUDA-event time of every pass after the first (warm-up) one.  Under ncu:
    ncu --set full --clock-control none --import-source on -k regex:k23_fused -s 1 -c 1 -o gpurun_out/prof python tools/run_once.py ...
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemoflux_b200 import _lib, nemoflux_gpu, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='C3')
ap.add_argument('--nt', type=int, default=16)
ap.add_argument('--dtype', default='f64', choices=['f64', 'f32'])
ap.add_argument('--pad', type=int, default=-1, help='-1 = 32-byte rows (4 doubles / 8 floats), 0 = dense')
ap.add_argument('--passes', type=int, default=3)
ap.add_argument('--classic', action='store_true')
ap.add_argument('--opt', action='append', default=[], help='nfx option as id=value, repeatable')
ap.add_argument('--transects', type=int, default=0)
ap.add_argument('--out', default='')
a = ap.parse_args()
for kv in a.opt:
    k, val = kv.split('=')
    _lib.set_option(int(k), int(val))
dev = torch.device('cuda', 0)
syn = synth.make(a.workload, **({'ntransects': a.transects} if a.transects else {}))
g = nemoflux_gpu.Grid()
g.setPoints(syn.points)
g.setCGridShape(syn.ny, syn.nx)
p = nemoflux_gpu.PolylineIntegral()
p.build(g)
p.computeWeights(syn.transects)
tdt = torch.float32 if a.dtype == 'f32' else torch.float64
pad = a.pad if a.pad >= 0 else (4 if a.dtype == 'f64' else 8)
u, v = syn.fill_device(0, a.nt, dev, dtype=tdt, pad=pad)
th, a1, a2 = (torch.from_numpy(x).to(dev) for x in (syn.thickness, syn.arc1, syn.arc2))
eflux = torch.empty((a.nt, 2 * syn.ncell), dtype=torch.float64, device=dev) if a.classic else None
out = torch.empty((a.nt, syn.ntransects), dtype=torch.float64, device=dev)
nbytes = (8.0 if a.dtype == 'f32' else 16.0) * syn.units_per_step() * a.nt
ms = []
for i in range(a.passes + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    p.fluxSeries(u, v, th, a1, a2, out=out, eflux=eflux)
    e1.record()
    torch.cuda.synchronize()
    if i:
        ms.append(e0.elapsed_time(e1))
st = p.seriesStatus() if hasattr(p, 'seriesStatus') else None
res = dict(workload=a.workload, nt=a.nt, dtype=a.dtype, pad=pad, fused=_lib.get_option(_lib.NFX_OPT_LAST_SERIES_PATH),
           panels=p.getNumberOfPanels(a.dtype)[0], ms=ms, gbs=[nbytes / m / 1e6 for m in ms], algorithmic_bytes=nbytes,
           status=st, finite=bool(torch.isfinite(out).all()))
print(json.dumps(res))
if a.out:
    os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
    json.dump(res, open(a.out, 'w'), indent=1)
