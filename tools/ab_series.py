"""A/B the flux-series passes on ONE box: classic (K2 -> HBM -> K3) vs fused persistent pass with options.

    python tools/ab_series.py --workload C3 [--nt-local 365] [--rounds 5]
"""
import argparse
import json
import os
import sys

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemoflux_b200 import _lib, nemoflux_gpu, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='C3')
    ap.add_argument('--nt-local', type=int, default=0)
    ap.add_argument('--rounds', type=int, default=5)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--slots', default='8')
    ap.add_argument('--out', default='')
    ap.add_argument('--nx', type=int, default=0)
    ap.add_argument('--dtype', default='f64', choices=['f64', 'f32'])
    ap.add_argument('--pad', type=int, default=0)
    ap.add_argument('--f64-ctas', default='', help='fused float64 kernel: register budgets (CTAs per SM) to compare, e.g. 2,3')
    ap.add_argument('--base-order', type=int, default=0, help='NFX_OPT_FUSED_ORDER of the modes that do not set it')
    ap.add_argument('--orders', default='', help='fused batch visiting orders / K3 unroll to compare (NFX_OPT_FUSED_ORDER), e.g. 0,1,2,3')
    ap.add_argument('--f32-shapes', default='', help='fused float32 kernel shapes to compare, e.g. 85,45,83,43')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    syn = synth.make(a.workload, **({'nx': a.nx} if a.nx else {}))
    nt = a.nt_local or syn.nt
    g = nemoflux_gpu.Grid()
    g.setPoints(syn.points)
    g.setCGridShape(syn.ny, syn.nx)
    p = nemoflux_gpu.PolylineIntegral()
    p.build(g)
    p.computeWeights(syn.transects)
    tdt = torch.float32 if a.dtype == 'f32' else torch.float64
    u, v = syn.fill_device(0, nt, dev, dtype=tdt, pad=a.pad)
    th, a1, a2 = (torch.from_numpy(x).to(dev) for x in (syn.thickness, syn.arc1, syn.arc2))
    eflux = torch.empty((nt, 2 * syn.ncell), dtype=torch.float64, device=dev)
    out = torch.empty((nt, syn.ntransects), dtype=torch.float64, device=dev)
    nbytes = (8.0 if a.dtype == 'f32' else 16.0) * syn.units_per_step() * nt
    modes = [('classic', dict(eflux=eflux), {})]
    for mb in [int(x) for x in a.slots.split(',')]:
        modes.append((f'fused slot={mb}MB', {}, {_lib.NFX_OPT_RING_SLOT_MB: mb, _lib.NFX_OPT_FAST_SERIES: 2}))
    for shape in [int(x) for x in a.f32_shapes.split(',') if x]:
        modes.append((f'fused f32 shape={shape}', {}, {_lib.NFX_OPT_RING_SLOT_MB: 0, _lib.NFX_OPT_FAST_SERIES: 2,
                                                         _lib.NFX_OPT_FUSED_F32_SHAPE: shape}))
    for o in [int(x) for x in a.orders.split(',') if x]:
        modes.append((f'fused order={o}', {}, {_lib.NFX_OPT_RING_SLOT_MB: 0, _lib.NFX_OPT_FAST_SERIES: 2,
                                               _lib.NFX_OPT_FUSED_ORDER: o}))
    for c in [int(x) for x in a.f64_ctas.split(',') if x]:
        modes.append((f'fused f64 ctas/SM={c}', {}, {_lib.NFX_OPT_FUSED_F64_CTAS: c}))
    times = {m[0]: [] for m in modes}
    ref = None
    for rnd in range(a.rounds):
        for name, kw, opts in modes:
            full = {_lib.NFX_OPT_RING_SLOT_MB: 0, _lib.NFX_OPT_FAST_SERIES: 2, _lib.NFX_OPT_FUSED_F32_SHAPE: 0,
                    _lib.NFX_OPT_FUSED_ORDER: a.base_order, _lib.NFX_OPT_FUSED_F64_CTAS: 0}
            full.update(opts)                       # every mode sets every knob: nothing leaks from the mode before
            for k, val in full.items():
                _lib.set_option(k, val)
            p.fluxSeries(u, v, th, a1, a2, out=out, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                p.fluxSeries(u, v, th, a1, a2, out=out, **kw)
            e1.record()
            torch.cuda.synchronize()
            times[name].append(e0.elapsed_time(e1) / a.reps)
            res = out.cpu().numpy().copy()
            if ref is None:
                ref = res
            assert numpy.abs(res - ref).max() <= 1e-12 * numpy.abs(ref).max(), name
    rows = []
    for name, ts in times.items():
        med = float(numpy.median(ts))
        rows.append(dict(mode=name, ms_median=med, ms_min=float(min(ts)), ms_max=float(max(ts)), gbs=nbytes / med / 1e6))
        print(f'{name:34s} median {med:8.3f} ms  min {min(ts):8.3f}  max {max(ts):8.3f}   {nbytes / med / 1e6:8.1f} GB/s')
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(dict(workload=a.workload, nt=nt, rows=rows), open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
