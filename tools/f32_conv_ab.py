"""Round-2 starter: the second float32 conversion flavour of the fused pass (NFX_OPT_FUSED_F32_CONV = 1, see
nfx_stream_ops.cuh::clean_scaled_v2 -- 37 % fewer ALU-pipe instructions per value in the SASS of the level loop).
1. bit-compares the flux series of both flavours on data with NaN land, +-0, denormals, FLT_MAX, infinities and a
   missing-value marker;  2. same-box A/B of the pass time.   python tools/f32_conv_ab.py [--workload C3] [--nt 64]"""
import argparse
import json
import os
import sys

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemoflux_b200 import _lib, nemoflux_gpu, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='C3')
ap.add_argument('--nt', type=int, default=64)
ap.add_argument('--rounds', type=int, default=7)
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--out', default='gpurun_out/f32_conv_ab.json')
a = ap.parse_args()
dev = torch.device('cuda', 0)
syn = synth.make(a.workload)
g = nemoflux_gpu.Grid()
g.setPoints(syn.points)
g.setCGridShape(syn.ny, syn.nx)
p = nemoflux_gpu.PolylineIntegral()
p.build(g)
p.computeWeights(syn.transects)
th, a1, a2 = (torch.from_numpy(x).to(dev) for x in (syn.thickness, syn.arc1, syn.arc2))
u, v = syn.fill_device(0, a.nt, dev, dtype=torch.float32, pad=8)


def series(conv, uu, vv, fill=float('nan')):
    _lib.set_option(_lib.NFX_OPT_FUSED_F32_CONV, conv)
    try:
        return p.fluxSeries(uu, vv, th, a1, a2, fill=fill)
    finally:
        _lib.set_option(_lib.NFX_OPT_FUSED_F32_CONV, 0)


# 1. bits: special values sprinkled over a copy of the first 4 steps
n = min(4, a.nt)
us, vs = u[:n].clone(), v[:n].clone()
rng = torch.Generator(device=dev)
rng.manual_seed(3)
r = torch.rand(us.shape, device=dev, generator=rng)
for lo, hi, val in ((0.00, 0.02, 0.0), (0.02, 0.04, -0.0), (0.04, 0.06, 1.0e-41), (0.06, 0.08, -1.4e-45),
                    (0.08, 0.10, 3.4028234663852886e38), (0.10, 0.12, 1.0e20), (0.12, 0.1202, float('inf')),
                    (0.1202, 0.1204, float('-inf')), (0.1204, 0.14, float('nan'))):
    m = (r >= lo) & (r < hi)
    us[m] = val
    vs[m.flip(-1)] = val
same = {}
for fill in (float('nan'), 1.0e20):
    s0, s1 = series(0, us, vs, fill).cpu().numpy(), series(1, us, vs, fill).cpu().numpy()
    assert _lib.get_option(_lib.NFX_OPT_LAST_SERIES_PATH) == 1
    same[str(fill)] = bool(numpy.array_equal(s0, s1, equal_nan=True))
print('bit-identical series:', same, flush=True)

# 2. time, interleaved
ms = {0: [], 1: []}
for _ in range(a.rounds):
    for conv in (0, 1):
        _lib.set_option(_lib.NFX_OPT_FUSED_F32_CONV, conv)
        for _ in range(2):
            p.fluxSeries(u, v, th, a1, a2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            p.fluxSeries(u, v, th, a1, a2)
        e1.record()
        torch.cuda.synchronize()
        ms[conv].append(e0.elapsed_time(e1) / a.reps)
_lib.set_option(_lib.NFX_OPT_FUSED_F32_CONV, 0)
nbytes = 2.0 * 4 * syn.units_per_step() * a.nt
res = dict(workload=a.workload, nt=a.nt, bit_identical=same,
           median_ms={k: float(numpy.median(x)) for k, x in ms.items()},
           GBps={k: nbytes / float(numpy.median(x)) / 1e6 for k, x in ms.items()})
print(json.dumps(res, indent=1))
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, 'w'), indent=1)
