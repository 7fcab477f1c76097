"""float32 storage: time the stand-alone K2 launch per ALU-conversion mask (NFX_OPT_K2_ALU_MASK) and check that
every mask gives the same bits as the all-F2F kernel, on inputs with NaN, +-0, denormals and infinities."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemoflux_b200 import _lib, nemoflux_gpu, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--shape', default='C3')
ap.add_argument('--nt', type=int, default=64)
ap.add_argument('--reps', type=int, default=10)
ap.add_argument('--masks', default='0,1,9,21,31,-1')
a = ap.parse_args()
cfg = synth.CONFIGS[a.shape]
ncell, nz = cfg['nx'] * cfg['ny'], cfg['nz']
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(7)


def field(nasty):
    x = torch.randn((a.nt, nz, ncell), dtype=torch.float32, device=dev, generator=g)
    r = torch.rand((a.nt, nz, ncell), device=dev, generator=g)
    x[r < 0.30] = float('nan')
    x[(r >= 0.30) & (r < 0.31)] = 0.0
    x[(r >= 0.31) & (r < 0.32)] = -0.0
    if not nasty:                                    # the timed data: land as NaN, zeros, ordinary velocities
        return x
    x[(r >= 0.32) & (r < 0.325)] = 1.0e-41          # denormal
    x[(r >= 0.325) & (r < 0.3251)] = float('inf')
    x[(r >= 0.3251) & (r < 0.3252)] = float('-inf')
    x[(r >= 0.3252) & (r < 0.33)] = 3.0e38
    x[(r >= 0.33) & (r < 0.335)] = 1.2e-38          # smallest normals
    return x


u, v = field(False), field(False)
un, vn = field(True), field(True)                    # bit-exactness is checked on these
th = torch.rand(nz, dtype=torch.float64, device=dev, generator=g)
a1 = torch.rand(ncell, dtype=torch.float64, device=dev, generator=g)
a2 = torch.rand(ncell, dtype=torch.float64, device=dev, generator=g)
out = torch.empty((a.nt, 2 * ncell), dtype=torch.float64, device=dev)
ref = None
res = []
nbytes = u.numel() * 4 * 2 + out.numel() * 8
for m in [int(s) for s in a.masks.split(',')]:
    _lib.set_option(_lib.NFX_OPT_K2_ALU_MASK, m)
    for _ in range(3):
        nemoflux_gpu.edgeFluxAssemble(u, v, th, a1, a2, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.reps):
        nemoflux_gpu.edgeFluxAssemble(u, v, th, a1, a2, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    nemoflux_gpu.edgeFluxAssemble(un, vn, th, a1, a2, out=out)
    bits = out.view(torch.int64).clone()
    if ref is None:
        ref = bits
    same = bool(torch.equal(bits, ref))
    res.append({'mask': m, 'ms': round(ms, 4), 'GBps': round(nbytes / ms / 1e6, 1), 'bit_identical_to_mask0': same})
    print(json.dumps(res[-1]), flush=True)
for fill in (1.0e20,):                               # fill-value variant of the masking
    uf = torch.nan_to_num(un, nan=fill, posinf=float('inf'), neginf=float('-inf'))
    vf = torch.nan_to_num(vn, nan=fill, posinf=float('inf'), neginf=float('-inf'))
    outs = []
    for m in (0, 9):
        _lib.set_option(_lib.NFX_OPT_K2_ALU_MASK, m)
        nemoflux_gpu.edgeFluxAssemble(uf, vf, th, a1, a2, fill=fill, out=out)
        outs.append(out.view(torch.int64).clone())
    print(json.dumps({'fill': fill, 'mask9_vs_mask0_identical': bool(torch.equal(outs[0], outs[1])),
                      'fill_vs_nan_identical': bool(torch.equal(outs[0], ref))}), flush=True)
