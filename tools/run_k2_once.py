"""run the stand-alone K2 launch a few times on random data (for ncu captures): --dtype f32|f64 --shape C3 --nt 32"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemoflux_b200 import _lib, nemoflux_gpu, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--dtype', default='f32')
ap.add_argument('--shape', default='C3')
ap.add_argument('--nt', type=int, default=32)
ap.add_argument('--reps', type=int, default=4)
ap.add_argument('--alu-mask', type=int, default=-1)
a = ap.parse_args()
cfg = synth.CONFIGS[a.shape]
_lib.set_option(_lib.NFX_OPT_K2_ALU_MASK, a.alu_mask)
ncell, nz = cfg['nx'] * cfg['ny'], cfg['nz']
dev = torch.device('cuda', 0)
tdt = torch.float32 if a.dtype == 'f32' else torch.float64
u = torch.randn((a.nt, nz, ncell), dtype=tdt, device=dev)
v = torch.randn((a.nt, nz, ncell), dtype=tdt, device=dev)
th = torch.rand(nz, dtype=torch.float64, device=dev)
a1 = torch.rand(ncell, dtype=torch.float64, device=dev)
a2 = torch.rand(ncell, dtype=torch.float64, device=dev)
out = torch.empty((a.nt, 2 * ncell), dtype=torch.float64, device=dev)
for _ in range(a.reps):
    nemoflux_gpu.edgeFluxAssemble(u, v, th, a1, a2, out=out)
torch.cuda.synchronize()
print('ok', float(out[0, 0]))
