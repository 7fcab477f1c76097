"""Sweep the K2 (edge-flux assembly) kernel variants on one GPU: CUDA-event time and achieved GB/s.

    python tools/k2_sweep.py [--shapes C3,C4,C5] [--gb 12] [--reps 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemoflux_b200 import _lib, nemoflux_gpu, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shapes', default='C3,C4,C5')
    ap.add_argument('--gb', type=float, default=12.0, help='u+v bytes resident per case')
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--dtypes', default='f64,f32')
    ap.add_argument('--out', default='gpurun_out/k2_sweep.json')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs'] if os.path.exists('MEASURED_PEAKS.json') else 6650.
    results = []
    for name in a.shapes.split(','):
        cfg = synth.CONFIGS[name]
        nx, ny, nz = cfg['nx'], cfg['ny'], cfg['nz']
        ncell = nx * ny
        for dt in a.dtypes.split(','):
            tdt, es = (torch.float64, 8) if dt == 'f64' else (torch.float32, 4)
            nt = max(2, int(a.gb * 1e9 / (2 * es * nz * ncell)))
            u = torch.randn((nt, nz, ncell), dtype=tdt, device=dev)
            v = torch.randn((nt, nz, ncell), dtype=tdt, device=dev)
            th = torch.rand(nz, dtype=torch.float64, device=dev)
            a1 = torch.rand(ncell, dtype=torch.float64, device=dev)
            a2 = torch.rand(ncell, dtype=torch.float64, device=dev)
            out = torch.empty((nt, 2 * ncell), dtype=torch.float64, device=dev)
            nbytes = 2.0 * es * nz * ncell * nt
            ref = None
            for variant, unroll, block in [(1, 5, 256), (1, 3, 256), (1, 8, 256), (1, 15, 256), (1, 5, 128), (1, 8, 128),
                                           (1, 15, 128), (1, 5, 512), (1, 3, 512), (3, 5, 256), (3, 8, 256), (3, 15, 256),
                                           (3, 15, 128), (3, 5, 512)] + [(2, c, 0) for c in range(9)]:
                try:
                    _lib.set_option(_lib.NFX_OPT_K2_VARIANT, variant)
                    _lib.set_option(_lib.NFX_OPT_K2_UNROLL, unroll)
                    _lib.set_option(_lib.NFX_OPT_K2_BLOCK, block)
                    for _ in range(3):
                        nemoflux_gpu.edgeFluxAssemble(u, v, th, a1, a2, out=out)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(a.reps):
                        nemoflux_gpu.edgeFluxAssemble(u, v, th, a1, a2, out=out)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / a.reps
                    chk = out.double().sum().item()
                    if ref is None:
                        ref = out.clone()
                        same = True
                    else:
                        same = bool(torch.equal(ref, out))
                    r = dict(shape=name, dtype=dt, nt=nt, variant=variant, unroll=unroll, block=block, ms=ms,
                             gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / peak, same_as_first=same)
                except Exception as e:  # unsupported combination
                    r = dict(shape=name, dtype=dt, variant=variant, unroll=unroll, block=block, error=str(e)[:100])
                results.append(r)
                print(json.dumps(r), flush=True)
            del u, v, out
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(results, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
