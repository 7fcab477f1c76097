"""K1 (computeWeights) at full size on one GPU: time, sizes, and a bit-exact check of a subset of transects
against the C oracle.

    python tools/k1_scale.py --workload C5 --check 24
"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='C5')
    ap.add_argument('--check', type=int, default=24)
    ap.add_argument('--out', default='gpurun_out/k1_scale.json')
    a = ap.parse_args()
    t0 = time.perf_counter()
    syn = synth.make(a.workload)
    t_syn = time.perf_counter() - t0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g = nemoflux_gpu.Grid()
    g.setPoints(syn.points)
    g.setCGridShape(syn.ny, syn.nx)
    torch.cuda.synchronize()
    t_up = time.perf_counter() - t0
    p = nemoflux_gpu.PolylineIntegral()
    t0 = time.perf_counter()
    p.build(g)
    torch.cuda.synchronize()
    t_loc = time.perf_counter() - t0
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        p.computeWeights(syn.transects)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    sub = p.getSubsegments()
    mp = p.getMap()
    per = numpy.diff(sub['offsets'])
    res = dict(workload=a.workload, ncell=syn.ncell, transects=len(syn.transects),
               segments=int(sum(len(t) - 1 for t in syn.transects)), subsegments=int(sub['cell'].size),
               map_entries=int(mp['keys'].size), max_subsegments_per_transect=int(per.max()),
               synth_s=t_syn, upload_s=t_up, locator_s=t_loc, compute_weights_s=times,
               gpu_mem_gb=torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9)
    # coverage diagnostic: sum (tb-ta)*coeff per transect = number of segments when fully inside the grid
    tot = numpy.add.reduceat((sub['tb'] - sub['ta']) * sub['coeff'], sub['offsets'][:-1][per > 0])
    nseg = numpy.array([len(t) - 1 for t in syn.transects])[per > 0]
    res['coverage_max_dev'] = float(numpy.abs(tot - nseg).max())
    # oracle subset
    t0 = time.perf_counter()
    og = O.Grid(syn.points)
    ok, n = True, 0
    idx = numpy.linspace(0, len(syn.transects) - 1, min(a.check, len(syn.transects))).astype(int)
    for m in idx:
        op = O.PolylineIntegral(og)
        op.computeWeights(syn.transects[m])
        aa, bb = sub['offsets'][m], sub['offsets'][m + 1]
        o = op.subsegs
        same = (bb - aa == len(o)) and numpy.array_equal(sub['cell'][aa:bb], o['cell']) and \
            numpy.array_equal(sub['w'][aa:bb].view(numpy.int64), o['w'].view(numpy.int64)) and \
            numpy.array_equal(sub['ta'][aa:bb], o['ta'])
        keys, ws = op.merged_map()
        ma, mb = mp['offsets'][m], mp['offsets'][m + 1]
        same = same and numpy.array_equal(mp['keys'][ma:mb], keys) and \
            numpy.array_equal(mp['w'][ma:mb].view(numpy.int64), ws.view(numpy.int64))
        ok = ok and bool(same)
        n += len(o)
    res.update(oracle_checked_transects=int(len(idx)), oracle_checked_subsegments=int(n), bit_exact=bool(ok),
               oracle_s=time.perf_counter() - t0)
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
