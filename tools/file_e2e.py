"""File-to-series end to end, the reference's real workflow (fluxplot.py): NetCDF T/U/V files on disk -> (nt, M)
flux series, GPU tool vs the CPU oracle restatement reading the same files.

    python tools/file_e2e.py --workload C3 --nt 8 --dir /tmp/nfx_e2e --out gpurun_out/file_e2e.json
"""
import argparse
import json
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nemoflux_b200 import ncio, synth  # noqa: E402


def write_files(syn, nt, d, dtype, fmt='classic'):
    os.makedirs(d, exist_ok=True)
    w = ncio.Writer(os.path.join(d, 'T.nc'))
    for name, n in (('z', syn.nz), ('y', syn.ny), ('x', syn.nx), ('nvertex', 4), ('axis_nbounds', 2)):
        w.createDimension(name, n)
    w.createVariable('deptht_bounds', 'float64', ('z', 'axis_nbounds'), data=numpy.stack([syn.ztop, syn.zbot], 1))
    w.createVariable('bounds_lat', 'float64', ('y', 'x', 'nvertex'), data=syn.bounds_lat)
    w.createVariable('bounds_lon', 'float64', ('y', 'x', 'nvertex'), data=syn.bounds_lon)
    w.close()
    if fmt == 'hdf5':      # NetCDF-4 style container (little-endian, contiguous), written by the test helper
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
        import h5build
        for fname, vname, k in (('U.nc', 'uo', 0), ('V.nc', 'vo', 1)):
            a = numpy.stack([syn.uv_host(t)[k].astype(dtype) for t in range(nt)])
            a[numpy.isnan(a)] = 1.e20
            h5build.write(os.path.join(d, fname), {vname: dict(data=a, attrs={'_FillValue': a.dtype.type(1.e20)})})
        return
    for fname, vname, k in (('U.nc', 'uo', 0), ('V.nc', 'vo', 1)):
        w = ncio.Writer(os.path.join(d, fname))
        for name, n in (('t', nt), ('z', syn.nz), ('y', syn.ny), ('x', syn.nx)):
            w.createDimension(name, n)
        var = w.createVariable(vname, dtype, ('t', 'z', 'y', 'x'), fill_value=1.e20)
        for t in range(nt):
            a = syn.uv_host(t)[k].astype(dtype)
            a[numpy.isnan(a)] = 1.e20                       # land as NEMO writes it
            var[t] = a
        w.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='C3')
    ap.add_argument('--nt', type=int, default=8)
    ap.add_argument('--dtype', default='float64')
    ap.add_argument('--dir', default='/tmp/nfx_e2e')
    ap.add_argument('--chunk-steps', default='0', help='comma list; the first is the reported one (0 = default)')
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--out', default='gpurun_out/file_e2e.json')
    ap.add_argument('--format', default='classic', choices=['classic', 'hdf5'])
    ap.add_argument('--skip-cpu', action='store_true')
    a = ap.parse_args()
    syn = synth.make(a.workload)
    t0 = time.perf_counter()
    write_files(syn, a.nt, a.dir, a.dtype, a.format)
    t_write = time.perf_counter() - t0
    T, U, V = (os.path.join(a.dir, f) for f in ('T.nc', 'U.nc', 'V.nc'))
    units = syn.units_per_step() * a.nt

    # GPU tool
    import torch
    from nemoflux_b200.field import Field
    torch.cuda.init()
    t0 = time.perf_counter()
    fld = Field(T, U, V, syn.transects, verbose=False)
    t_init = time.perf_counter() - t0
    sweep = {}
    for cs in [int(x) for x in a.chunk_steps.split(',')]:
        fld.fluxSeries(chunk_steps=cs)                    # warm-up (page cache, allocations)
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            s_gpu = fld.fluxSeries(chunk_steps=cs)
            ts.append(time.perf_counter() - t0)
        sweep[cs] = dict(median_s=float(numpy.median(ts)), min_s=min(ts))
        print('chunk_steps', cs, sweep[cs], flush=True)
    t_gpu = sweep[int(a.chunk_steps.split(',')[0])]['median_s']

    if a.skip_cpu:
        fbytes = int(os.path.getsize(U) + os.path.getsize(V))
        res = dict(workload=a.workload, nt=a.nt, dtype=a.dtype, format=a.format, transects=len(syn.transects),
                   file_bytes=fbytes, write_s=t_write, chunk_sweep=sweep,
                   gpu=dict(init_s=t_init, series_s=t_gpu, units_per_s=units / t_gpu, file_gbs=fbytes / t_gpu / 1e9),
                   series_abs_max=float(numpy.abs(s_gpu).max()))
        print(json.dumps(res, indent=1))
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(res, open(a.out, 'w'), indent=1)
        for f in (T, U, V):
            os.unlink(f)
        return
    # CPU: the reference's loop restated (oracle), reading the same files through the same reader
    from oracle import oracle as O
    t0 = time.perf_counter()
    og = O.Grid(syn.points)
    plis = []
    for xyz in syn.transects:
        p = O.PolylineIntegral(og)
        p.computeWeights(xyz)
        plis.append(p)
    arc = O.arc_lengths(syn.points)
    t_cpu_init = time.perf_counter() - t0
    ncU, ncV = ncio.open_dataset(U), ncio.open_dataset(V)
    t0 = time.perf_counter()
    s_cpu = numpy.zeros_like(s_gpu)
    for t in range(a.nt):
        u = ncU['uo'][t].astype(numpy.float64)            # decodes 1e20 -> NaN like xarray
        v = ncV['vo'][t].astype(numpy.float64)
        Uz = O.read_field(u, syn.thickness)
        Vz = O.read_field(v, syn.thickness)
        iV, _, _ = O.integrated_flux(Uz, Vz, arc)
        for m, p in enumerate(plis):
            s_cpu[t, m] = p.getIntegral(iV)
    t_cpu = time.perf_counter() - t0
    err = float(numpy.abs(s_gpu - s_cpu).max() / numpy.abs(s_cpu).max())
    res = dict(workload=a.workload, nt=a.nt, dtype=a.dtype, transects=len(syn.transects),
               file_bytes=int(os.path.getsize(U) + os.path.getsize(V)), write_s=t_write,
               chunk_sweep=sweep,
               gpu=dict(init_s=t_init, series_s=t_gpu, units_per_s=units / t_gpu,
                        file_gbs=(os.path.getsize(U) + os.path.getsize(V)) / t_gpu / 1e9),
               cpu=dict(init_s=t_cpu_init, series_s=t_cpu, units_per_s=units / t_cpu, cores=os.cpu_count()),
               speedup_series=t_cpu / t_gpu, max_rel_diff=err)
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, 'w'), indent=1)
    for f in (T, U, V):
        os.unlink(f)


if __name__ == '__main__':
    main()
