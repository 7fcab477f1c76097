"""fluxexact -- drop-in for nemoflux/fluxexact.py: the analytic flux of a stream-function flow between the
first and the last target point, sum_k (psi_k(B) - psi_k(A)) * thickness_k per time step (fluxexact.py:21-46).

    python -m nemoflux_b200.fluxexact --potentialFunction="x" --lonLatPointsStr="(-180,-70),(180,40)"
"""
import argparse
import ast

import numpy

from .datagen import eval_stream_function


def exactFlux(potentialFunction, xyVals, zmin=0., zmax=1., nz=5, nt=1):
    dz = (zmax - zmin) / float(nz)
    zhalf = numpy.array([zmin + (k + 0.5) * dz for k in range(nz)])
    ztop = numpy.array([zmin + (k + 0) * dz for k in range(nz)])
    zbot = numpy.array([zmin + (k + 1) * dz for k in range(nz)])
    thickness = -(ztop - zbot)
    xyBeg, xyEnd = xyVals[0, :], xyVals[-1, :]
    out = numpy.zeros(nt)
    for t in range(nt):
        flux = 0.
        for k in range(nz):
            phiA = eval_stream_function(potentialFunction, xyBeg[0], xyBeg[1], zhalf[k], t, nt, zmin, zmax, k)
            phiB = eval_stream_function(potentialFunction, xyEnd[0], xyEnd[1], zhalf[k], t, nt, zmin, zmax, k)
            flux += (phiB - phiA) * thickness[k]
        out[t] = flux
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description='Exact flux of a stream function flow')
    ap.add_argument('-p', '--potentialFunction', default="(cos(t*2*pi/nt)+2)*(0.5*(y/180)**2 + sin(2*pi*x/360))")
    ap.add_argument('--zmin', type=float, default=0.)
    ap.add_argument('--zmax', type=float, default=1.)
    ap.add_argument('--nz', type=int, default=5)
    ap.add_argument('--nt', type=int, default=1)
    ap.add_argument('-d', '--deltaDeg', default='(0.,0.)')
    ap.add_argument('-l', '--lonLatPointsStr', required=True, help='target points "(lon0, lat0), (lon1, lat1),..."')
    a = ap.parse_args(argv)
    xyVals = numpy.array(ast.literal_eval(a.lonLatPointsStr), numpy.float64)
    print(f'zmin/zmax = {a.zmin}/{a.zmax}')
    print(f'beg/end target points: {xyVals[0, :]} {xyVals[-1, :]}')
    print('time_index                 flux')
    res = exactFlux(a.potentialFunction, xyVals, a.zmin, a.zmax, a.nz, a.nt)
    for t, flux in enumerate(res):
        print(f'{t:10d} {flux:20.10g}')
    return res


if __name__ == '__main__':
    main()
