"""fluxplot -- drop-in for nemoflux/fluxplot.py: flux time series of every transect (same flags).

    python -m nemoflux_b200.fluxplot -t T.nc -u U.nc -v V.nc -i "../data/nz/S*.txt" [-s] [--output series.csv]

The reference loops ``for itime: fld.update(); for pli: pli.getIntegral(...)`` (fluxplot.py:51-59); here the
whole (nt, M) series comes from Field.fluxSeries() (K2+K3 over chunks of time steps).  The series is printed
and optionally written as CSV; it is plotted only when matplotlib is importable and --plot is given.
"""
import argparse
import os

import numpy

from .field import Field
from .fluxviz import parseTransects


def main(argv=None):
    ap = argparse.ArgumentParser(description='Flux time series')
    ap.add_argument('-t', '--tFile', required=True)
    ap.add_argument('-u', '--uFile', required=True)
    ap.add_argument('-v', '--vFile', required=True)
    ap.add_argument('-l', '--lonLatPoints', default='')
    ap.add_argument('-i', '--iFiles', default='')
    ap.add_argument('-s', '--sverdrup', action='store_true')
    ap.add_argument('-o', '--output', default='', help='write the series as CSV (time, one column per transect)')
    ap.add_argument('--plot', action='store_true')
    ap.add_argument('--meshFile', default='', help='optional mesh_mask file: e3u_0/e3v_0 (and e2u/e1v) scale factors')
    a = ap.parse_args(argv)
    pts, names = parseTransects(a.lonLatPoints, a.iFiles)
    print(f'target points:\n {pts}')
    fld = Field(a.tFile, a.uFile, a.vFile, pts, a.sverdrup, meshFile=a.meshFile or None)
    series = fld.fluxSeries()
    timeVals = [fld.timeObj.getTimeAsDate(i) for i in range(fld.nt)]
    units = 'Sv' if a.sverdrup else 'A m^2/s'
    labels = [os.path.basename(n) for n in names]
    print('time,' + ','.join(labels) + f'   [{units}]')
    for t, row in zip(timeVals, series):
        print(f'{t},' + ','.join(f'{x:.15g}' for x in row))
    if a.output:
        with open(a.output, 'w') as f:
            f.write('time,' + ','.join(labels) + '\n')
            for t, row in zip(timeVals, series):
                f.write(f'{t},' + ','.join(repr(float(x)) for x in row) + '\n')
    if a.plot:
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            print('matplotlib is not installed: no plot')
        else:
            lineTypes = ['b-', 'm--', 'c-.', 'r:', 'g-', 'k--']
            for i in range(series.shape[1]):
                plt.plot(timeVals, series[:, i], lineTypes[i % len(lineTypes)])
            if len(labels) > 1:
                plt.legend(labels)
            plt.title('Water flow')
            plt.ylabel(units)
            plt.show()
    return numpy.asarray(series)


if __name__ == '__main__':
    main()
