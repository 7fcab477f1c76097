"""Cut an index window out of a set of NEMO T/U/V files -- the data-preparation tool of the reference
(nemoflux/subsetNEMO.py:6-93), same flags: --tfile --ufile --vfile --outputdir --jmin --jmax --imin --imax.

Writes T.nc (bounds_lon, bounds_lat cut to [jmin:jmax, imin:imax]; deptht, deptht_bounds whole), U.nc and V.nc
(uo / vo cut in their last two axes, fill value kept; the time variables whole) into `outputdir`, with the
dimensions x and y shrunk accordingly and all variable attributes copied.  Differences from the reference, on
purpose: variables a file does not have (datagen's mock files carry no `deptht` and no time axis) are skipped with a
note instead of a KeyError, and the output is NetCDF classic (64-bit offset, uncompressed) when the netCDF4 package
is absent -- which is what nemoflux_b200.ncio reads fastest (memory-mapped, bytes uploaded as stored).
"""
import argparse
import datetime
import os
import sys

import numpy

from . import ncio

T_VARS = ('bounds_lon', 'bounds_lat', 'deptht', 'deptht_bounds')
TIME_VARS = ('time_counter', 'time_centered', 'time_centered_bounds')
SLAB_BYTES = 256 << 20      # copy big variables in slabs of their first axis


def _dimension_sizes(ds):
    sizes = {}
    for var in ds.variables.values():
        for name, n in zip(var.dimensions, var.shape):
            sizes[name] = n
    return sizes


def _copy(ds, out, names, window, verbose):
    """copy the variables `names` of dataset ds into the Writer out; window = (jmin, jmax, imin, imax) is applied
    to the last two axes of every variable whose last two dimensions are (y, x) -- or (y, x, nvertex) bounds"""
    jmin, jmax, imin, imax = window
    for name, n in _dimension_sizes(ds).items():
        out.createDimension(name, imax - imin if name == 'x' else jmax - jmin if name == 'y' else n)
    for name in names:
        if name not in ds:
            if verbose:
                print(f'variable {name} is not in {ds.path}: skipped')
            continue
        var = ds[name]
        if verbose:
            print(f'creating variable {name}')
        attrs = {k: v for k, v in var.attrs.items() if k != '_FillValue'}
        fill = var.attrs.get('_FillValue', None)
        if fill is not None:
            fill = numpy.asarray(fill).reshape(-1)[0]
        dst = out.createVariable(name, var.dtype.newbyteorder('='), var.dimensions, fill_value=fill, attrs=attrs)
        dims = var.dimensions
        if 'y' in dims and 'x' in dims and dims.index('x') == dims.index('y') + 1:
            ay = dims.index('y')
            cut = (slice(None),) * ay + (slice(jmin, jmax), slice(imin, imax))
        else:
            cut = ()
        if len(var.shape) == 0:
            dst.assignValue(var.raw()) if hasattr(dst, 'assignValue') else dst.__setitem__(Ellipsis, var.raw())
            continue
        if cut and cut[0] == slice(jmin, jmax):          # (y, x, ...) itself is the first axis: one piece
            dst[:] = var.raw(cut)
            continue
        row_bytes = max(1, int(numpy.prod(var.shape[1:], dtype=numpy.int64)) * var.dtype.itemsize)
        step = max(1, SLAB_BYTES // row_bytes)
        for a in range(0, var.shape[0], step):
            b = min(var.shape[0], a + step)
            dst[a:b] = var.raw((slice(a, b),) + cut[1:])


def subset(tfile='', ufile='', vfile='', outputdir='./', jmin=0, jmax=0, imin=0, imax=0, verbose=True):
    if not (0 <= jmin < jmax and 0 <= imin < imax):
        raise ValueError(f'empty index window j {jmin}:{jmax}, i {imin}:{imax}')
    stamp = f'generated on {datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S")}'
    jobs = [(tfile, 'T.nc', T_VARS), (ufile, 'U.nc', TIME_VARS + ('uo',)), (vfile, 'V.nc', TIME_VARS + ('vo',))]
    os.makedirs(outputdir, exist_ok=True)
    for src, target, names in jobs:
        if not src:
            continue
        if verbose:
            print(f'{target[0]} file: {src}')
        with ncio.open_dataset(src) as ds:
            sizes = _dimension_sizes(ds)
            if jmax > sizes.get('y', jmax) or imax > sizes.get('x', imax):
                raise ValueError(f'{src}: window j {jmin}:{jmax}, i {imin}:{imax} exceeds the grid '
                                 f'({sizes.get("y")}, {sizes.get("x")})')
            out = ncio.Writer(os.path.join(outputdir, target))
            out.setAttr('command', ' '.join(sys.argv))
            out.setAttr('timestamp', stamp)
            for key, val in ds.attrs.items():           # e.g. earthRadius (datagen.py:196)
                if key not in ('command', 'timestamp'):
                    out.setAttr(key, val)
            _copy(ds, out, names, (jmin, jmax, imin, imax), verbose)
            out.close()


def main(argv=None):
    ap = argparse.ArgumentParser(description='subset nemo data')
    ap.add_argument('-t', '--tfile', default='', help='name of the netCDF file containing T cell grid data')
    ap.add_argument('-u', '--ufile', default='', help='name of the netCDF file containing u data')
    ap.add_argument('-v', '--vfile', default='', help='name of the netCDF file containing v data')
    ap.add_argument('-o', '--outputdir', default='./', help='output directory, the files are saved as T.nc, U.nc and V.nc')
    ap.add_argument('--jmin', type=int, required=True, help='min j index')
    ap.add_argument('--jmax', type=int, required=True, help='max j index')
    ap.add_argument('--imin', type=int, required=True, help='min i index')
    ap.add_argument('--imax', type=int, required=True, help='max i index')
    a = ap.parse_args(argv)
    subset(a.tfile, a.ufile, a.vfile, a.outputdir, a.jmin, a.jmax, a.imin, a.imax)


if __name__ == '__main__':
    main()
