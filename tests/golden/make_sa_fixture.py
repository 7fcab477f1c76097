"""Writes tests/golden/sa_T_grid.npz from the reference's only data file with a real NEMO grid,
/root/reference/data/sa/T.nc (NetCDF-4/HDF5, a 100 x 100 x 75 window of ORCA025 south of Africa cut with
subsetNEMO.py), read with the package's own HDF5 reader (nemoflux_b200/h5lite.py through ncio).  Run in the build
container (the file does not travel to the GPU box):

    python tests/golden/make_sa_fixture.py
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from nemoflux_b200 import ncio  # noqa: E402

with ncio.open_dataset('/root/reference/data/sa/T.nc') as nc:
    out = {name: nc[name][:] for name in ('bounds_lon', 'bounds_lat', 'deptht_bounds', 'deptht')}
    dims = {name: '|'.join(nc[name].dimensions) for name in out}
numpy.savez_compressed(os.path.join(HERE, 'sa_T_grid.npz'), dims=numpy.array(sorted(dims.items())), **out)
print({k: (v.shape, v.dtype) for k, v in out.items()}, dims)
