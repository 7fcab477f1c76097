"""Horizontal grid ingest: T.nc cell bounds -> the device grid K1 works on.

Interface of nemoflux/horizgrid.py (HorizGrid(tFile), getMintGrid, getNumCells, getPoints, getPoint, dump),
with nemoflux_gpu.Grid in the place of mint.Grid (horizgrid.py:23-24).  ``bounds_lon`` / ``bounds_lat`` are
(y, x, 4) with the vertices SW, SE, NE, NW; the grid wants one private copy of the 4 vertices per cell,
(ny*nx, 4, 3) = (lon, lat, 0), cell id = j*nx + i (horizgrid.py:17-22).
"""
import argparse

import numpy

from . import ncio

VTK_QUAD = 9


def cell_vertices(bounds_lon, bounds_lat):
    """(ny, nx, 4) lon and lat bounds -> (ny*nx, 4, 3) vertex array with z = 0"""
    lon = numpy.asarray(bounds_lon, numpy.float64)
    lat = numpy.asarray(bounds_lat, numpy.float64)
    if lon.shape != lat.shape or lon.ndim != 3:
        raise RuntimeError(f'ERROR: bounds_lon {lon.shape} and bounds_lat {lat.shape} must both be (y, x, 4)')
    if lon.shape[2] != 4:
        raise RuntimeError(f'ERROR: cells must have 4 vertices, got {lon.shape[2]}')
    xyz = numpy.stack([lon, lat, numpy.zeros_like(lon)], axis=-1)
    return numpy.ascontiguousarray(xyz.reshape(-1, 4, 3))


class HorizGrid(object):

    def __init__(self, tFile=None, bounds_lon=None, bounds_lat=None, device_grid=True):
        """from a T file, or from the two bounds arrays; device_grid=False keeps everything on the host (dump)"""
        if tFile is not None:
            with ncio.open_dataset(tFile) as nc:
                bounds_lon, bounds_lat = nc['bounds_lon'][:], nc['bounds_lat'][:]
        self.ny, self.nx = numpy.shape(bounds_lat)[:2]
        self.points = cell_vertices(bounds_lon, bounds_lat)
        self.grid = None
        if device_grid:
            from .nemoflux_gpu import Grid
            self.grid = Grid()
            self.grid.setPoints(self.points)
            self.grid.setCGridShape(self.ny, self.nx)

    def getMintGrid(self):
        return self.grid

    getGrid = getMintGrid

    def getNumCells(self):
        return len(self.points) if self.grid is None else self.grid.getNumberOfCells()

    def getPoints(self):
        return self.points

    def getPoint(self, cellId, vertex):
        return self.points[cellId, vertex]

    def dump(self, fileName):
        """legacy VTK file (ASCII, unstructured quads with private vertices), the format mint.Grid.dump writes"""
        ncells = len(self.points)
        connectivity = numpy.empty((ncells, 5), numpy.int64)
        connectivity[:, 0] = 4
        connectivity[:, 1:] = numpy.arange(4 * ncells).reshape(ncells, 4)
        with open(fileName, 'w') as out:
            out.write('# vtk DataFile Version 4.2\nnemoflux grid\nASCII\nDATASET UNSTRUCTURED_GRID\n')
            out.write(f'POINTS {4 * ncells} double\n')
            numpy.savetxt(out, self.points.reshape(-1, 3), fmt='%.17g')
            out.write(f'CELLS {ncells} {connectivity.size}\n')
            numpy.savetxt(out, connectivity, fmt='%d')
            out.write(f'CELL_TYPES {ncells}\n')
            numpy.savetxt(out, numpy.full(ncells, VTK_QUAD), fmt='%d')


def main(argv=None):
    """the reference's command line (horizgrid.py:45-52): T.nc -> T.vtk next to it"""
    ap = argparse.ArgumentParser(description='Create grid')
    ap.add_argument('-t', '--tFile', required=True, help='netCDF file containing t grid data')
    a = ap.parse_args(argv)
    target = a.tFile[:-3] + '.vtk' if a.tFile.endswith('.nc') else a.tFile + '.vtk'
    HorizGrid(a.tFile, device_grid=False).dump(target)


if __name__ == '__main__':
    main()
