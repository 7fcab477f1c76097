"""Horizontal grid ingest -- mirror of nemoflux/horizgrid.py backed by nemoflux_gpu.Grid.

T.nc ``bounds_lon`` / ``bounds_lat`` (y, x, 4) -> points (ny*nx, 4, 3) = (lon, lat, 0) with vertex order
SW, SE, NE, NW (horizgrid.py:17-22) -> device grid (mint.Grid in the reference, horizgrid.py:23-24).
"""
import argparse
import re

import numpy

from . import ncio


class HorizGrid(object):

    def __init__(self, tFile=None, bounds_lon=None, bounds_lat=None, device_grid=True):
        if tFile is not None:
            with ncio.open_dataset(tFile) as nc:
                bounds_lat = nc['bounds_lat'][:]
                bounds_lon = nc['bounds_lon'][:]
        bounds_lat = numpy.asarray(bounds_lat, numpy.float64)
        bounds_lon = numpy.asarray(bounds_lon, numpy.float64)
        ny, nx, nvertex = bounds_lat.shape
        if nvertex != 4:
            raise RuntimeError(f'ERROR: cells must have 4 vertices, got {nvertex}')
        self.ny, self.nx = ny, nx
        numCells = ny * nx
        self.points = numpy.zeros((ny, nx, nvertex, 3), numpy.float64)
        self.points[..., 0] = bounds_lon
        self.points[..., 1] = bounds_lat
        self.points = self.points.reshape((numCells, nvertex, 3))
        self.grid = None
        if device_grid:
            from .nemoflux_gpu import Grid
            self.grid = Grid()
            self.grid.setPoints(self.points)
            self.grid.setCGridShape(ny, nx)

    def getMintGrid(self):
        return self.grid

    getGrid = getMintGrid

    def getNumCells(self):
        return self.grid.getNumberOfCells() if self.grid is not None else self.points.shape[0]

    def getPoints(self):
        return self.points

    def getPoint(self, cellId, vertex):
        return self.points[cellId, vertex, :]

    def dump(self, fileName):
        """legacy-VTK (ASCII, unstructured quads) dump of the grid, what mint.Grid.dump writes"""
        n = self.points.shape[0]
        with open(fileName, 'w') as f:
            f.write('# vtk DataFile Version 4.2\nnemoflux grid\nASCII\nDATASET UNSTRUCTURED_GRID\n')
            f.write(f'POINTS {4 * n} double\n')
            numpy.savetxt(f, self.points.reshape(-1, 3), fmt='%.17g')
            f.write(f'CELLS {n} {5 * n}\n')
            ids = numpy.arange(4 * n).reshape(n, 4)
            numpy.savetxt(f, numpy.concatenate([numpy.full((n, 1), 4), ids], 1), fmt='%d')
            f.write(f'CELL_TYPES {n}\n')
            numpy.savetxt(f, numpy.full(n, 9), fmt='%d')


def main(argv=None):
    ap = argparse.ArgumentParser(description='Create grid')
    ap.add_argument('-t', '--tFile', required=True, help='netCDF file containing t grid data')
    a = ap.parse_args(argv)
    gr = HorizGrid(a.tFile, device_grid=False)
    gr.dump(re.sub('.nc', '.vtk', a.tFile))


if __name__ == '__main__':
    main()
