"""Mock NEMO data generator -- drop-in for nemoflux/datagen.py (same flags, same files, same maths).

    python -m nemoflux_b200.datagen --streamFunction="x"
    python -m nemoflux_b200.datagen --streamFunction="cos(2*pi*y/360) + sin(2*pi*x/360)" \
           --nx=360 --ny=180 --deltaDeg="20,30"

writes <prefix>T.nc (bounds_lon, bounds_lat (y,x,4), deptht_bounds (z,2)), <prefix>U.nc (uo (t,z,y,x)) and
<prefix>V.nc (vo (t,z,y,x)), float64, _FillValue 1e20 -- the conventions of datagen.py:168-208.

Formulas follow /root/reference/nemoflux/datagen.py: vertical axis :35-40 (ztop=(k+1)dz, zbot=(k+2)dz),
horizontal axes :42-66 (the reference spaces y with dx, :49 -- kept, ``dy`` overrides it for grids
where that is not wanted), stream function on the UN-rotated logical nodes :69-82, u = dpsi/ds on the
east edge and v = -dpsi/ds on the north edge :85-113, pole displacement :116-166.  Vectorised; the
reference loops vertex by vertex with math.asin/atan2, so coordinates can differ in the last ulp.
"""
import argparse
import math  # noqa: F401  (visible to stream-function expressions, datagen.py:4)

import numpy
from numpy import pi, cos, sin, arctan2, arctan  # noqa: F401  (datagen.py:3)

from . import geo
from . import ncio

REAL = 'float64'      # datagen.py:9

_EXPR_NAMES = dict(pi=pi, cos=cos, sin=sin, arctan2=arctan2, arctan=arctan, numpy=numpy, math=math)


def eval_stream_function(expr, x, y, z, t, nt, zmin, zmax, k=0):
    """evaluate a stream-function expression with the names datagen.py:70-78 exposes, nothing else"""
    env = dict(_EXPR_NAMES)
    env.update(x=x, y=y, z=z, t=t, nt=nt, A=geo.EARTH_RADIUS, zmin=zmin, zmax=zmax, k=k)
    return eval(expr, {'__builtins__': {'abs': abs, 'min': min, 'max': max, 'float': float, 'int': int}}, env)


def rotate_lonlat(lon, lat, deltaDeg):
    """pole displacement of datagen.py:116-157 applied to arrays of lon/lat in degrees"""
    alpha = numpy.pi * deltaDeg[1] / 180.
    beta = numpy.pi * deltaDeg[0] / 180.
    ca, sa, cb, sb = numpy.cos(alpha), numpy.sin(alpha), numpy.cos(beta), numpy.sin(beta)
    m = numpy.dot(numpy.array([[cb, sb, 0.], [-sb, cb, 0.], [0., 0., 1.]]),
                  numpy.array([[ca, 0., sa], [0., 1., 0.], [-sa, 0., ca]]))
    the = numpy.pi * lat / 180.
    lam = numpy.pi * lon / 180.
    rho = numpy.cos(the)
    xo, yo, zo = rho * numpy.cos(lam), rho * numpy.sin(lam), numpy.sin(the)
    xn = (m[0, 0] * xo + m[0, 1] * yo) + m[0, 2] * zo
    yn = (m[1, 0] * xo + m[1, 1] * yo) + m[1, 2] * zo
    zn = (m[2, 0] * xo + m[2, 1] * yo) + m[2, 2] * zo
    return 180. * numpy.arctan2(yn, xn) / numpy.pi, 180. * numpy.arcsin(numpy.clip(zn, -1., 1.)) / numpy.pi


class DataGen(object):

    def __init__(self, prefix=''):
        self.prefix = prefix
        self.dy = None

    def setBoundingBox(self, xmin, xmax, ymin, ymax, zmin, zmax):
        self.xmin, self.xmax, self.ymin, self.ymax, self.zmin, self.zmax = xmin, xmax, ymin, ymax, zmin, zmax

    def setSizes(self, nx, ny, nz, nt):
        self.nx, self.ny, self.nz, self.nt = nx, ny, nz, nt

    def build(self):
        self.buildVertical()
        self.buildUniformHorizontal()

    def buildVertical(self):
        dz = (self.zmax - self.zmin) / float(self.nz)
        print(f'zmin/zmax = {self.zmin}/{self.zmax}')
        ks = range(self.nz)
        self.zhalf = numpy.array([self.zmin + (k + 0.5) * dz for k in ks])
        self.ztop = numpy.array([self.zmin + (k + 1) * dz for k in ks])
        self.zbot = numpy.array([self.zmin + (k + 2) * dz for k in ks])

    @staticmethod
    def _cellCorners(nodes):
        c = numpy.zeros((nodes.shape[0] - 1, nodes.shape[1] - 1, 4), numpy.float64)
        c[..., 0], c[..., 1], c[..., 2], c[..., 3] = nodes[:-1, :-1], nodes[:-1, 1:], nodes[1:, 1:], nodes[1:, :-1]
        return c

    def buildUniformHorizontal(self):
        dx = (self.xmax - self.xmin) / float(self.nx)
        dy = dx if self.dy is None else self.dy          # the reference uses dx for both (datagen.py:49)
        x = numpy.array([self.xmin + i * dx for i in range(self.nx + 1)])
        y = numpy.array([self.ymin + j * dy for j in range(self.ny + 1)])
        self.xx, self.yy = numpy.meshgrid(x, y, indexing='xy')
        self.bounds_lon = self._cellCorners(self.xx)
        self.bounds_lat = self._cellCorners(self.yy)

    def rotatePole(self, deltaDeg=(0., 0.)):
        lon, lat = rotate_lonlat(self.bounds_lon, self.bounds_lat, deltaDeg)
        dlon = lon - lon[..., 0:1]          # date line fix relative to vertex 0 (datagen.py:161-166)
        lon = numpy.where(dlon > +270., lon - 360., lon)
        lon = numpy.where(dlon < -270., lon + 360., lon)
        self.bounds_lon, self.bounds_lat = lon, lat

    def potentialAtNodes(self, streamFunction, t, k):
        pot = eval_stream_function(streamFunction, self.xx, self.yy, self.zhalf[k], t, self.nt, self.zmin, self.zmax, k)
        return pot + numpy.zeros_like(self.xx)

    def applyStreamFunction(self, streamFunction):
        self.streamFunction = streamFunction

    def edgeMetric(self):
        """arc lengths of the east (ds21) and north (ds23) edges on the UN-rotated nodes (datagen.py:89-104)"""
        def xyz(lon, lat):
            p = numpy.zeros(lon.shape + (3,), numpy.float64)
            p[..., 0], p[..., 1] = lon, lat
            return geo.lonLat2XYZArray(p, radius=geo.EARTH_RADIUS)
        p1 = xyz(self.xx[:-1, 1:], self.yy[:-1, 1:])
        p2 = xyz(self.xx[1:, 1:], self.yy[1:, 1:])
        p3 = xyz(self.xx[1:, :-1], self.yy[1:, :-1])
        ds21 = geo.getArcLengthArray(p2, p1, radius=geo.EARTH_RADIUS)
        ds23 = numpy.clip(geo.getArcLengthArray(p2, p3, radius=geo.EARTH_RADIUS), 1.e-12, None)
        return ds21, ds23

    def computeUVFromPotential(self):
        ds21, ds23 = self.edgeMetric()
        shape = (self.nt, self.nz, self.ny, self.nx)
        self.u = numpy.zeros(shape, numpy.float64)
        self.v = numpy.zeros(shape, numpy.float64)
        for t in range(self.nt):
            for k in range(self.nz):
                pot = self.potentialAtNodes(self.streamFunction, t, k)
                se, ne, nw = pot[:-1, 1:], pot[1:, 1:], pot[1:, :-1]
                self.u[t, k] = (ne - se) / ds21
                self.v[t, k] = -(ne - nw) / ds23

    def save(self):
        w = ncio.Writer(self.prefix + 'T.nc')
        for name, n in (('z', self.nz), ('y', self.ny), ('x', self.nx), ('nvertex', 4), ('axis_nbounds', 2)):
            w.createDimension(name, n)
        w.createVariable('deptht_bounds', REAL, ('z', 'axis_nbounds'), data=numpy.stack([self.ztop, self.zbot], 1))
        w.createVariable('bounds_lat', REAL, ('y', 'x', 'nvertex'), data=self.bounds_lat)
        w.createVariable('bounds_lon', REAL, ('y', 'x', 'nvertex'), data=self.bounds_lon)
        w.close()
        for fname, vname, data, std in (('U.nc', 'uo', self.u, 'sea_water_x_velocity'),
                                        ('V.nc', 'vo', self.v, 'sea_water_y_velocity')):
            w = ncio.Writer(self.prefix + fname)
            for name, n in (('t', self.nt), ('z', self.nz), ('y', self.ny), ('x', self.nx), ('axis_nbounds', 2)):
                w.createDimension(name, n)
            w.createVariable(vname, REAL, ('t', 'z', 'y', 'x'), fill_value=1.e20,
                             attrs=dict(standard_name=std, units='m/s'), data=data)
            if vname == 'uo':
                w.setAttr('earthRadius', f'earth radius = {geo.EARTH_RADIUS} in metres')
            w.close()


    def saveMeshMask(self, partialCells=0.0):
        """<prefix>mesh_mask.nc with NEMO's names (not part of the reference's datagen; feeds Field(meshFile=...)):
        e3u_0, e3v_0 (t=1, z, y, x): the layer thickness zbot-ztop of every level, the deepest level thinned to
        (1 - partialCells * r(j, i)) of it with a fixed pseudo-random r in [0, 1) -- a partial bottom cell;
        e2u, e1v (t=1, y, x): lengths in metres of the east and north edges of the cells as field.py:170-181
        measures them on the saved bounds, times the earth radius of field.py:12."""
        dz = numpy.asarray(self.zbot, numpy.float64) - numpy.asarray(self.ztop, numpy.float64)
        e3 = numpy.broadcast_to(dz[:, None, None], (self.nz, self.ny, self.nx))
        e3u, e3v = e3.copy(), e3.copy()
        if partialCells > 0.:
            jj, ii = numpy.meshgrid(numpy.arange(self.ny), numpy.arange(self.nx), indexing='ij')
            r = ((jj * 7919 + ii * 104729 + 12345) % 1000) / 1000.0
            e3u[-1] *= 1.0 - partialCells * r
            e3v[-1] *= 1.0 - partialCells * r[::-1, ::-1]
        points = numpy.zeros((self.ny * self.nx, 4, 3), numpy.float64)
        points[:, :, 0] = numpy.asarray(self.bounds_lon).reshape(-1, 4)
        points[:, :, 1] = numpy.asarray(self.bounds_lat).reshape(-1, 4)
        arc = geo.cellArcLengths(points)
        radius = 6371000.0
        w = ncio.Writer(self.prefix + 'mesh_mask.nc')
        for name, n in (('t', 1), ('z', self.nz), ('y', self.ny), ('x', self.nx)):
            w.createDimension(name, n)
        w.createVariable('e3u_0', REAL, ('t', 'z', 'y', 'x'), data=e3u[None])
        w.createVariable('e3v_0', REAL, ('t', 'z', 'y', 'x'), data=e3v[None])
        w.createVariable('e2u', REAL, ('t', 'y', 'x'), data=(arc[:, 1] * radius).reshape(1, self.ny, self.nx))
        w.createVariable('e1v', REAL, ('t', 'y', 'x'), data=(arc[:, 2] * radius).reshape(1, self.ny, self.nx))
        w.close()
        return e3u, e3v


def parseDeltaDeg(text):
    vals = [float(s) for s in str(text).replace('(', ' ').replace(')', ' ').replace(',', ' ').split()]
    if len(vals) != 2:
        raise ValueError(f'deltaDeg must hold two numbers, got {text!r}')
    return tuple(vals)


def main(*, streamFunction="(cos(t*2*pi/nt)+2)*(0.5*(y/180)**2 + sin(2*pi*x/360))", prefix='', xmin=-180.,
         xmax=180., ymin=-90., ymax=90., zmin=0., zmax=1.0, nx=36, ny=18, nz=1, nt=1, deltaDeg="(0.,0.)",
         meshMask=False, partialCells=0.0):
    """Generate data (datagen.py:211-242)"""
    gen = DataGen(prefix)
    gen.setSizes(nx, ny, nz, nt)
    gen.setBoundingBox(xmin=xmin, xmax=xmax, ymin=ymin, ymax=ymax, zmin=zmin, zmax=zmax)
    gen.build()
    delta = parseDeltaDeg(deltaDeg)
    if delta[0] != 0 or delta[1] != 0:
        gen.rotatePole(deltaDeg=delta)
    gen.applyStreamFunction(streamFunction)
    gen.computeUVFromPotential()
    gen.save()
    if meshMask:
        gen.saveMeshMask(partialCells=partialCells)
    return gen


def cli(argv=None):
    ap = argparse.ArgumentParser(description='Generate data')
    ap.add_argument('-s', '--streamFunction', default="(cos(t*2*pi/nt)+2)*(0.5*(y/180)**2 + sin(2*pi*x/360))",
                    help='potential expression of x (logical lon), y (logical lat), z (depth) and t (time index)')
    ap.add_argument('-p', '--prefix', default='', help='data will be saved as <prefix>T.nc, <prefix>U.nc, <prefix>V.nc')
    for name, default, hlp in (('xmin', -180., 'min longitude'), ('xmax', 180., 'max longitude'),
                               ('ymin', -90., 'min latitude'), ('ymax', 90., 'max latitude'),
                               ('zmin', 0., 'min depth'), ('zmax', 1., 'max depth')):
        ap.add_argument('--' + name, type=float, default=default, help=hlp)
    for name, default, hlp in (('nx', 36, 'number of cells in longitude'), ('ny', 18, 'number of cells in latitude'),
                               ('nz', 1, 'number of vertical cells'), ('nt', 1, 'number of time steps')):
        ap.add_argument('--' + name, type=int, default=default, help=hlp)
    ap.add_argument('-d', '--deltaDeg', default='(0.,0.)', help='longitude, latitude pole displacement')
    ap.add_argument('--meshMask', action='store_true', help='also write <prefix>mesh_mask.nc (e3u_0, e3v_0, e2u, e1v)')
    ap.add_argument('--partialCells', type=float, default=0., help='with --meshMask: thin the deepest cells by up to this fraction')
    a = ap.parse_args(argv)
    main(**vars(a))


if __name__ == '__main__':
    cli()
