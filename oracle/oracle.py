"""CPU ORACLE -- test infrastructure, NOT product code.

Python side of the oracle: numpy restatements of the reference's numpy code (datagen.py, geo.py,
field.py) and ctypes bindings to oracle/libnfx_oracle.so (the C restatement of the mint pieces and a
sequential/threaded K2).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; nemoflux_b200/ never does.

Parity status (also in oracle/nfx_oracle.c and DESIGN.md): the K2 part is pinned by tests/golden/*
(outputs of the reference's own code); the mint part (K1/K3) is a restatement from memory of a library
that is neither vendored in /root/reference nor installed here -> "parity unpinned" against mint itself,
anchored on README.md:39,56,68 and the invariants of SURVEY.md section 8c.
"""
import ctypes
import os
import subprocess

import numpy
from numpy import pi, cos, sin, arctan2, arctan  # noqa: F401  (names visible to stream functions)
import math  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SUBSEG_DTYPE = numpy.dtype([('cell', '<i8'), ('seg', '<i4'), ('img', '<i4'), ('ta', '<f8'), ('tb', '<f8'),
                            ('coeff', '<f8'), ('xia', '<f8', (2,)), ('xib', '<f8', (2,)), ('w', '<f8', (4,))])
assert SUBSEG_DTYPE.itemsize == 104


def build():
    """compile oracle/libnfx_oracle.so (gcc; see oracle/Makefile)"""
    subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libnfx_oracle.so')
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        L.orc_grid_new.restype = ctypes.c_void_p
        L.orc_grid_new.argtypes = [dp, ctypes.c_int64]
        L.orc_grid_del.argtypes = [ctypes.c_void_p]
        L.orc_pli_compute_weights.argtypes = [ctypes.c_void_p, dp, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                              ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                              ctypes.POINTER(ctypes.c_int64)]
        L.orc_free.argtypes = [ctypes.c_void_p]
        L.orc_pli_merge.restype = ctypes.c_int64
        L.orc_pli_merge.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), dp]
        L.orc_pli_get_integral.restype = ctypes.c_double
        L.orc_pli_get_integral.argtypes = [ctypes.c_void_p, ctypes.c_int64, dp, ctypes.c_int]
        L.orc_edgeflux_step.argtypes = [dp, dp, dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_double, dp, dp, dp]
        fp = ctypes.POINTER(ctypes.c_float)
        L.orc_edgeflux_step_f32.argtypes = [fp, fp, dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_float, dp, dp, dp]
        L.orc_edgeflux_step_e3.argtypes = [dp, dp, dp, dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_double, dp, dp]
        L.orc_vinterp_find_points.argtypes = [ctypes.c_void_p, dp, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                              ctypes.POINTER(ctypes.c_int64), dp]
        L.orc_vinterp_face_vectors.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), dp, ctypes.c_int64, dp, dp]
        L.orc_flux_index.restype = ctypes.c_int64
        L.orc_flux_index.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.orc_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ------------------------------------------------------------------------------------------------
# datagen.py restated (datagen.py:35-166).  Vectorised; the reference loops vertex by vertex with
# math.asin/math.atan2, which differ from numpy's vector functions in the last ulp, so golden
# comparisons use a tolerance and skip the pole points (whose longitude is arbitrary).
# ------------------------------------------------------------------------------------------------
class DataGen(object):

    def __init__(self, nx=36, ny=18, nz=1, nt=1, xmin=-180., xmax=180., ymin=-90., ymax=90., zmin=0., zmax=1.,
                 deltaDeg=(0., 0.), dy=None):
        self.nx, self.ny, self.nz, self.nt = nx, ny, nz, nt
        self.xmin, self.xmax, self.ymin, self.ymax, self.zmin, self.zmax = xmin, xmax, ymin, ymax, zmin, zmax
        # vertical axis, datagen.py:35-40 (note ztop=(k+1)dz, zbot=(k+2)dz: thickness is still dz)
        dz = (zmax - zmin) / float(nz)
        self.zhalf = numpy.array([zmin + (k + 0.5) * dz for k in range(nz)])
        self.ztop = numpy.array([zmin + (k + 1) * dz for k in range(nz)])
        self.zbot = numpy.array([zmin + (k + 2) * dz for k in range(nz)])
        # horizontal axes, datagen.py:42-66.  The reference uses dx for the y spacing too
        # (datagen.py:49); pass dy explicitly for grids where that is not wanted (ORCA shapes).
        dx = (xmax - xmin) / float(nx)
        if dy is None:
            dy = dx
        x = numpy.array([xmin + i * dx for i in range(nx + 1)])
        y = numpy.array([ymin + j * dy for j in range(ny + 1)])
        self.xx, self.yy = numpy.meshgrid(x, y, indexing='xy')
        self.bounds_lon = self._corners(self.xx)
        self.bounds_lat = self._corners(self.yy)
        if deltaDeg[0] != 0 or deltaDeg[1] != 0:
            self._rotate_pole(deltaDeg)

    @staticmethod
    def _corners(a):
        out = numpy.zeros((a.shape[0] - 1, a.shape[1] - 1, 4), numpy.float64)
        out[..., 0] = a[:-1, :-1]
        out[..., 1] = a[:-1, 1:]
        out[..., 2] = a[1:, 1:]
        out[..., 3] = a[1:, :-1]
        return out

    def _rotate_pole(self, deltaDeg):
        # datagen.py:116-166
        alpha = numpy.pi * deltaDeg[1] / 180.
        beta = numpy.pi * deltaDeg[0] / 180.
        ca, sa, cb, sb = numpy.cos(alpha), numpy.sin(alpha), numpy.cos(beta), numpy.sin(beta)
        rot_alp = numpy.array([[ca, 0., sa], [0., 1., 0.], [-sa, 0., ca]])
        rot_bet = numpy.array([[cb, sb, 0.], [-sb, cb, 0.], [0., 0., 1.]])
        m = numpy.dot(rot_bet, rot_alp)
        the = numpy.pi * self.bounds_lat / 180.
        lam = numpy.pi * self.bounds_lon / 180.
        rho = numpy.cos(the)
        xo, yo, zo = rho * numpy.cos(lam), rho * numpy.sin(lam), numpy.sin(the)
        xn = (m[0, 0] * xo + m[0, 1] * yo) + m[0, 2] * zo
        yn = (m[1, 0] * xo + m[1, 1] * yo) + m[1, 2] * zo
        zn = (m[2, 0] * xo + m[2, 1] * yo) + m[2, 2] * zo
        self.bounds_lat = 180. * numpy.arcsin(numpy.clip(zn, -1., 1.)) / numpy.pi
        lon = 180. * numpy.arctan2(yn, xn) / numpy.pi
        dlon = lon - lon[..., 0:1]     # date line fix relative to vertex 0, datagen.py:161-166
        lon = numpy.where(dlon > +270., lon - 360., lon)
        lon = numpy.where(dlon < -270., lon + 360., lon)
        self.bounds_lon = lon

    def potential_nodes(self, streamFunction, t, k):
        """stream function on the (ny+1, nx+1) UN-rotated logical nodes, datagen.py:69-78"""
        zmin, zmax = self.zmin, self.zmax  # noqa: F841
        A = 1.0                            # noqa: F841  geo.EARTH_RADIUS, geo.py:3
        nt = self.nt                       # noqa: F841
        z = self.zhalf[k]                  # noqa: F841
        x, y = self.xx, self.yy            # noqa: F841
        pot = eval(streamFunction)
        return pot + numpy.zeros_like(self.xx)

    def metric(self):
        """ds21, ds23 of datagen.py:89-104 (on the UN-rotated coordinates)"""
        def xyz(lon, lat):
            p = numpy.zeros(lon.shape + (3,))
            p[..., 0], p[..., 1] = lon, lat
            return lonLat2XYZArray(p, 1.0)
        xyz1 = xyz(self.xx[:-1, 1:], self.yy[:-1, 1:])
        xyz2 = xyz(self.xx[1:, 1:], self.yy[1:, 1:])
        xyz3 = xyz(self.xx[1:, :-1], self.yy[1:, :-1])
        ds21 = getArcLengthArray(xyz2, xyz1, 1.0)
        ds23 = getArcLengthArray(xyz2, xyz3, 1.0)
        ds23 = numpy.clip(ds23, 1.e-12, None)
        return ds21, ds23

    def uv(self, streamFunction):
        """(nt,nz,ny,nx) u and v, datagen.py:85-113"""
        ds21, ds23 = self.metric()
        u = numpy.zeros((self.nt, self.nz, self.ny, self.nx), numpy.float64)
        v = numpy.zeros((self.nt, self.nz, self.ny, self.nx), numpy.float64)
        for t in range(self.nt):
            for k in range(self.nz):
                pot = self.potential_nodes(streamFunction, t, k)
                p1, p2, p3 = pot[:-1, 1:], pot[1:, 1:], pot[1:, :-1]
                u[t, k] = (p2 - p1) / ds21
                v[t, k] = -(p2 - p3) / ds23
        return u, v

    def points(self):
        """(ncell,4,3) vertex array handed to mint, horizgrid.py:17-22"""
        p = numpy.zeros((self.ny, self.nx, 4, 3), numpy.float64)
        p[..., 0] = self.bounds_lon
        p[..., 1] = self.bounds_lat
        return p.reshape((self.ny * self.nx, 4, 3))

    def thickness(self):
        return self.zbot - self.ztop     # field.py:51 on datagen.py:176-178


# ------------------------------------------------------------------------------------------------
# geo.py:6-27 and field.py:170-181
# ------------------------------------------------------------------------------------------------
DEG2RAD = numpy.pi / 180.


def lonLat2XYZArray(p, radius):
    lam = p[..., 0] * DEG2RAD
    the = p[..., 1] * DEG2RAD
    rho = radius * numpy.cos(the)
    xyz = numpy.zeros(p.shape, numpy.float64)
    xyz[..., 0] = rho * numpy.cos(lam)
    xyz[..., 1] = rho * numpy.sin(lam)
    xyz[..., 2] = radius * numpy.sin(the)
    return xyz


def getArcLengthArray(xyzA, xyzB, radius):
    angle = numpy.arccos(numpy.sum(xyzA * xyzB, axis=-1) / (radius * radius))
    return numpy.fabs(radius * angle)


def arc_lengths(points):
    """(ncell,4) unit-sphere arc length of edge i0 -> i0+1, field.py:170-181"""
    xyz = lonLat2XYZArray(points, 1.0)
    arc = numpy.zeros((points.shape[0], 4), numpy.float64)
    for i0 in range(4):
        i1 = (i0 + 1) % 4
        arc[:, i0] = getArcLengthArray(xyz[:, i0, :], xyz[:, i1, :], 1.0)
    return arc


# ------------------------------------------------------------------------------------------------
# field.py:145-163 (readField) and field.py:183-234 (computeIntegratedFlux), numpy as the reference
# ------------------------------------------------------------------------------------------------
EARTH_RADIUS = 6371000.0     # field.py:12


def read_field(field_zyx, thickness):
    f = numpy.where(numpy.isnan(field_zyx), 0.0, field_zyx)           # fillna(0.0)
    return numpy.tensordot(thickness, f, axes=(0, 0))


def integrated_flux(U, V, arc, sverdrup=False):
    """returns iV (ncell,4), signed eU, eV (ncell,) exactly as field.py:195-228 computes them"""
    ny, nx = U.shape
    ncell = ny * nx
    eUa = + U.reshape((ncell,)) * arc[:, 1]
    eVa = - V.reshape((ncell,)) * arc[:, 2]
    iVa = numpy.zeros((ncell, 4), numpy.float64)
    iVa[:, 1] = eUa
    iVa[:, 2] = eVa
    eU, eV, iV = eUa.reshape((ny, nx)), eVa.reshape((ny, nx)), iVa.reshape((ny, nx, 4))
    iV[1:, :, 0] = eV[:-1, :]
    iV[:, 1:, 3] = eU[:, :-1]
    iV[:, 0, 3] = eU[:, -1]
    if sverdrup:
        eU *= EARTH_RADIUS / 1.e6
        eV *= EARTH_RADIUS / 1.e6
        iV *= EARTH_RADIUS / 1.e6
    return iVa, eUa, eVa


def edgeflux_step_c(u_zyx, v_zyx, thickness, arc1, arc2, sverdrup=False, fill=float('nan'), want_iV=True):
    """sequential-k C restatement (bit-exact reference for the CUDA kernel); OpenMP over cells"""
    nz, ny, nx = u_zyx.shape
    ncell = ny * nx
    eU = numpy.empty(ncell)
    eV = numpy.empty(ncell)
    iV = numpy.empty((ncell, 4)) if want_iV else None
    th = numpy.ascontiguousarray(thickness, numpy.float64)
    a1 = numpy.ascontiguousarray(arc1, numpy.float64)
    a2 = numpy.ascontiguousarray(arc2, numpy.float64)
    L = lib()
    if u_zyx.dtype == numpy.float32:
        fp = ctypes.POINTER(ctypes.c_float)
        u = numpy.ascontiguousarray(u_zyx)
        v = numpy.ascontiguousarray(v_zyx)
        L.orc_edgeflux_step_f32(u.ctypes.data_as(fp), v.ctypes.data_as(fp), _dp(th), _dp(a1), _dp(a2), nz, ny, nx,
                                int(sverdrup), fill, _dp(eU), _dp(eV), _dp(iV) if want_iV else None)
    else:
        u = numpy.ascontiguousarray(u_zyx, numpy.float64)
        v = numpy.ascontiguousarray(v_zyx, numpy.float64)
        L.orc_edgeflux_step(_dp(u), _dp(v), _dp(th), _dp(a1), _dp(a2), nz, ny, nx, int(sverdrup), fill, _dp(eU),
                            _dp(eV), _dp(iV) if want_iV else None)
    return iV, eU, eV


def edgeflux_step_c_e3(u_zyx, v_zyx, e3u_zyx, e3v_zyx, arc1, arc2, sverdrup=False, fill=float('nan')):
    """per-column scale factors: U = sum_k e3u[k,c]*u[k,c], sequential k, fp64"""
    nz, ny, nx = u_zyx.shape
    ncell = ny * nx
    eU, eV = numpy.empty(ncell), numpy.empty(ncell)
    arrs = [numpy.ascontiguousarray(a, numpy.float64) for a in (u_zyx, v_zyx, e3u_zyx, e3v_zyx, arc1, arc2)]
    lib().orc_edgeflux_step_e3(*[_dp(a) for a in arrs], nz, ny, nx, int(sverdrup), fill, _dp(eU), _dp(eV))
    return eU, eV


# ------------------------------------------------------------------------------------------------
# mint.Grid / mint.PolylineIntegral restated (C), call sites horizgrid.py:23-24, field.py:44-49,102
# ------------------------------------------------------------------------------------------------
class Grid(object):
    def __init__(self, points):
        self.points = numpy.ascontiguousarray(points, numpy.float64)
        assert self.points.ndim == 3 and self.points.shape[1:] == (4, 3)
        self.h = lib().orc_grid_new(_dp(self.points), self.points.shape[0])

    def __del__(self):
        if getattr(self, 'h', None):
            try:
                lib().orc_grid_del(self.h)
            except TypeError:       # interpreter shutdown: the module globals are already gone
                pass
            self.h = None

    def getNumberOfCells(self):
        return self.points.shape[0]


class PolylineIntegral(object):
    """one transect; computeWeights returns nothing, results in .subsegs (SUBSEG_DTYPE)"""

    def __init__(self, grid, periodX=360., filter_mode=1):
        self.grid = grid
        self.periodX = periodX
        self.filter_mode = filter_mode
        self.subsegs = numpy.zeros(0, SUBSEG_DTYPE)

    def computeWeights(self, xyz, counterclock=False):
        xyz = numpy.ascontiguousarray(xyz, numpy.float64)
        if xyz.ndim != 2 or xyz.shape[1] != 3:
            raise ValueError('xyz must be (npoints, 3)')
        out = ctypes.c_void_p()
        n = ctypes.c_int64()
        ier = lib().orc_pli_compute_weights(self.grid.h, _dp(xyz), xyz.shape[0], float(self.periodX),
                                            int(counterclock), int(self.filter_mode), ctypes.byref(out),
                                            ctypes.byref(n))
        if ier != 0:
            raise RuntimeError(f'orc_pli_compute_weights failed ({ier})')
        if n.value > 0:
            buf = (ctypes.c_char * (n.value * SUBSEG_DTYPE.itemsize)).from_address(out.value)
            self.subsegs = numpy.frombuffer(buf, SUBSEG_DTYPE).copy()
        else:
            self.subsegs = numpy.zeros(0, SUBSEG_DTYPE)
        if out.value:
            lib().orc_free(out)
        return 0

    def emission_list(self):
        """(cellId, edgeIndex, weight), 4 entries per sub-segment, in emission order"""
        n = self.subsegs.shape[0]
        cells = numpy.repeat(self.subsegs['cell'], 4)
        edges = numpy.tile(numpy.arange(4, dtype=numpy.int32), n)
        return cells, edges, self.subsegs['w'].reshape(-1).copy()

    def merged_map(self):
        """mint's std::map view: unique (cell*4+edge) keys ascending and the accumulated weights"""
        n = self.subsegs.shape[0]
        keys = numpy.zeros(4 * n + 1, numpy.int64)
        ws = numpy.zeros(4 * n + 1, numpy.float64)
        s = numpy.ascontiguousarray(self.subsegs)
        m = lib().orc_pli_merge(s.ctypes.data_as(ctypes.c_void_p), n,
                                keys.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _dp(ws))
        return keys[:m].copy(), ws[:m].copy()

    def getIntegral(self, data, order='map'):
        data = numpy.ascontiguousarray(data, numpy.float64).reshape(-1)
        s = numpy.ascontiguousarray(self.subsegs)
        return lib().orc_pli_get_integral(s.ctypes.data_as(ctypes.c_void_p), s.shape[0], _dp(data),
                                          1 if order == 'map' else 0)

    def segment_totals(self, nseg):
        """sum of (tb-ta)*coeff per segment: 1 when the segment lies inside the grid"""
        tot = numpy.zeros(nseg)
        numpy.add.at(tot, self.subsegs['seg'], (self.subsegs['tb'] - self.subsegs['ta']) * self.subsegs['coeff'])
        return tot


class VectorInterp(object):
    """mint.VectorInterp restated (field.py:90-95): findPoints + getFaceVectors"""

    def __init__(self, grid, periodX=360.):
        self.grid, self.periodX = grid, periodX
        self.cell = numpy.zeros(0, numpy.int64)
        self.xi = numpy.zeros((0, 2))

    def findPoints(self, points, tol2=1.e-12):
        pts = numpy.ascontiguousarray(points, numpy.float64).reshape(-1, 3)
        n = pts.shape[0]
        self.cell = numpy.zeros(n, numpy.int64)
        self.xi = numpy.zeros((n, 2))
        tol = max(float(tol2), 10 * numpy.finfo(numpy.float64).eps)
        lib().orc_vinterp_find_points(self.grid.h, _dp(pts), n, float(self.periodX), tol,
                                      self.cell.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _dp(self.xi))
        return int((self.cell < 0).sum())

    def getFaceVectors(self, data, placement=0):
        d = numpy.ascontiguousarray(data, numpy.float64).reshape(-1)
        vec = numpy.zeros((self.cell.shape[0], 3))
        lib().orc_vinterp_face_vectors(self.grid.h, self.cell.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                       _dp(self.xi), self.cell.shape[0], _dp(d), _dp(vec))
        return vec


def transect_vector_points(lonLatZPoints, dx):
    """sample points along the transects and their unit direction vectors, field.py:72-87"""
    pts, dirs = [], []
    for line in lonLatZPoints:
        line = numpy.asarray(line, numpy.float64)
        for i in range(len(line) - 1):
            beg, end = line[i], line[i + 1]
            u = end - beg
            dist = numpy.sqrt(u.dot(u))
            u = u / dist
            nv = max(2, int(dist / dx))
            vdx = dist / float(nv - 1)
            for j in range(nv):
                pts.append(beg + u * j * vdx)
                dirs.append(u)
    return numpy.array(pts), numpy.array(dirs)


def flux_series(points, transects, u, v, thickness, sverdrup=False, periodX=360., order='map', use_c=False):
    """the fluxplot.py:48-59 loop restated: (nt, M) series.  u, v are (nt,nz,ny,nx)."""
    nt, nz, ny, nx = u.shape
    grid = Grid(points)
    arc = arc_lengths(points)
    plis = []
    for xyz in transects:
        p = PolylineIntegral(grid, periodX)
        p.computeWeights(numpy.asarray(xyz, numpy.float64))
        plis.append(p)
    out = numpy.zeros((nt, len(plis)))
    for t in range(nt):
        if use_c:
            iV, _, _ = edgeflux_step_c(u[t], v[t], thickness, arc[:, 1], arc[:, 2], sverdrup)
        else:
            U = read_field(u[t], thickness)
            V = read_field(v[t], thickness)
            iV, _, _ = integrated_flux(U, V, arc, sverdrup)
        for m, p in enumerate(plis):
            out[t, m] = p.getIntegral(iV, order)
    return out
