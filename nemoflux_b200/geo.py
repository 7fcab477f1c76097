"""Spherical geometry helpers -- host mirror of nemoflux/geo.py.

Arc lengths are the e2u / e1v metric factors of the edge-flux kernel (field.py:176-181).  They are
computed ONCE per grid, on the host, with the reference's own formula (arccos of a dot product,
geo.py:24-27): that formula is ill-conditioned (one ulp in the dot product moves a 1/12 degree arc by
~5e-11 relative), so reproducing the reference to 1e-12 needs the very same libm/numpy calls; the
result is uploaded and consumed by K2.
"""
import numpy

EARTH_RADIUS = 1.0          # geo.py:3 -- unit sphere; field.py:12 carries the metres
DEG2RAD = numpy.pi / 180.


def lonLat2XYZ(p, radius):
    """single point (lon, lat[, elev]) in degrees -> Cartesian (geo.py:6-12)"""
    lam, the = p[0] * DEG2RAD, p[1] * DEG2RAD
    rho = radius * numpy.cos(the)
    return numpy.array([rho * numpy.cos(lam), rho * numpy.sin(lam), radius * numpy.sin(the)])


def lonLat2XYZArray(p, radius):
    """(..., 3) lon, lat, elev -> (..., 3) x, y, z (geo.py:14-22)"""
    lam = p[..., 0] * DEG2RAD
    the = p[..., 1] * DEG2RAD
    rho = radius * numpy.cos(the)
    out = numpy.zeros(p.shape, numpy.float64)
    out[..., 0] = rho * numpy.cos(lam)
    out[..., 1] = rho * numpy.sin(lam)
    out[..., 2] = radius * numpy.sin(the)
    return out


def getArcLengthArray(xyzA, xyzB, radius=EARTH_RADIUS):
    """great-circle distance between rows of xyzA and xyzB (geo.py:24-27)"""
    angle = numpy.arccos(numpy.sum(xyzA * xyzB, axis=-1) / (radius * radius))
    return numpy.fabs(radius * angle)


def getArcLength(xyzA, xyzB, radius=EARTH_RADIUS):
    return abs(radius * numpy.arccos(xyzA.dot(xyzB) / radius**2))


def cellArcLengths(points, chunk=1 << 20):
    """(ncell, 4, 3) lon-lat points -> (ncell, 4): length of edge i -> i+1 (field.py:170-181).
    Processed in chunks of cells so that an ORCA12 grid does not need several GB of temporaries."""
    ncell = points.shape[0]
    arc = numpy.zeros((ncell, 4), numpy.float64)
    for a in range(0, ncell, chunk):
        xyz = lonLat2XYZArray(points[a:a + chunk], radius=EARTH_RADIUS)
        for i0 in range(4):
            i1 = (i0 + 1) % 4
            arc[a:a + chunk, i0] = getArcLengthArray(xyz[:, i0, :], xyz[:, i1, :], radius=EARTH_RADIUS)
    return arc
