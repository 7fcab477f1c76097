"""shared helpers of the test-suite (transect builders, bitwise comparison)"""
import numpy


def tr(pts):
    """[(lon, lat), ...] -> (n, 3) float64 with z = 0 (what fluxviz.py:378 builds)"""
    return numpy.array([(p[0], p[1], 0.0) for p in pts], numpy.float64)


README_C1 = [(-180, -70), (-160, -10), (-35, 40), (20, -50), (60, 50), (180, 40)]      # README.md:32
README_SINGULAR = [(-180, -80), (-10, -80), (-10, 80), (-180, 80)]                     # README.md:51
README_LOOP = [(-100, -80), (100, -80), (0, 80), (-100, -80)]                          # README.md:66
README_C2 = [(-100, -80), (100, -80), (0, 80)]                                         # README.md:91
SF_C2 = '(1+10*z)*(t+1)*(cos(2*pi*y/360) + sin(2*pi*x/360))'                           # README.md:89


def bits(a):
    return numpy.ascontiguousarray(a, numpy.float64).view(numpy.int64)


def assert_bitwise(a, b, what=''):
    a = numpy.ascontiguousarray(a, numpy.float64)
    b = numpy.ascontiguousarray(b, numpy.float64)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    bad = numpy.nonzero(bits(a).reshape(-1) != bits(b).reshape(-1))[0]
    # -0.0 and +0.0 are the same number; anything else must match bit for bit
    bad = [i for i in bad if not (a.reshape(-1)[i] == 0.0 and b.reshape(-1)[i] == 0.0)]
    assert not bad, f'{what}: {len(bad)} values differ, first at {bad[0]}: {a.reshape(-1)[bad[0]]!r} vs {b.reshape(-1)[bad[0]]!r}'


def random_transects(rng, n, lon_range=(-170., 170.), lat_range=(-75., 75.), max_pts=6, nodes=None):
    """n random polylines; a quarter closed loops; with `nodes` (ny+1, nx+1, 2) a quarter snapped to nodes"""
    out = []
    for m in range(n):
        k = int(rng.integers(2, max_pts + 1))
        lon = rng.uniform(lon_range[0], lon_range[1], k)
        lat = rng.uniform(lat_range[0], lat_range[1], k)
        pts = numpy.stack([lon, lat], 1)
        if nodes is not None and m % 4 == 1:
            jj = rng.integers(1, nodes.shape[0] - 1, k)
            ii = rng.integers(1, nodes.shape[1] - 1, k)
            pts = nodes[jj, ii]
        if m % 4 == 2 and k >= 3:
            pts = numpy.concatenate([pts, pts[:1]], 0)
        out.append(tr(pts))
    return out
