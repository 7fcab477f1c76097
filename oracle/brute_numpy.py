"""Second, independently written oracle for K1 -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Pure numpy, brute force over ALL cells for every segment (no candidate filter), written with different
formulations from oracle/nfx_oracle.c on purpose (Cramer solves vectorised over cells, the inverse
bilinear map solved in closed form through its quadratic instead of by Newton iteration).  It checks
the C oracle and the CUDA kernel at small sizes: cell/edge ids exactly, weights to ~1e-13.
Same recalled mint algorithm, so this is an implementation cross-check, not a pin
against mint itself.
"""
import numpy

EPS = 10 * numpy.finfo(numpy.float64).eps
EPS100 = 100 * EPS


def _inside(vx, vy, px, py):
    ok = numpy.ones(vx.shape[0], bool)
    for i0 in range(4):
        i1 = (i0 + 1) % 4
        cross = (px - vx[:, i0]) * (py - vy[:, i1]) - (py - vy[:, i0]) * (px - vx[:, i1])
        ok &= cross >= -EPS
    return ok


def _segment_cells(vx, vy, a, b):
    """returns (cell ids, ta, tb) of the cells met by segment a->b"""
    n = vx.shape[0]
    lam = numpy.full((n, 10), numpy.nan)
    lam[_inside(vx, vy, a[0], a[1]), 0] = 0.0
    lam[_inside(vx, vy, b[0], b[1]), 1] = 1.0
    d = b - a
    for i0 in range(4):
        i1 = (i0 + 1) % 4
        ex, ey = vx[:, i1] - vx[:, i0], vy[:, i1] - vy[:, i0]          # edge direction q1 - q0
        rx, ry = vx[:, i0] - a[0], vy[:, i0] - a[1]
        det = d[0] * (-ey) - (-ex) * d[1]
        with numpy.errstate(divide='ignore', invalid='ignore'):
            l = (rx * (-ey) - (-ex) * ry) / det
            mu = (d[0] * ry - d[1] * rx) / det
        regular = numpy.abs(det) > EPS
        hit = regular & (l >= -EPS100) & (l <= 1 + EPS100) & (mu >= -EPS100) & (mu <= 1 + EPS100)
        lam[hit, 2 + 2 * i0] = l[hit]
        # collinear overlap
        col = (~regular) & (numpy.abs(rx * (-ey) - (-ex) * ry) < EPS) & (numpy.abs(d[0] * ry - d[1] * rx) < EPS)
        if col.any():
            d2 = d.dot(d)
            lA = (rx * d[0] + ry * d[1]) / d2
            lB = ((vx[:, i1] - a[0]) * d[0] + (vy[:, i1] - a[1]) * d[1]) / d2
            lo, hi = numpy.minimum(lA, lB), numpy.maximum(lA, lB)
            keep = col & ~((lo > 1 + EPS) | (hi < -EPS))
            la, lb = numpy.maximum(lo, 0.0), numpy.minimum(hi, 1.0)
            keep &= numpy.abs(lb - la) > EPS
            lam[keep, 2 + 2 * i0] = la[keep]
            lam[keep, 3 + 2 * i0] = lb[keep]
    cnt = (~numpy.isnan(lam)).sum(axis=1)
    with numpy.errstate(all='ignore'):
        ta, tb = numpy.nanmin(lam, axis=1), numpy.nanmax(lam, axis=1)
    sel = (cnt >= 2) & (numpy.abs(tb - ta) > EPS100)
    ids = numpy.nonzero(sel)[0]
    return ids, ta[ids], tb[ids]


def _inverse_bilinear(vx, vy, px, py):
    """closed-form inverse of p = v0 + B s + C t + D s t for arrays of quads/points"""
    bx, by = vx[:, 1] - vx[:, 0], vy[:, 1] - vy[:, 0]
    cx, cy = vx[:, 3] - vx[:, 0], vy[:, 3] - vy[:, 0]
    dx = vx[:, 0] - vx[:, 1] + vx[:, 2] - vx[:, 3]
    dy = vy[:, 0] - vy[:, 1] + vy[:, 2] - vy[:, 3]
    hx, hy = px - vx[:, 0], py - vy[:, 0]
    # eliminate s: quadratic in t:  k2 t^2 + k1 t + k0 = 0
    k2 = cx * dy - cy * dx
    k1 = bx * cy - by * cx + hy * dx - hx * dy
    k0 = hx * by - hy * bx
    t = numpy.empty_like(px)
    lin = numpy.abs(k2) < 1e-14 * (numpy.abs(k1) + 1e-300)
    with numpy.errstate(all='ignore'):
        t[lin] = -k0[lin] / k1[lin]
        disc = numpy.sqrt(numpy.maximum(k1 * k1 - 4 * k2 * k0, 0.0))
        q = -0.5 * (k1 + numpy.sign(k1) * disc)
        t1, t2 = q / k2, k0 / q
        pick = numpy.where(numpy.abs(t1 - 0.5) <= numpy.abs(t2 - 0.5), t1, t2)
    t[~lin] = pick[~lin]
    den_x, den_y = bx + dx * t, by + dy * t
    with numpy.errstate(all='ignore'):
        s = numpy.where(numpy.abs(den_x) >= numpy.abs(den_y), (hx - cx * t) / den_x, (hy - cy * t) / den_y)
    # Newton polish of the closed-form root (cancellation in the quadratic costs digits on skewed cells)
    for _ in range(12):
        fx = vx[:, 0] + bx * s + cx * t + dx * s * t - px
        fy = vy[:, 0] + by * s + cy * t + dy * s * t - py
        j00, j01, j10, j11 = bx + dx * t, cx + dx * s, by + dy * t, cy + dy * s
        det = j00 * j11 - j01 * j10
        s = s - (fx * j11 - fy * j01) / det
        t = t - (fy * j00 - fx * j10) / det
    return s, t


def compute_weights(points, xyz, periodX=360., counterclock=False):
    """-> dict(cell, seg, img, ta, tb, coeff, w (n,4)) in emission order (segment, ta, cell, image)"""
    vx, vy = points[:, :, 0], points[:, :, 1]
    out = dict(cell=[], seg=[], img=[], ta=[], tb=[], coeff=[], w=[])
    imgs = (-1, 0, 1) if periodX > 0 else (0,)
    for sgi in range(len(xyz) - 1):
        p0, p1 = numpy.array(xyz[sgi][:2], float), numpy.array(xyz[sgi + 1][:2], float)
        if (p0 == p1).all():
            continue
        recs = []
        for img in imgs:
            sh = numpy.array([img * periodX, 0.0])
            ids, ta, tb = _segment_cells(vx, vy, p0 + sh, p1 + sh)
            recs += [(ta[i], int(ids[i]), img, tb[i]) for i in range(len(ids))]
        recs.sort(key=lambda r: (r[0], r[1], r[2]))
        n = len(recs)
        if n == 0:
            continue
        ta = numpy.array([r[0] for r in recs])
        tb = numpy.array([r[3] for r in recs])
        cell = numpy.array([r[1] for r in recs])
        img = numpy.array([r[2] for r in recs])
        coeff = numpy.ones(n)
        ov = numpy.maximum(0.0, numpy.minimum(tb[:-1], tb[1:]) - numpy.maximum(ta[:-1], ta[1:]))
        coeff[:-1] = 1.0 - ov / (tb[:-1] - ta[:-1])
        shx = img * periodX
        d = p1 - p0
        pax, pay = p0[0] + shx + ta * d[0], p0[1] + ta * d[1]
        pbx, pby = p0[0] + shx + tb * d[0], p0[1] + tb * d[1]
        sa, tq = _inverse_bilinear(vx[cell], vy[cell], pax, pay)
        sb, tr = _inverse_bilinear(vx[cell], vy[cell], pbx, pby)
        d0, d1 = sb - sa, tr - tq
        m0, m1 = 0.5 * (sa + sb), 0.5 * (tq + tr)
        sgn = -1.0 if counterclock else 1.0
        w = numpy.stack([d0 * (1 - m1) * coeff, d1 * m0 * coeff, sgn * d0 * m1 * coeff, sgn * d1 * (1 - m0) * coeff], 1)
        out['cell'].append(cell)
        out['seg'].append(numpy.full(n, sgi))
        out['img'].append(img)
        out['ta'].append(ta)
        out['tb'].append(tb)
        out['coeff'].append(coeff)
        out['w'].append(w)
    if not out['cell']:
        return dict(cell=numpy.zeros(0, int), seg=numpy.zeros(0, int), img=numpy.zeros(0, int), ta=numpy.zeros(0),
                    tb=numpy.zeros(0), coeff=numpy.zeros(0), w=numpy.zeros((0, 4)))
    return {k: numpy.concatenate(v) for k, v in out.items()}
