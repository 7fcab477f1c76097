#!/usr/bin/env python
"""Pin the K1 / K3 / K4 restatement against the real mint library -- one command, the moment `import mint` works.

    python tools/pin_against_mint.py [--gpu] [--out profiles/mint_pin_report.json] [--cases c1,c2,...]

mint (`python-mint>=1.24.4`, /root/reference/README.md:12) is the third-party C++ library the reference delegates
every geometric step to (call sites /root/reference/nemoflux/field.py:44-49, 90-95, 102; horizgrid.py:23-24).  It is
neither vendored in /root/reference nor installable in the build container, so `oracle/` restates its published
algorithm and DESIGN.md section 2 lists the decisions that restatement had to take.  This script settles them:

  for every case (BASELINE config 1 and 2, the singular stream function, both closed loops, the reference's real
  grid data/sa with data/sa/S3_sa.txt, seam / node / grid-line stress transects) it drives mint exactly as
  field.py:44-49 does -- Grid.setPoints, PolylineIntegral.setGrid / buildLocator(128, 360., False) /
  computeWeights(xyz, counterclock=False) -- and recovers mint's weight map through its only public read-out,
  getIntegral(data, CELL_BY_CELL_DATA), with unit data vectors (getIntegral is linear: data = e_key returns the
  accumulated weight of that (cell, edge) key bit for bit); then diffs (cell, edge, weight) against the oracle
  (and, with --gpu, against libnemoflux_gpu.so), compares getIntegral on the case's own edge fluxes and on
  cancellation-heavy random data in both summation orders, and VectorInterp.findPoints / getFaceVectors on sample
  points of the transect (field.py:71-95).

The report answers, per case: same key set? weights bit-equal / max ulp distance? which summation order reproduces
mint's getIntegral bit for bit?  and lists the DESIGN section 2 decision each difference points at.
Exit code 0 = everything within the parity bars (keys identical, weights <= 1e-12 relative, integrals <= 1e-12 of
sum |w f|), 1 = a bar is broken, 2 = mint is not importable (nothing was compared).
"""
import argparse
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

WEIGHT_BAR = 1e-12      # relative, per merged (cell, edge) weight
INTEGRAL_BAR = 1e-12    # relative to sum |w f|

# which open decision of DESIGN.md section 2 a finding points at
DECISIONS = {
    'keys': 'acceptance window of the line-line solve (+-100 eps here, SURVEY 8c recalls +-eps), the absolute collinearity '
            'test |cross| < eps on un-normalised cross products, the drop rule |tb - ta| <= 100 eps, x-period images',
    'weights': 'inverse bilinear map (Newton from (1/2, 1/2), <= 20 steps, |dxi| < 1e-15; mint uses '
               'vtkQuad::EvaluatePosition), the duplicity coefficient 1 - overlap/(tb - ta)',
    'order': 'summation order of getIntegral: "map" = keys ascending with contributions merged in emission order '
             '(std::map) vs "list" = emission order; sub-segment tie-break (ta, cell, image)',
    'vinterp': 'findPoints: lowest containing cell id, tolerance; getFaceVectors sign / normalisation (SURVEY appendix B)',
}


def tr(pts):
    return numpy.array([(p[0], p[1], 0.0) for p in pts], numpy.float64)


def ulp_distance(a, b):
    """distance in units of the last place between two float64 arrays (same sign assumed where it matters)"""
    ia = numpy.ascontiguousarray(a, numpy.float64).view(numpy.int64)
    ib = numpy.ascontiguousarray(b, numpy.float64).view(numpy.int64)
    return numpy.abs(ia - ib)


# ---------------------------------------------------------------------------------------------------
def build_cases(O, wanted=None):
    """[dict(name, points, ny, nx, xyz, data (ncell,4), expect)]: inputs come from the oracle's restatement of
    datagen.py / field.py, which IS pinned against the reference's own code (tests/golden/make_golden.py)"""
    cases = []

    def edge_data(d, sf, t=0):
        u, v = d.uv(sf)
        arc = O.arc_lengths(d.points())
        iV, _, _ = O.integrated_flux(O.read_field(u[t], d.thickness()), O.read_field(v[t], d.thickness()), arc)
        return iV

    def add(name, d, sf, pts, expect=None, t=0):
        if wanted and name not in wanted:
            return
        cases.append(dict(name=name, points=d.points(), ny=d.ny, nx=d.nx, xyz=tr(pts), data=edge_data(d, sf, t),
                          expect=expect))

    from helpers import README_C1, README_C2, README_LOOP, README_SINGULAR, SF_C2
    add('c1', O.DataGen(), 'x', README_C1, 360.0)                                               # README.md:26-39
    add('singular', O.DataGen(), 'arctan2(y, x+180)/(2*pi)', README_SINGULAR, 0.5)                # README.md:50-58
    loop_sf = 'cos(2*pi*y/360) + sin(2*pi*x/360)'
    add('loop_rect', O.DataGen(nx=360, ny=180), loop_sf, README_LOOP, 0.0)                        # README.md:65-68
    add('loop_rot', O.DataGen(nx=360, ny=180, deltaDeg=(20., 30.)), loop_sf, README_LOOP, 0.0)    # README.md:77-79
    add('c2', O.DataGen(nx=360, ny=180, nz=10, nt=20, deltaDeg=(20., 30.)), SF_C2, README_C2, None)   # README.md:89-91
    # stress transects on the default 36 x 18 grid: the x-period seam, exact nodes, grid lines, a point 50 eps off a node
    eps = numpy.finfo(numpy.float64).eps
    add('seam', O.DataGen(), 'x', [(170., -35.), (190., 35.)])
    add('seam_neg', O.DataGen(), 'x', [(-190., 15.), (-170., -15.)])
    add('gridline', O.DataGen(), 'y', [(-180., 0.), (180., 0.)])
    add('nodes', O.DataGen(), 'x*y', [(-100., -30.), (-50., 20.), (0., 20.), (40., -60.)])
    add('near_node', O.DataGen(), 'x', [(-100. * (1 + 50 * eps), -30.), (33.3, 47.1)])
    add('inside_one_cell', O.DataGen(), 'x', [(1., 1.), (9., 9.)])
    add('zero_length', O.DataGen(), 'x', [(5., 5.), (5., 5.), (25., 15.)])
    if not wanted or 'sa' in wanted:                                                              # data/sa/T.nc + S3_sa.txt
        g = numpy.load(os.path.join(ROOT, 'tests', 'golden', 'sa_T_grid.npz'))
        lon, lat = g['bounds_lon'].astype(numpy.float64), g['bounds_lat'].astype(numpy.float64)
        pts = numpy.zeros(lon.shape[:2] + (4, 3))
        pts[..., 0], pts[..., 1] = lon, lat
        ny, nx = lon.shape[:2]
        rng = numpy.random.default_rng(5)
        cases.append(dict(name='sa', points=pts.reshape(-1, 4, 3), ny=ny, nx=nx,
                          xyz=tr([(16., -40.4), (28., -34.5), (31., -28.5), (36., -30.5)]),      # data/sa/S3_sa.txt:5-8
                          data=rng.normal(size=(ny * nx, 4)), expect=None))
    return cases


# ---------------------------------------------------------------------------------------------------
def mint_map(mint, points, xyz, keys):
    """mint's accumulated weight of every key in `keys` (cell*4 + edge), recovered with unit data vectors"""
    grid = mint.Grid()
    grid.setPoints(points)                                               # horizgrid.py:23-24
    pli = mint.PolylineIntegral()
    pli.setGrid(grid)
    pli.buildLocator(numCellsPerBucket=128, periodX=360., enableFolding=False)      # field.py:47
    pli.computeWeights(xyz, counterclock=False)                                      # field.py:48
    ncell = points.shape[0]
    data = numpy.zeros((ncell, 4), numpy.float64)
    flat = data.reshape(-1)
    out = numpy.zeros(len(keys), numpy.float64)
    for n, k in enumerate(keys):
        flat[k] = 1.0
        out[n] = pli.getIntegral(data, mint.CELL_BY_CELL_DATA)          # field.py:102
        flat[k] = 0.0
    return pli, grid, out


def candidate_keys(keys_oracle, ncell, ny, nx, budget):
    """keys to probe: every key when the grid is small, else the oracle's keys plus all edges of those cells and of
    their 8 structured neighbours (a key mint holds and the oracle does not must sit next to the path)"""
    if ncell * 4 <= budget:
        return numpy.arange(ncell * 4, dtype=numpy.int64)
    cells = numpy.unique(keys_oracle // 4)
    j, i = cells // nx, cells % nx
    nb = set()
    for dj in (-1, 0, 1):
        for di in (-1, 0, 1):
            jj, ii = j + dj, (i + di) % nx
            ok = (jj >= 0) & (jj < ny)
            nb.update((jj[ok] * nx + ii[ok]).tolist())
    cells = numpy.array(sorted(nb), numpy.int64)
    return (cells[:, None] * 4 + numpy.arange(4)[None]).reshape(-1)


def compare_case(mint, O, case, gpu=None, budget=60000, rng=None):
    rng = rng or numpy.random.default_rng(1)
    pts, xyz = case['points'], case['xyz']
    ncell = pts.shape[0]
    og = O.Grid(pts)
    op = O.PolylineIntegral(og, 360.)
    op.computeWeights(xyz)
    okeys, ow = op.merged_map()
    probe = candidate_keys(okeys, ncell, case['ny'], case['nx'], budget)
    pli, grid, mw = mint_map(mint, pts, xyz, probe)
    mkeys = probe[mw != 0.0]
    mvals = mw[mw != 0.0]
    rep = dict(name=case['name'], ncell=int(ncell), probed_keys=int(probe.size), keys_mint=int(mkeys.size),
               keys_oracle=int(okeys.size))
    only_m = numpy.setdiff1d(mkeys, okeys)
    only_o = numpy.setdiff1d(okeys, mkeys)
    # a key whose oracle weight is exactly 0 (duplicate sub-segment with coefficient 0) is not a difference
    only_o = numpy.array([k for k in only_o if ow[numpy.searchsorted(okeys, k)] != 0.0], numpy.int64)
    rep['keys_only_mint'] = [int(k) for k in only_m[:20]]
    rep['keys_only_oracle'] = [int(k) for k in only_o[:20]]
    rep['keys_identical'] = bool(only_m.size == 0 and only_o.size == 0)
    common = numpy.intersect1d(mkeys, okeys)
    a = mvals[numpy.searchsorted(mkeys, common)]
    b = ow[numpy.searchsorted(okeys, common)]
    if common.size:
        rel = numpy.abs(a - b) / numpy.maximum(numpy.abs(a), 1e-300)
        ulp = ulp_distance(a, b)
        rep.update(weights_bit_equal=int((ulp == 0).sum()), weights_compared=int(common.size),
                   weights_max_rel=float(rel.max()), weights_max_ulp=int(ulp.max()))
    else:
        rep.update(weights_bit_equal=0, weights_compared=0, weights_max_rel=0.0, weights_max_ulp=0)
    # were all of mint's keys probed?  random data: sum over the probed map must reproduce mint's integral
    rnd = rng.normal(size=(ncell, 4))
    f_m = pli.getIntegral(rnd, mint.CELL_BY_CELL_DATA)
    f_probe = float(numpy.sum(mvals * rnd.reshape(-1)[mkeys]))
    sc = float(numpy.abs(mvals * rnd.reshape(-1)[mkeys]).sum()) or 1.0
    rep['probe_covers_mint_map'] = bool(abs(f_m - f_probe) <= 1e-11 * sc)
    # the case's own edge fluxes
    data = numpy.ascontiguousarray(case['data'], numpy.float64)
    f_mint = pli.getIntegral(data, mint.CELL_BY_CELL_DATA)
    scale = float(numpy.abs(ow * data.reshape(-1)[okeys]).sum()) or 1.0
    rep['integral_mint'] = float(f_mint)
    for order in ('map', 'list'):
        f_o = op.getIntegral(data, order)
        rep[f'integral_oracle_{order}'] = float(f_o)
        rep[f'integral_rel_err_{order}'] = float(abs(f_o - f_mint) / scale)
        rep[f'integral_bit_equal_{order}'] = bool(f_o == f_mint)
    if case.get('expect') is not None:
        rep['expect'] = case['expect']
        rep['mint_vs_expect'] = float(abs(f_mint - case['expect']))
    # summation order: cancellation-heavy data make the two orders differ in the last bits
    hits = dict(map=0, list=0)
    trials = 16
    for _ in range(trials):
        big = rng.normal(size=(ncell, 4)) * 10.0 ** rng.integers(-3, 9, size=(ncell, 4))
        fm = pli.getIntegral(big, mint.CELL_BY_CELL_DATA)
        for order in ('map', 'list'):
            hits[order] += int(op.getIntegral(big, order) == fm)
    rep['order_trials'] = trials
    rep['order_bit_equal_hits'] = hits
    # VectorInterp at sample points of the transect (field.py:71-95)
    try:
        keep = [0] + [i for i in range(1, len(xyz)) if not numpy.array_equal(xyz[i], xyz[i - 1])]
        vp, _ = O.transect_vector_points([xyz[keep]], 5.0)      # field.py:72-87 (zero-length segments dropped)
        vi = mint.VectorInterp()
        vi.setGrid(grid)
        vi.buildLocator(numCellsPerBucket=128, periodX=360., enableFolding=False)      # field.py:91-92
        vi.findPoints(vp, tol2=1.e-12)                                                 # field.py:93
        vm = numpy.asarray(vi.getFaceVectors(data, placement=mint.CELL_BY_CELL_DATA))  # field.py:95,120
        ovi = O.VectorInterp(og, 360.)
        ovi.findPoints(vp, tol2=1.e-12)
        vo = numpy.asarray(ovi.getFaceVectors(data))
        vs = float(numpy.abs(vm).max()) or 1.0
        rep['vinterp_points'] = int(vp.shape[0])
        rep['vinterp_max_rel'] = float(numpy.abs(vm[:, :2] - vo[:, :2]).max() / vs)
    except Exception as e:    # an older / newer mint may name things differently: report, do not hide
        rep['vinterp_error'] = repr(e)
    if gpu is not None:
        g = gpu.Grid()
        g.setPoints(pts)
        p = gpu.PolylineIntegral()
        p.build(g, periodX=360.)
        p.computeWeights(xyz)
        mp = p.getMap()
        gk, gw = mp['keys'], mp['w']
        nz = gw != 0.0
        rep['gpu_keys_identical_to_mint'] = bool(numpy.array_equal(numpy.sort(gk[nz]), numpy.sort(mkeys)))
        cg = numpy.intersect1d(gk, mkeys)
        if cg.size:
            ga = gw[numpy.searchsorted(gk, cg)]
            ma = mvals[numpy.searchsorted(mkeys, cg)]
            rep['gpu_weights_max_ulp_vs_mint'] = int(ulp_distance(ga, ma).max())
        rep['gpu_integral_rel_err'] = float(abs(p.getIntegral(data) - f_mint) / scale)
    rep['ok'] = bool(rep['keys_identical'] and rep['probe_covers_mint_map'] and rep['weights_max_rel'] <= WEIGHT_BAR and
                     rep['integral_rel_err_map'] <= INTEGRAL_BAR and rep['integral_rel_err_list'] <= INTEGRAL_BAR)
    rep['decisions_in_question'] = [DECISIONS[k] for k, bad in (
        ('keys', not rep['keys_identical']), ('weights', rep['weights_max_ulp'] > 0),
        ('order', hits['map'] != trials and hits['list'] != trials), ('vinterp', rep.get('vinterp_max_rel', 0.0) > 1e-12)) if bad]
    return rep


def run(mint, O, gpu=None, wanted=None, budget=60000):
    reps = [compare_case(mint, O, c, gpu, budget) for c in build_cases(O, wanted)]
    hits = {o: sum(r['order_bit_equal_hits'][o] for r in reps) for o in ('map', 'list')}
    return dict(mint_version=getattr(mint, '__version__', 'unknown'), cases=reps, all_ok=all(r['ok'] for r in reps),
                summation_order_votes=hits,
                summation_order_verdict=('map' if hits['map'] > hits['list'] else 'list' if hits['list'] > hits['map'] else 'undecided'),
                bars=dict(weight_rel=WEIGHT_BAR, integral_rel=INTEGRAL_BAR))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpu', action='store_true', help='also diff libnemoflux_gpu.so (needs a CUDA device)')
    ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'mint_pin_report.json'))
    ap.add_argument('--cases', default='')
    ap.add_argument('--budget', type=int, default=60000, help='max unit-vector probes per case')
    a = ap.parse_args()
    try:
        import mint
    except ImportError as e:
        print(f'mint is not importable ({e}): nothing compared.  Install python-mint>=1.24.4 (README.md:12) and rerun.')
        return 2
    from oracle import oracle as O
    O.lib()
    gpu = None
    if a.gpu:
        from nemoflux_b200 import nemoflux_gpu as gpu
    rep = run(mint, O, gpu, set(a.cases.split(',')) if a.cases else None, a.budget)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rep, open(a.out, 'w'), indent=1)
    for r in rep['cases']:
        print(f"{r['name']:16s} keys {'same' if r['keys_identical'] else 'DIFFER'} ({r['keys_mint']} / {r['keys_oracle']})  "
              f"weights bit-equal {r['weights_bit_equal']}/{r['weights_compared']} max ulp {r['weights_max_ulp']}  "
              f"integral rel err map {r['integral_rel_err_map']:.2e} list {r['integral_rel_err_list']:.2e}  "
              f"{'ok' if r['ok'] else 'BAR BROKEN'}")
    print('summation order reproducing mint bit for bit:', rep['summation_order_votes'], '->', rep['summation_order_verdict'])
    print('report:', a.out)
    return 0 if rep['all_ok'] else 1


if __name__ == '__main__':
    sys.exit(main())
