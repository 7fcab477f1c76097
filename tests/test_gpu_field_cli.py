"""-m gpu: the reference-facing surface (Field / fluxviz / fluxplot on NetCDF files written by datagen)."""
import os

import numpy
import pytest

from conftest import GOLDEN
from helpers import README_C1, README_SINGULAR, SF_C2, tr

pytestmark = pytest.mark.gpu


def _files(tmp_path, *args):
    from nemoflux_b200 import datagen
    prefix = str(tmp_path) + '/'
    datagen.cli(list(args) + [f'--prefix={prefix}'])
    return prefix + 'T.nc', prefix + 'U.nc', prefix + 'V.nc'


def test_field_simple_example_360(gpu, tmp_path, capsys):
    """README.md:26-39 end to end: datagen --streamFunction="x", fluxviz ... -> 360"""
    from nemoflux_b200 import fluxviz
    T, U, V = _files(tmp_path, '--streamFunction=x')
    fv = fluxviz.main(['-t', T, '-u', U, '-v', V, '--lonLatPoints=(-180,-70),(-160,-10),(-35,40),(20,-50),(60,50),(180,40)'])
    f = fv.field
    out = capsys.readouterr().out
    assert 'flux =  360 (A m^2/s)' in out and 'lon-lat box: -180.0, -90.0 -> 180.0, 90.0' in out
    assert f.getFluxText() == ' 360 (A m^2/s) '
    assert (f.nt, f.nz, f.ny, f.nx) == (1, 1, 18, 36)
    g = numpy.load(os.path.join(GOLDEN, 'c1_simple.npz'))       # the reference Field's own arrays
    assert numpy.array_equal(f.arcLengths, g['arcLengths'])
    assert numpy.abs(f.integratedVelocity - g['iV']).max() <= 1e-14 * numpy.abs(g['iV']).max()
    assert numpy.allclose(f.edgeFluxesUArray, g['absEU'], rtol=1e-14) and numpy.allclose(f.edgeFluxesVArray, g['absEV'], rtol=1e-14)
    assert abs(f.maxAbsFlux - 10.0) < 1e-12
    assert abs(f.plis[0].getIntegral(f.integratedVelocity) - 360.) < 1e-10
    U0, V0 = f.getUV()
    assert numpy.allclose(U0, g['U'], rtol=1e-14) and numpy.allclose(V0, g['V'], rtol=1e-14)


def test_field_singular_and_sverdrup(gpu, tmp_path):
    from nemoflux_b200.field import Field
    T, U, V = _files(tmp_path, '--streamFunction=arctan2(y, x+180)/(2*pi)')
    f = Field(T, U, V, [tr(README_SINGULAR)], sverdrup=False, verbose=False)
    assert f.getFluxText() == ' 0.5 (A m^2/s) '
    fs = Field(T, U, V, [tr(README_SINGULAR), tr(README_C1)], sverdrup=True, verbose=False)
    assert fs.getFluxText().endswith('(Sv) ')
    assert abs(fs.fluxes[0] - 0.5 * 6.371) < 1e-14 and len(fs.plis) == 2
    assert abs(fs.maxAbsFlux - 0.125 * 6.371) < 1e-14           # pictures/singular.png colour bar 0.125


def test_fluxplot_series_and_time_stepping(gpu, oracle, tmp_path, capsys):
    from nemoflux_b200 import fluxplot, fluxviz
    T, U, V = _files(tmp_path, f'--streamFunction={SF_C2}', '--nx=72', '--ny=36', '--nz=5', '--nt=6', '--deltaDeg=20,30')
    csv = str(tmp_path / 'series.csv')
    pts = '[(-100,-70),(100,-70),(0,70)],[(-150,-20),(-20,35),(60,-40)]'
    s = fluxplot.main(['-t', T, '-u', U, '-v', V, '-l', pts, '-o', csv])
    assert s.shape == (6, 2)
    d = oracle.DataGen(nx=72, ny=36, nz=5, nt=6, deltaDeg=(20., 30.))
    u, v = d.uv(SF_C2)
    ref = oracle.flux_series(d.points(), [tr([(-100, -70), (100, -70), (0, 70)]), tr([(-150, -20), (-20, 35), (60, -40)])],
                             u, v, d.thickness())
    assert numpy.abs(s - ref).max() <= 1e-12 * numpy.abs(ref).max()
    rows = open(csv).read().strip().split('\n')
    assert rows[0] == 'time,line0,line1' and len(rows) == 7
    assert numpy.allclose([float(x) for x in rows[3].split(',')[1:]], s[2])
    # the viewer's 't' key: one update per time index gives the same numbers as the batched series
    fv = fluxviz.main(['-t', T, '-u', U, '-v', V, '-l', pts, '--allTimes'])
    assert fv.field.timeIndex == 5 and numpy.allclose(fv.field.fluxes, s[5], rtol=1e-13)
    fv.update('t')
    assert fv.field.timeIndex == 0 and numpy.allclose(fv.field.fluxes, s[0], rtol=1e-13)
    fv.update('T')
    assert fv.field.timeIndex == 5
    assert fv.field.maxAbsFlux >= numpy.abs(fv.field.edgeFluxesUArray).max()       # running maximum (field.py:234)
    out = capsys.readouterr().out
    assert out.count('flux = ') >= 6 and 'time index 5' in out


def test_field_float32_and_land(gpu, oracle, tmp_path):
    """real NEMO files: float32 uo/vo with _FillValue 1e20 over land -> zero flux there (README.md:118)"""
    from nemoflux_b200 import ncio
    from nemoflux_b200.field import Field
    T, U, V = _files(tmp_path, f'--streamFunction={SF_C2}', '--nx=36', '--ny=18', '--nz=4', '--nt=2')
    d = oracle.DataGen(nx=36, ny=18, nz=4, nt=2)
    u, v = d.uv(SF_C2)
    land = numpy.zeros((4, 18, 36), bool)
    land[:, 5:9, 10:20] = True
    land[3] = True
    for fname, vname, a in ((U, 'uo', u), (V, 'vo', v)):
        a32 = a.astype(numpy.float32)
        a32[:, land] = numpy.float32(1.e20)
        w = ncio.Writer(fname)
        for name, n in (('t', 2), ('z', 4), ('y', 18), ('x', 36)):
            w.createDimension(name, n)
        w.createVariable(vname, 'float32', ('t', 'z', 'y', 'x'), fill_value=1.e20, data=a32)
        w.close()
    path = [tr([(-150, -50), (-20, 35), (60, -40), (170, 60)])]
    f = Field(T, U, V, path, verbose=False)
    s = f.fluxSeries(chunk_steps=1)
    un = u.astype(numpy.float32).astype(numpy.float64)
    vn = v.astype(numpy.float32).astype(numpy.float64)
    un[:, land] = numpy.nan
    vn[:, land] = numpy.nan
    ref = oracle.flux_series(d.points(), path, un, vn, d.thickness())
    assert numpy.abs(s - ref).max() <= 1e-12 * numpy.abs(ref).max()
    assert numpy.isfinite(f.integratedVelocity).all() and (f.integratedVelocity.reshape(18, 36, 4)[6, 12] == 0).all()


def test_integration_md_binding_runs(gpu, oracle, tmp_path):
    """the ctypes stub INTEGRATION.md shows to nemoflux maintainers works as written: mint-style calls only"""
    import re
    import types
    from conftest import ROOT
    from nemoflux_b200 import _lib
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    code = re.search(r'```python\n(# nemoflux/mint_gpu.py.*?)```', text, re.S).group(1)
    code = code.replace("'libnemoflux_gpu.so'", repr(_lib.LIB_PATH))
    mint = types.ModuleType('mint_gpu')
    exec(compile(code, 'mint_gpu.py', 'exec'), mint.__dict__)
    d = oracle.DataGen()
    u, v = d.uv('x')
    arc = oracle.arc_lengths(d.points())
    iV, _, _ = oracle.integrated_flux(oracle.read_field(u[0], d.thickness()), oracle.read_field(v[0], d.thickness()), arc)
    g = mint.Grid()
    g.setPoints(d.points())
    assert g.getNumberOfCells() == 648
    pli = mint.PolylineIntegral()
    pli.setGrid(g)
    pli.buildLocator(numCellsPerBucket=128, periodX=360., enableFolding=False)
    pli.computeWeights(tr(README_C1), counterclock=False)
    assert abs(pli.getIntegral(iV, mint.CELL_BY_CELL_DATA) - 360.) < 1e-10
    with pytest.raises(RuntimeError):
        pli.buildLocator(enableFolding=True)


def test_vector_interp_matches_oracle_and_pictures(gpu, oracle, tmp_path):
    """mint.VectorInterp mirror (field.py:90-95): findPoints + getFaceVectors bit-exact against the oracle; the
    directions the reference's screenshots show (simple.png: south; singular.png: away from the singularity)"""
    from nemoflux_b200.field import Field
    rng = numpy.random.default_rng(8)
    for delta in ((0., 0.), (20., 30.)):
        g = oracle.DataGen(nx=60, ny=30, deltaDeg=delta)
        P = g.points()
        data = rng.standard_normal((60 * 30, 4))
        pts = numpy.zeros((400, 3))
        pts[:, 0] = rng.uniform(-200., 200., 400)        # some need the x-period, some lie outside in y
        pts[:, 1] = rng.uniform(-80., 95., 400)
        pts[:40, :2] = P[rng.integers(0, P.shape[0], 40), rng.integers(0, 4, 40), :2]     # exactly on grid nodes
        grid = gpu.Grid()
        grid.setPoints(P)
        vi = gpu.VectorInterp()
        vi.setGrid(grid)
        vi.buildLocator(numCellsPerBucket=128, periodX=360.)
        nbad = vi.findPoints(pts, tol2=1.e-12)
        ov = oracle.VectorInterp(oracle.Grid(P))
        assert ov.findPoints(pts, tol2=1.e-12) == nbad
        cell, xi = vi.getCells()
        assert numpy.array_equal(cell, ov.cell) and (cell < 0).sum() == nbad
        from helpers import assert_bitwise
        assert_bitwise(xi, ov.xi, 'parametric coordinates')
        assert_bitwise(vi.getFaceVectors(data), ov.getFaceVectors(data), 'face vectors')
    # through Field: README simple example, arrows point south with unit length
    T, U, V = _files(tmp_path, '--streamFunction=x')
    f = Field(T, U, V, [tr(README_C1)], verbose=False)
    assert f.vectorPoints.shape[0] == len(f.uVectors) > 50
    assert numpy.allclose(f.vectorValues[:, 0], 0., atol=1e-13) and numpy.allclose(f.vectorValues[:, 1], -1., rtol=1e-12)
    pts_ref, dirs_ref = oracle.transect_vector_points([tr(README_C1)], f.dx)
    assert numpy.allclose(f.vectorPoints, pts_ref) and numpy.allclose(numpy.array(f.uVectors), dirs_ref)
    T, U, V = _files(tmp_path, '--streamFunction=arctan2(y, x+180)/(2*pi)')
    f = Field(T, U, V, [tr(README_SINGULAR)], verbose=False)
    radial = f.vectorPoints[:, :2] - numpy.array([-180., 0.])
    inside = numpy.linalg.norm(radial, axis=1) > 15.
    assert ((f.vectorValues[inside, :2] * radial[inside]).sum(axis=1) > 0).all()


def test_field_with_mesh_file_scale_factors(gpu, oracle, tmp_path, capsys):
    """SURVEY 8f rank 4 through the tools: datagen --meshMask writes e3u_0/e3v_0/e2u/e1v, Field(meshFile=...) uses them
    instead of deptht_bounds and the arc lengths.  Full cells: the same fluxes as without the file (to rounding:
    arc*R/R); partial bottom cells: the oracle's series from its e3 edge fluxes; float32 files; the fluxplot flag"""
    from nemoflux_b200 import fluxplot, ncio
    from nemoflux_b200.field import EARTH_RADIUS, Field
    args = [f'--streamFunction={SF_C2}', '--nx=72', '--ny=36', '--nz=5', '--nt=6', '--deltaDeg=20,30']
    lines = [tr([(-100, -70), (100, -70), (0, 70)]), tr([(-150, -20), (-20, 35), (60, -40)])]
    T, U, V = _files(tmp_path, *args, '--meshMask')
    mesh = str(tmp_path) + '/mesh_mask.nc'
    base = Field(T, U, V, lines, verbose=False)
    s_base = base.fluxSeries(chunk_steps=4)
    full = Field(T, U, V, lines, verbose=False, meshFile=mesh)
    assert full.e3u.shape == (5, 36, 72) and numpy.allclose(full.arcLengths[:, 1:3], base.arcLengths[:, 1:3], rtol=1e-15)
    s_full = full.fluxSeries(chunk_steps=4)
    scale = numpy.abs(s_base).max()
    assert numpy.abs(s_full - s_base).max() <= 1e-13 * scale
    assert numpy.abs(full.fluxes - s_full[0]).max() <= 1e-13 * scale            # update() of time index 0
    U0, _ = full.getUV()
    Ub, _ = base.getUV()
    assert numpy.allclose(U0, Ub, rtol=1e-14, atol=1e-300)
    # partial bottom cells
    sub = tmp_path / 'partial'
    sub.mkdir()
    T, U, V = _files(sub, *args, '--meshMask', '--partialCells=0.6')
    mesh = str(sub) + '/mesh_mask.nc'
    f = Field(T, U, V, lines, sverdrup=True, verbose=False, meshFile=mesh)
    s = f.fluxSeries(chunk_steps=4)
    with ncio.open_dataset(mesh) as nc:
        e3u, e3v = nc['e3u_0'][:][0], nc['e3v_0'][:][0]
        a1, a2 = nc['e2u'][:].reshape(-1) / EARTH_RADIUS, nc['e1v'][:].reshape(-1) / EARTH_RADIUS
    assert e3u[-1].min() < 0.5 * e3u[0].max() and numpy.array_equal(e3u[0], numpy.full((36, 72), e3u[0, 0, 0]))
    d = oracle.DataGen(nx=72, ny=36, nz=5, nt=6, deltaDeg=(20., 30.))
    u, v = d.uv(SF_C2)
    og = oracle.Grid(d.points())
    ref, l1 = numpy.zeros((6, 2)), numpy.zeros((6, 2))
    for m, xyz in enumerate(lines):
        op = oracle.PolylineIntegral(og)
        op.computeWeights(xyz)
        keys, ws = op.merged_map()
        for t in range(6):
            eU, eV = oracle.edgeflux_step_c_e3(u[t], v[t], e3u, e3v, a1, a2, True)
            iV = numpy.zeros((36, 72, 4))
            iV[:, :, 1], iV[:, :, 2] = eU.reshape(36, 72), eV.reshape(36, 72)
            iV[1:, :, 0] = iV[:-1, :, 2]
            iV[:, 1:, 3] = iV[:, :-1, 1]
            iV[:, 0, 3] = iV[:, -1, 1]
            ref[t, m] = (ws * iV.reshape(-1)[keys]).sum()
            l1[t, m] = numpy.abs(ws * iV.reshape(-1)[keys]).sum()
    assert (numpy.abs(s - ref) <= 1e-12 * l1).all()
    assert numpy.abs(s - 6.371 * s_base).max() > 1e-3 * scale                  # the thinner cells do change the flux
    assert numpy.abs(f.fluxes - ref[0]).max() <= 1e-12 * l1[0].max()
    # float32 velocity files: the scale factors are taken in the storage precision of uo/vo
    for fname, vname, a in ((U, 'uo', u), (V, 'vo', v)):
        w = ncio.Writer(fname)
        for name, n in (('t', 6), ('z', 5), ('y', 36), ('x', 72)):
            w.createDimension(name, n)
        w.createVariable(vname, 'float32', ('t', 'z', 'y', 'x'), fill_value=1.e20, data=a.astype(numpy.float32))
        w.close()
    f32 = Field(T, U, V, lines, sverdrup=True, verbose=False, meshFile=mesh)
    assert f32.e3u.dtype == numpy.float32
    s32 = f32.fluxSeries(chunk_steps=4)
    assert numpy.abs(s32 - ref).max() <= 1e-5 * numpy.abs(ref).max()
    # the command-line flag
    out = fluxplot.main(['-t', T, '-u', U, '-v', V, '-s', '--meshFile', mesh,
                         '-l', '[(-100,-70),(100,-70),(0,70)],[(-150,-20),(-20,35),(60,-40)]'])
    assert numpy.array_equal(out, s32)
    with pytest.raises(RuntimeError):
        Field(T, U, V, lines, verbose=False, meshFile=T)          # no e3u_0/e3v_0 in there


def test_field_reads_netcdf4_velocity_files(gpu, oracle, tmp_path):
    """real NEMO output is NetCDF-4: uo/vo chunked per time step and level slab, shuffled and deflated, float32 with
    _FillValue over land.  Field streams such files through the package's own HDF5 reader into the same pipeline
    (here written by tests/h5build.py) and gives the series of the identical data in classic files, bit for bit"""
    import h5build
    from nemoflux_b200 import ncio
    from nemoflux_b200.field import Field
    T, U, V = _files(tmp_path, f'--streamFunction={SF_C2}', '--nx=48', '--ny=24', '--nz=6', '--nt=5', '--deltaDeg=20,30')
    d = oracle.DataGen(nx=48, ny=24, nz=6, nt=5, deltaDeg=(20., 30.))
    u, v = d.uv(SF_C2)
    u32, v32 = u.astype(numpy.float32), v.astype(numpy.float32)
    land = numpy.zeros((6, 24, 48), bool)
    land[:, 8:12, 10:20] = True
    land[4:] |= numpy.random.default_rng(3).random((24, 48)) < 0.3
    u32[:, land] = numpy.float32(1.e20)
    v32[:, land] = numpy.float32(1.e20)
    lines = [tr([(-150, -50), (-20, 35), (60, -40), (170, 60)]), tr([(-100, -70), (100, -70), (0, 70)])]
    for fname, vname, a in ((U, 'uo', u32), (V, 'vo', v32)):               # classic NetCDF-3 first
        w = ncio.Writer(fname)
        for name, n in (('t', 5), ('z', 6), ('y', 24), ('x', 48)):
            w.createDimension(name, n)
        w.createVariable(vname, 'float32', ('t', 'z', 'y', 'x'), fill_value=1.e20, data=a)
        w.close()
    s3 = Field(T, U, V, lines, verbose=False).fluxSeries(chunk_steps=2)
    U4, V4 = str(tmp_path / 'U4.nc'), str(tmp_path / 'V4.nc')
    for fname, vname, a in ((U4, 'uo', u32), (V4, 'vo', v32)):
        h5build.write(fname, {vname: dict(data=a, chunks=(1, 2, 24, 48), deflate=True, shuffle=True,
                                          attrs={'_FillValue': numpy.float32(1.e20), 'units': 'm/s'}),
                              'time_counter': dict(data=numpy.arange(5.), chunks=(1,))})
    f4 = Field(T, U4, V4, lines, verbose=False)
    s4 = f4.fluxSeries(chunk_steps=2)
    assert numpy.array_equal(s4, s3)
    # deflated chunks travel compressed and are inflated + unshuffled on the device (csrc/nfx_inflate.cu) ...
    assert f4.last_ingest['path'] == 'device decode'
    assert 0 < f4.last_ingest['bytes_compressed'] < f4.last_ingest['bytes_decoded'] == 2 * u32.nbytes
    # ... with the bits of the host decode (zlib + numpy unshuffle on reader threads), whatever the block size
    assert numpy.array_equal(f4.fluxSeries(chunk_steps=2, device_decode=False), s3)
    assert numpy.array_equal(f4.fluxSeries(chunk_steps=5, device_decode=True, prefetch=False), s3)
    assert numpy.abs(f4.fluxes - s4[0]).max() <= 1e-13 * numpy.abs(s4).max()
    un, vn = u32.astype(numpy.float64), v32.astype(numpy.float64)
    un[:, land] = numpy.nan
    vn[:, land] = numpy.nan
    ref = oracle.flux_series(d.points(), lines, un, vn, d.thickness())
    assert numpy.abs(s4 - ref).max() <= 1e-12 * numpy.abs(ref).max()
