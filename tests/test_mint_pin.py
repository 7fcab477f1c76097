"""K1 / K3 / K4 against the real mint library (python-mint>=1.24.4, /root/reference/README.md:12).

mint is not installable in the build container, so the first test is skipped there; it runs the moment `import mint`
works (same comparisons as tools/pin_against_mint.py, which also writes the report).  The second test runs the same
machinery against a stand-in module that answers mint's Python API from the oracle, so that the pin script itself is
known to work: key recovery through unit data vectors, the diff, the summation-order vote, the report."""
import os
import sys
import types

import numpy
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, 'tools'))
import pin_against_mint as pin  # noqa: E402


def test_k1_k3_k4_against_real_mint(oracle):
    """call sites /root/reference/nemoflux/field.py:44-49,90-95,102: same (cell, edge) keys, weights to 1e-12, integrals
    to 1e-12 of sum |w f| on C1, C2, the singular case, both closed loops, data/sa and the seam / node stress cases"""
    mint = pytest.importorskip('mint')
    rep = pin.run(mint, oracle)
    bad = [r for r in rep['cases'] if not r['ok']]
    assert not bad, [(r['name'], r['keys_only_mint'], r['keys_only_oracle'], r['weights_max_rel'],
                      r['integral_rel_err_map'], r['decisions_in_question']) for r in bad]
    for r in rep['cases']:
        if r.get('expect') is not None:
            assert r['mint_vs_expect'] <= 1e-10 * max(1.0, abs(r['expect']))


@pytest.mark.gpu
def test_gpu_against_real_mint(oracle, gpu):
    mint = pytest.importorskip('mint')
    rep = pin.run(mint, oracle, gpu)
    for r in rep['cases']:
        assert r['gpu_keys_identical_to_mint'] and r['gpu_integral_rel_err'] <= 1e-12, r


def _stub_mint(O, order):
    """mint's Python surface (the part nemoflux uses) answered by the oracle, summing in `order`"""
    m = types.ModuleType('mint')
    m.CELL_BY_CELL_DATA = 0
    m.__version__ = f'stub-{order}'

    class Grid(object):
        def setPoints(self, points):
            self.g = O.Grid(points)

    class PolylineIntegral(object):
        def setGrid(self, grid):
            self.grid = grid

        def buildLocator(self, numCellsPerBucket=128, periodX=360., enableFolding=False):
            assert numCellsPerBucket == 128 and periodX == 360. and enableFolding is False      # field.py:47
            self.p = O.PolylineIntegral(self.grid.g, periodX)

        def computeWeights(self, xyz, counterclock=False):
            assert counterclock is False                                                         # field.py:48
            self.p.computeWeights(xyz)

        def getIntegral(self, data, placement):
            assert placement == m.CELL_BY_CELL_DATA
            return self.p.getIntegral(data, order)

    class VectorInterp(object):
        def setGrid(self, grid):
            self.grid = grid

        def buildLocator(self, numCellsPerBucket=128, periodX=360., enableFolding=False):
            self.v = O.VectorInterp(self.grid.g, periodX)

        def findPoints(self, points, tol2=1.e-12):
            return self.v.findPoints(points, tol2)

        def getFaceVectors(self, data, placement=0):
            return self.v.getFaceVectors(data)

    m.Grid, m.PolylineIntegral, m.VectorInterp = Grid, PolylineIntegral, VectorInterp
    return m


@pytest.mark.parametrize('order', ['map', 'list'])
def test_pin_script_machinery_on_a_stand_in(oracle, order):
    """the pin script recovers a weight map through getIntegral alone and votes for the right summation order"""
    rep = pin.run(_stub_mint(oracle, order), oracle, wanted={'c1', 'singular', 'seam', 'gridline', 'zero_length', 'sa'},
                  budget=3000)
    assert rep['all_ok'] and len(rep['cases']) == 6
    for r in rep['cases']:
        assert r['keys_identical'] and r['probe_covers_mint_map'] and r['weights_max_ulp'] == 0, r
        assert r['weights_bit_equal'] == r['weights_compared'] > 0
        assert r[f'integral_bit_equal_{order}'] and r['vinterp_max_rel'] == 0.0
        assert r['order_bit_equal_hits'][order] == r['order_trials']
    other = 'list' if order == 'map' else 'map'
    assert rep['summation_order_votes'][order] > rep['summation_order_votes'][other]
    assert rep['summation_order_verdict'] == order
    c1 = [r for r in rep['cases'] if r['name'] == 'c1'][0]
    assert abs(c1['integral_mint'] - 360.0) < 1e-10 and c1['keys_mint'] == c1['keys_oracle']
    # a stand-in that drops one key is caught
    bad = _stub_mint(oracle, order)
    real_get = bad.PolylineIntegral.getIntegral

    def lossy(self, data, placement):
        d = numpy.array(data, numpy.float64, copy=True).reshape(-1, 4)
        d[72, 0] = 0.0                      # first cell of the C1 path
        return real_get(self, d, placement)
    bad.PolylineIntegral.getIntegral = lossy
    rep = pin.run(bad, oracle, wanted={'c1'}, budget=3000)
    assert not rep['all_ok'] and rep['cases'][0]['keys_only_oracle'] == [72 * 4]
