"""CPU: product-side init-time tables (ace_b200/legendre.py) vs the independent oracle restatement."""
import numpy as np
import pytest

from ace_b200 import legendre as prod
from oracle import legendre as oleg
from oracle import quadrature as oquad
from oracle import sht as osht
from tests.util import load_golden


@pytest.mark.parametrize("n", [2, 3, 9, 16, 33, 180])
def test_quadrature_matches_oracle(n):
    for mine, ref in [
        (prod.legendre_gauss_nodes, oquad.legendre_gauss_weights),
        (prod.clenshaw_curtis_nodes, oquad.clenshaw_curtiss_weights),
    ] + ([(prod.lobatto_nodes, oquad.lobatto_weights)] if n > 2 else []):
        x, w = mine(n)
        xr, wr = ref(n)
        np.testing.assert_allclose(x, xr, rtol=0, atol=1e-15)
        np.testing.assert_allclose(w, wr, rtol=1e-13, atol=1e-16)
        assert abs(w.sum() - 2.0) < 1e-13


@pytest.mark.parametrize("mmax,lmax,n", [(10, 9, 9), (17, 16, 16), (46, 45, 45), (91, 90, 90)])
@pytest.mark.parametrize("inverse,csphase,norm", [(False, True, "ortho"), (True, True, "ortho"), (False, False, "four-pi"), (True, True, "schmidt")])
def test_vectorised_recursion_is_bit_identical_to_loop(mmax, lmax, n, inverse, csphase, norm):
    x = np.cos(np.linspace(0.01, np.pi - 0.01, n))
    a = prod.legendre_table(mmax, lmax, x, norm=norm, inverse=inverse, csphase=csphase)
    b = oleg.legpoly(mmax, lmax, x, norm=norm, inverse=inverse, csphase=csphase)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("grid", ["legendre-gauss", "equiangular", "lobatto"])
def test_fp32_tables_match_oracle_and_live_reference_vectors(grid):
    for nlat, nlon in [(9, 18), (30, 60)]:
        fwd, inv, lmax, mmax = prod.sht_tables(nlat, nlon, grid=grid)
        of, ol, om = osht.forward_table(nlat, nlon, grid=grid)
        oi, _, _ = osht.inverse_table(nlat, nlon, grid=grid)
        assert (lmax, mmax) == (ol, om)
        # after the reference's cast to fp32 the tables agree to 1 ulp (different but equivalent quadrature code)
        np.testing.assert_allclose(fwd.astype(np.float32), of.astype(np.float32), rtol=2e-7, atol=1e-37)
        np.testing.assert_array_equal(inv.astype(np.float32), oi.astype(np.float32))
    g = load_golden("ref_live_sht_cases.npz")
    for i in range(3):  # the three 9x18 cases carry the live reference's fp32 tables
        if str(g[f"c{i}.grid"]) != grid:
            continue
        fwd, inv, _, _ = prod.sht_tables(9, 18, grid=grid)
        np.testing.assert_allclose(fwd.astype(np.float32), g[f"c{i}.fwd_table"], rtol=2e-7, atol=1e-37)
        np.testing.assert_array_equal(inv.astype(np.float32), g[f"c{i}.inv_table"])


def test_table_properties():
    fwd, inv, lmax, mmax = prod.sht_tables(24, 48, grid="legendre-gauss")
    m, l = np.meshgrid(np.arange(mmax), np.arange(lmax), indexing="ij")
    assert np.all(fwd[m > l] == 0.0) and np.all(inv[m > l] == 0.0)  # P_l^m = 0 for l < m
    # discrete orthonormality under Gauss quadrature: sum_k fwd[m,l,k] inv[m,l',k] = delta / (2 pi)
    for mm in (0, 3, 11):
        gram = fwd[mm] @ inv[mm].T * 2 * np.pi
        np.testing.assert_allclose(gram[mm:, mm:], np.eye(lmax - mm), atol=1e-12)
