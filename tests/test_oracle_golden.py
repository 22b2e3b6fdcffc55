"""CPU: the oracle restatement against the committed golden vectors (reference-stored and live-reference)."""
import numpy as np
import pytest
import torch

from oracle import sht as osht
from tests.util import NET_GOLDENS, load_golden, oracle_net_from_golden


def test_oracle_sht_matches_reference_stored_golden():
    # fme/core/benchmark/test_benchmark.py:43-57 ; default grid "lobatto"
    g = load_golden("ref_stored_sht_regression.npz")
    x = torch.from_numpy(g["x"])
    y = osht.RealSHT(9, 18)(x)
    torch.testing.assert_close(y, torch.from_numpy(g["sht_output"]))
    z = osht.InverseRealSHT(9, 18)(y)
    torch.testing.assert_close(z, torch.from_numpy(g["isht_output"]))


def test_oracle_sht_all_grids_vs_live_reference_vectors():
    g = load_golden("ref_live_sht_cases.npz")
    for i in range(int(g["ncases"])):
        nlat, nlon, lmax, mmax = (int(v) for v in g[f"c{i}.meta"])
        grid = str(g[f"c{i}.grid"])
        fwd = osht.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid)
        inv = osht.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid)
        torch.testing.assert_close(fwd(torch.from_numpy(g[f"c{i}.x"])), torch.from_numpy(g[f"c{i}.sht"]), rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(inv(torch.from_numpy(g[f"c{i}.spec_in"])), torch.from_numpy(g[f"c{i}.isht"]), rtol=1e-6, atol=1e-6)
        if f"c{i}.fwd_table" in g.files:
            np.testing.assert_array_equal(fwd.weights.numpy(), g[f"c{i}.fwd_table"])
            np.testing.assert_array_equal(inv.pct.numpy(), g[f"c{i}.inv_table"])


@pytest.mark.parametrize("name", NET_GOLDENS)
def test_oracle_net_matches_golden(name):
    g = load_golden(name)
    net = oracle_net_from_golden(g)
    with torch.no_grad():
        y = net(torch.from_numpy(g["x"]))
    ref = torch.from_numpy(g["output"])
    torch.testing.assert_close(y, ref, rtol=1e-5, atol=1e-5 * float(ref.abs().max()))


def test_oracle_net_seeded_construction_matches_reference_stored_golden():
    # fme/ace/models/modulus/test_sfnonet.py:13-36: manual_seed(0), build, randn input
    from oracle import sfno as osfno

    g = load_golden("ref_stored_sfnonet_output_is_unchanged.npz")
    torch.manual_seed(0)
    net = osfno.SphericalFourierNeuralOperatorNet(
        (9, 18), 2, 3, embed_dim=16, num_layers=2, operator_type="diagonal", data_grid="equiangular"
    )
    x = torch.randn(4, 2, 9, 18)
    torch.testing.assert_close(x, torch.from_numpy(g["x"]), rtol=0, atol=0)
    with torch.no_grad():
        torch.testing.assert_close(net(x), torch.from_numpy(g["output"]))


def test_sht_properties():
    # fme/test_harmonics.py:10-42: constant field -> only (0,0); round trip idempotent
    for grid in ["legendre-gauss", "equiangular"]:
        fwd = osht.RealSHT(16, 32, grid=grid)
        inv = osht.InverseRealSHT(16, 32, grid=grid)
        c = fwd(torch.ones(1, 16, 32))
        c00 = c[0, 0, 0].clone()
        c[0, 0, 0] = 0
        assert c.abs().max() < 1e-6 and abs(c00.real - 2 * np.sqrt(np.pi)) < 1e-5
        if grid == "legendre-gauss":  # exact quadrature -> isht(sht(.)) is a projection
            torch.manual_seed(0)
            x = inv(fwd(torch.randn(2, 16, 32)))
            torch.testing.assert_close(inv(fwd(x)), x, rtol=0, atol=2e-6)
