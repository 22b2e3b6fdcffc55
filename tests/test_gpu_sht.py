"""GPU: RealSHT / InverseRealSHT through the C ABI vs the oracle and the committed reference vectors."""
import numpy as np
import pytest
import torch

from tests.util import load_golden

pytestmark = pytest.mark.gpu

# split-bf16 storage (2^-17) + 3-term products through two GEMM stages
RTOL_FIELD = 5e-5


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def test_reference_stored_golden_lobatto():
    import ace_b200

    g = load_golden("ref_stored_sht_regression.npz")
    x = torch.from_numpy(g["x"]).cuda()
    y = ace_b200.RealSHT(9, 18)(x)
    ref = torch.from_numpy(g["sht_output"])
    assert y.dtype == torch.complex64 and tuple(y.shape) == (1, 8, 10)
    assert _rel(torch.view_as_real(y).cpu(), torch.view_as_real(ref)) < RTOL_FIELD
    z = ace_b200.InverseRealSHT(9, 18)(y)
    assert _rel(z.cpu(), torch.from_numpy(g["isht_output"])) < RTOL_FIELD


def test_live_reference_vectors_all_grids():
    import ace_b200

    g = load_golden("ref_live_sht_cases.npz")
    for i in range(int(g["ncases"])):
        nlat, nlon, lmax, mmax = (int(v) for v in g[f"c{i}.meta"])
        grid = str(g[f"c{i}.grid"])
        fwd = ace_b200.RealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid)
        inv = ace_b200.InverseRealSHT(nlat, nlon, lmax=lmax, mmax=mmax, grid=grid)
        y = fwd(torch.from_numpy(g[f"c{i}.x"]).cuda())
        ref = torch.from_numpy(g[f"c{i}.sht"])
        assert _rel(torch.view_as_real(y).cpu(), torch.view_as_real(ref)) < RTOL_FIELD, (i, grid)
        spec = torch.from_numpy(g[f"c{i}.spec_in"]).cuda()
        x = inv(spec)
        assert _rel(x.cpu(), torch.from_numpy(g[f"c{i}.isht"])) < RTOL_FIELD, (i, grid)


@pytest.mark.parametrize("shape,nf", [((48, 96), 8), ((180, 360), 8), ((64, 128), 12), ((45, 96), 8)])  # odd nlat: element-wise stores
def test_tcgen05_path_vs_oracle(shape, nf):
    """Aligned shapes: every GEMM of the transform must run on the tcgen05 kernel and match the oracle."""
    import ace_b200
    from ace_b200 import _lib
    from oracle import sht as osht

    nlat, nlon = shape
    torch.manual_seed(0)
    x = torch.randn(nf, nlat, nlon)
    fwd, inv = ace_b200.RealSHT(nlat, nlon, grid="legendre-gauss"), ace_b200.InverseRealSHT(nlat, nlon, grid="legendre-gauss")
    ofwd, oinv = osht.RealSHT(nlat, nlon, grid="legendre-gauss"), osht.InverseRealSHT(nlat, nlon, grid="legendre-gauss")
    u0, s0 = _lib.get_option("count_umma"), _lib.get_option("count_simt")
    y = fwd(x.cuda())
    z = inv(y)
    torch.cuda.synchronize()
    assert _lib.get_option("count_umma") - u0 == 4 and _lib.get_option("count_simt") == s0
    yo = ofwd(x)
    assert _rel(torch.view_as_real(y).cpu(), torch.view_as_real(yo)) < RTOL_FIELD
    assert _rel(z.cpu(), oinv(yo)) < RTOL_FIELD
    # same transform on the SIMT kernel agrees
    _lib.set_option("force_simt", 1)
    try:
        y2 = fwd(x.cuda())
    finally:
        _lib.set_option("force_simt", 0)
    assert _rel(torch.view_as_real(y2).cpu(), torch.view_as_real(y).cpu()) < RTOL_FIELD


def test_properties_constant_field_and_projection():
    # fme/test_harmonics.py:10-42
    import ace_b200

    fwd, inv = ace_b200.RealSHT(32, 64, grid="legendre-gauss"), ace_b200.InverseRealSHT(32, 64, grid="legendre-gauss")
    c = fwd(torch.ones(8, 32, 64, device="cuda"))
    c00 = c[:, 0, 0].clone()
    c[:, 0, 0] = 0
    assert c.abs().max() < 1e-4 and (c00.real - 2 * np.sqrt(np.pi)).abs().max() < 1e-4
    torch.manual_seed(1)
    x = inv(fwd(torch.randn(8, 32, 64, device="cuda")))
    assert _rel(inv(fwd(x)), x) < RTOL_FIELD


def test_batch_shapes_and_reuse():
    import ace_b200
    from oracle import sht as osht

    fwd = ace_b200.RealSHT(16, 32, grid="equiangular")
    o = osht.RealSHT(16, 32, grid="equiangular")
    torch.manual_seed(2)
    for lead in [(3,), (2, 5), (), (1, 1, 4)]:
        x = torch.randn(*lead, 16, 32)
        y = fwd(x.cuda())
        assert tuple(y.shape) == (*lead, 16, 17)
        assert _rel(torch.view_as_real(y).cpu(), torch.view_as_real(o(x))) < RTOL_FIELD
    assert fwd(torch.zeros(0, 16, 32, device="cuda")).shape == (0, 16, 17)
    with pytest.raises(ace_b200.AceError):
        fwd(torch.zeros(1, 16, 32))  # CPU tensor: no fallback
