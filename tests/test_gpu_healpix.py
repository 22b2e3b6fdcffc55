"""GPU: HEALPix SHT / iSHT (C ABI ace_hpx_forward / ace_hpx_inverse) vs the live-reference vectors and the oracle."""
import numpy as np
import pytest
import torch

from tests.util import load_golden

pytestmark = pytest.mark.gpu

RTOL = 5e-5  # split-bf16 Legendre GEMM + fp32 ring DFT, relative to the field / spectrum maximum


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("nside", [4, 8, 16])
def test_reference_vectors(nside):
    import ace_b200

    g = load_golden("ref_live_healpix.npz")
    lmax = mmax = 2 * nside - 1
    fwd = ace_b200.HealpixSHT(nside, lmax=lmax, mmax=mmax, quad_weights="ring", ring_weights=g[f"n{nside}.w"])
    inv = ace_b200.HealpixISHT(nside, lmax=lmax, mmax=mmax)
    c = fwd(torch.from_numpy(g[f"n{nside}.x"]).cuda())
    ref = torch.from_numpy(g[f"n{nside}.sht"])
    assert c.dtype == torch.complex64 and tuple(c.shape) == tuple(ref.shape)
    assert _rel(torch.view_as_real(c).cpu(), torch.view_as_real(ref)) < RTOL
    y = inv(torch.from_numpy(g[f"n{nside}.spec"]).cuda())
    assert _rel(y.cpu(), torch.from_numpy(g[f"n{nside}.isht"])) < RTOL


def test_config5_shape_vs_oracle_and_round_trip():
    """BASELINE configs[4] shape: nside 64, lmax = mmax = 127, batch 4 x channels; fme/core/cuhpx/test_sht.py:32-52 property."""
    import ace_b200
    from oracle import healpix as oh

    nside, lmax = 64, 127
    w = oh.uniform_weights(nside)  # the data-file ring weights are not available on the GPU box
    fwd = ace_b200.HealpixSHT(nside, lmax=lmax, mmax=lmax, quad_weights="none")
    inv = ace_b200.HealpixISHT(nside, lmax=lmax, mmax=lmax)
    torch.manual_seed(0)
    x = torch.randn(4, 3, 12 * nside**2)
    c = fwd(x.cuda())
    assert tuple(c.shape) == (4, 3, lmax, lmax)
    co = oh.SHT(nside, lmax, lmax, w)(x)
    assert _rel(torch.view_as_real(c).cpu(), torch.view_as_real(co)) < RTOL
    y = inv(c)
    assert _rel(y.cpu(), oh.iSHT(nside, lmax, lmax)(co)) < RTOL
    back, again = y, inv(fwd(y))
    rms = float((back - again).pow(2).mean().sqrt())
    assert rms < 1e-3


@pytest.mark.parametrize("nside,lmax,mmax,fields", [
    (1, 2, 2, 3),      # three rings of four pixels: the shortest rings (half = 2: the main loop of the ring kernels is empty)
    (2, 4, 5, 11),     # mmax = nlon / 2 + 1 (the module default): the Nyquist order of the equatorial rings
    (3, 6, 6, 9),      # ring lengths that are not powers of two, a ragged field group
    (40, 81, 70, 10),  # 160-pixel equatorial rings: 2 orders per lane in one pass + 1 in the next, polar rings on the 1-order path
])
def test_ring_length_dispatch_vs_oracle(nside, lmax, mmax, fields):
    """Every branch of the ring kernels' orders-per-lane dispatch and the shortest rings, against the oracle port of the
    reference's ring loop (fme/core/cuhpx/tools.py:34-83)."""
    import ace_b200
    from oracle import healpix as oh

    fwd = ace_b200.HealpixSHT(nside, lmax=lmax, mmax=mmax, quad_weights="none")
    inv = ace_b200.HealpixISHT(nside, lmax=lmax, mmax=mmax)
    torch.manual_seed(nside)
    x = torch.randn(fields, 12 * nside**2)
    co = oh.SHT(nside, lmax, mmax, oh.uniform_weights(nside))(x)
    c = fwd(x.cuda())
    assert tuple(c.shape) == (fields, lmax, mmax)
    assert _rel(torch.view_as_real(c).cpu(), torch.view_as_real(co)) < RTOL
    spec = torch.randn(fields, lmax, mmax, dtype=torch.complex64)
    assert _rel(inv(spec.cuda()).cpu(), oh.iSHT(nside, lmax, mmax)(spec)) < RTOL


def test_contract():
    import ace_b200

    with pytest.raises(ValueError):
        ace_b200.HealpixSHT(8, quad_weights="ring")  # ring weights must be supplied
    f = ace_b200.HealpixSHT(8, lmax=15, mmax=15, quad_weights="none")
    with pytest.raises(ace_b200.AceError):
        f(torch.zeros(768))  # CPU tensor: no fallback
    assert f(torch.zeros(0, 768, device="cuda")).shape == (0, 15, 15)
    c = f(torch.ones(768, device="cuda"))  # unbatched input works (the only shape the reference handles correctly)
    assert tuple(c.shape) == (15, 15)
    c00 = c[0, 0].clone()
    c[0, 0] = 0
    assert c.abs().max() < 5e-2 * abs(c00)  # a constant field is nearly pure l = m = 0 (uniform ring weights: inexact quadrature)
