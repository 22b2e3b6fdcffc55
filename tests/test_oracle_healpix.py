"""CPU, build container only: oracle/healpix.py against the reference's own cuHPX SHT / iSHT classes."""
import numpy as np
import pytest
import torch

from oracle import healpix as oh
from oracle import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="/root/reference not present (GPU box)")


@pytest.mark.parametrize("nside", [4, 8, 16])
def test_oracle_healpix_equals_reference(nside):
    ref = refload.load_cuhpx()
    lmax = mmax = 2 * nside - 1
    w = ref.apply_ring_weight(nside)
    r_f, r_i = ref.SHT(nside, lmax=lmax, mmax=mmax, quad_weights="ring"), ref.iSHT(nside, lmax=lmax, mmax=mmax)
    o_f, o_i = oh.SHT(nside, lmax, mmax, w), oh.iSHT(nside, lmax, mmax)
    assert torch.equal(r_f.weights, o_f.weights) and torch.equal(r_i.pct, o_i.pct)
    torch.manual_seed(nside)
    x = torch.randn(3, 12 * nside**2)
    c_o = o_f(x)
    # the reference's ring loops take the ring count from ``ftm.shape[0]`` (tools.py:40,59), which is only the ring axis for
    # UNBATCHED input (its own test, fme/core/cuhpx/test_sht.py:32-52, is 1-D); the oracle is the per-field transform, so
    # it is pinned field by field
    for i in range(x.shape[0]):
        c_r = r_f(x[i])
        # (batched vs unbatched einsum differ in summation order: 1 ulp)
        torch.testing.assert_close(torch.view_as_real(c_o[i]), torch.view_as_real(c_r), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(o_i(c_o)[i], r_i(c_r.clone()), rtol=1e-5, atol=1e-5)
    # uniform weights = the reference's quad_weights != "ring" branch
    r_u = ref.SHT(nside, lmax=lmax, mmax=mmax, quad_weights="none")
    assert torch.equal(r_u.weights, oh.SHT(nside, lmax, mmax, oh.uniform_weights(nside)).weights)
