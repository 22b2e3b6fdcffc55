"""CPU, build container only: oracle/metrics.py against the reference's own fme/core/metrics.py functions."""
import pytest
import torch

from oracle import metrics as om
from oracle import refload
from oracle import sht as osht

pytestmark = pytest.mark.skipif(not refload.available(), reason="/root/reference not present (GPU box)")


def _data(seed=0, shape=(3, 5, 18, 36)):
    g = torch.Generator().manual_seed(seed)
    x, t = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    lat = torch.linspace(-85, 85, shape[-2])
    w = torch.cos(torch.deg2rad(lat))[:, None].expand(shape[-2], shape[-1]).contiguous()
    return x, t, w


def test_oracle_metrics_equal_reference():
    ref = refload.load_metrics()
    x, t, w = _data()
    d = (-2, -1)
    assert torch.equal(om.weighted_sum(x, w), ref.weighted_sum(x, w, dim=d))
    assert torch.equal(om.weighted_mean(x, w), ref.weighted_mean(x, w, dim=d))
    assert torch.equal(om.weighted_mean(x, w, keepdim=True), ref.weighted_mean(x, w, dim=d, keepdim=True))
    assert torch.equal(om.weighted_std(x, w), ref.weighted_std(x, w, dim=d))
    assert torch.equal(om.weighted_mean_bias(t, x, w), ref.weighted_mean_bias(t, x, weights=w, dim=d))
    assert torch.equal(om.root_mean_squared_error(t, x, w), ref.root_mean_squared_error(t, x, weights=w, dim=d))
    # zero-weight NaNs are ignored by both
    w2 = w.clone()
    w2[0] = 0.0
    x2 = x.clone()
    x2[..., 0, :] = float("nan")
    assert torch.equal(om.weighted_mean(x2, w2), ref.weighted_mean(x2, w2, dim=d))
    assert torch.isfinite(om.weighted_mean(x2, w2)).all()


def test_oracle_power_spectrum_equal_reference():
    ref = refload.load_metrics()
    x, _, _ = _data(1)
    sht = osht.RealSHT(18, 36, grid="legendre-gauss")
    assert torch.equal(om.spherical_power_spectrum(x, sht), ref.spherical_power_spectrum(x, sht))
