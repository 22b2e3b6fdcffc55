"""GPU: aggregator reductions (ace_b200.metrics, C ABI ace_weighted_moments / ace_zonal_mean / ace_power_spectrum) vs the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _data(seed=0, shape=(2, 7, 48, 96)):
    g = torch.Generator().manual_seed(seed)
    x, t = torch.randn(shape, generator=g) + 0.3, torch.randn(shape, generator=g)
    lat = torch.linspace(-88, 88, shape[-2])
    w = torch.cos(torch.deg2rad(lat))[:, None].expand(shape[-2], shape[-1]).contiguous()
    return x, t, w


@pytest.mark.parametrize("shape", [(2, 7, 48, 96), (50, 180, 360), (3, 9, 18), (1, 1, 1, 45, 96)])
def test_area_weighted_reductions_match_oracle(shape):
    import ace_b200.metrics as am
    from oracle import metrics as om

    x, t, w = _data(0, shape)
    ops = am.LatLonOperations(w)
    xd, td = x.cuda(), t.cuda()
    tol = dict(rtol=2e-6, atol=2e-6)  # the oracle sums in fp32; the device path accumulates in fp64
    torch.testing.assert_close(ops.area_weighted_mean(xd).cpu(), om.weighted_mean(x, w), **tol)
    torch.testing.assert_close(ops.area_weighted_mean(xd, keepdim=True).cpu(), om.weighted_mean(x, w, keepdim=True), **tol)
    torch.testing.assert_close(ops.area_weighted_sum(xd).cpu(), om.weighted_sum(x, w), rtol=3e-6, atol=3e-4)
    torch.testing.assert_close(ops.area_weighted_std(xd).cpu(), om.weighted_std(x, w), **tol)
    torch.testing.assert_close(ops.area_weighted_mean_bias(td, xd).cpu(), om.weighted_mean_bias(t, x, w), **tol)
    torch.testing.assert_close(ops.area_weighted_rmse(td, xd).cpu(), om.root_mean_squared_error(t, x, w), **tol)
    torch.testing.assert_close(ops.zonal_mean(xd).cpu(), om.zonal_mean(x), **tol)
    st = ops.area_weighted_statistics(xd, td)
    torch.testing.assert_close(st["rmse"].cpu(), om.root_mean_squared_error(t, x, w), **tol)


def test_zero_weight_nans_are_ignored_and_contract():
    import ace_b200
    import ace_b200.metrics as am
    from oracle import metrics as om

    x, _, w = _data(1)
    w[0] = 0.0
    x[..., 0, :] = float("nan")
    ops = am.LatLonOperations(w)
    got = ops.area_weighted_mean(x.cuda()).cpu()
    assert torch.isfinite(got).all()
    torch.testing.assert_close(got, om.weighted_mean(x, w), rtol=2e-6, atol=2e-6)
    with pytest.raises(ace_b200.AceError):
        ops.area_weighted_mean(x)  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        am.LatLonOperations(torch.rand(8, 16))  # not longitudinally uniform (gridded_ops.py:305-311)
    assert ops.area_weighted_mean(torch.zeros(0, 48, 96, device="cuda")).shape == (0,)


def test_power_spectrum_matches_oracle():
    import ace_b200
    import ace_b200.metrics as am
    from oracle import metrics as om
    from oracle import sht as osht

    x, _, w = _data(2, (3, 5, 48, 96))
    ops = am.LatLonOperations(w, grid="legendre-gauss")
    sht = ops.get_real_sht()
    assert isinstance(sht, ace_b200.RealSHT)
    got = am.spherical_power_spectrum(x.cuda(), sht).cpu()
    ref = om.spherical_power_spectrum(x, osht.RealSHT(48, 96, grid="legendre-gauss"))
    assert got.shape == ref.shape == (3, 5, 48)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5 * float(ref.max()))


def test_window_aggregators_match_oracle():
    """Row f3: time-mean maps and per-step reduced metrics accumulated over three windows on the device vs the oracle
    restatement of fme/ace/aggregator/inference/{time_mean,reduced}.py (itself pinned to the reference's own class bodies)."""
    from ace_b200 import metrics as bm
    from oracle import aggregator as oa

    names = ["b", "a", "c"]
    B, T, H, W = 2, 4, 18, 36
    g = torch.Generator().manual_seed(7)
    lat = torch.linspace(-85, 85, H)
    w = torch.cos(torch.deg2rad(lat))[:, None].expand(H, W).contiguous()
    ops = bm.LatLonOperations(w)
    tm_dev, tm_ref = bm.TimeMeanAggregator(names), oa.TimeMean()
    ma_dev = bm.MeanAggregator(ops, names, n_timesteps=3 * T)
    refs = {k: oa.ReducedMetric(fn, 3 * T) for k, fn in oa.mean_aggregator_metrics(w).items()}
    for k in range(3):
        gen = torch.randn(B, T, len(names), H, W, generator=g) * 3.0 + 1.5
        tgt = gen + 0.3 * torch.randn(B, T, len(names), H, W, generator=g)
        tm_dev.record_batch(gen.cuda(), i_time_start=k * T)
        ma_dev.record_batch(tgt.cuda(), gen.cuda(), i_time_start=k * T)
        gd = {n: gen[:, :, i] for i, n in enumerate(names)}
        td = {n: tgt[:, :, i] for i, n in enumerate(names)}
        tm_ref.record_batch(gd, k * T)
        for r in refs.values():
            r.record(td, gd, k * T)
    got, want = tm_dev.get_data(), tm_ref.get_data()
    assert list(got) == list(want) == sorted(names)
    for n in names:
        torch.testing.assert_close(got[n].cpu(), want[n], rtol=2e-6, atol=2e-6)
    series = ma_dev.get()
    assert set(series) == set(refs)
    for m, r in refs.items():
        ref = r.get()
        for n in names:
            torch.testing.assert_close(series[m][n].cpu(), ref[n], rtol=5e-6, atol=5e-6)
    import ace_b200

    with pytest.raises(ace_b200.AceError):
        tm_dev.record_batch(torch.zeros(B, T, len(names), H, W))  # CPU tensor: no fallback
