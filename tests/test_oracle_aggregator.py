"""CPU: oracle/aggregator.py against the reference's own aggregator code (class members cut out of the reference files)."""
import os

import pytest
import torch

from oracle import aggregator as oa

pytestmark = pytest.mark.skipif(not os.path.exists(oa.REFERENCE_ROOT), reason="/root/reference not present (GPU box)")


def _windows(seed=0, B=2, T=4, H=6, W=12, names=("b", "a", "c")):
    g = torch.Generator().manual_seed(seed)
    wins = []
    for k in range(3):
        wins.append(({n: torch.randn(B, T, H, W, generator=g) for n in names}, {n: torch.randn(B, T, H, W, generator=g) for n in names}))
    lat = torch.linspace(-80, 80, H)
    w = torch.cos(torch.deg2rad(lat))[:, None].expand(H, W).contiguous()
    return wins, w


def test_time_mean_equals_reference_member():
    ref_add, _ = oa.reference_snippets()
    wins, _ = _windows()
    tm, ref = oa.TimeMean(), None
    i = 0
    for gen, _ in wins:
        tm.record_batch(gen, i)
        ref = ref_add(ref, gen, ignore_initial=(i == 0))
        i += 4
    assert tm._n_timesteps == 3 + 4 + 4 and tm._n_samples == 2
    for n in ref:
        assert torch.equal(tm._data[n], ref[n])
    assert list(tm.get_data()) == ["a", "b", "c"]


def test_reduced_metric_equals_reference_class():
    _, RefMetric = oa.reference_snippets()
    wins, w = _windows(1)
    fns = oa.mean_aggregator_metrics(w)
    for name, fn in fns.items():
        mine = oa.ReducedMetric(fn, n_timesteps=12)
        ref = RefMetric(device=torch.device("cpu"), compute_metric=fn, n_timesteps=12)
        for k, (gen, tgt) in enumerate(wins):
            mine.record(tgt, gen, 4 * k)
            ref.record(tgt, gen, 4 * k)
        a, b = mine.get(), ref.get()
        for n in b:
            assert torch.equal(a[n], b[n]), (name, n)
