"""CPU: the B200 noise-conditioned SFNO module's contract with the reference (no kernels run): parameter names / shapes /
seeded initial values equal to the oracle restatement (itself bit-identical to the reference), checkpoint layouts, option errors."""
import pytest
import torch

import ace_b200
from ace_b200 import csfno as bc
from oracle import csfno as oc
from tests.util import CSFNO_GOLDENS, load_csfno_case


def _b200_from_case(kwargs, dims):
    kw = dict(kwargs)
    grid, shape, ci, co = kw.pop("data_grid"), kw.pop("img_shape"), kw.pop("in_chans"), kw.pop("out_chans")
    return bc.get_lat_lon_sfnonet(bc.SFNONetConfig(**kw), ci, co, shape, grid, bc.ContextConfig(**dims))


def test_seeded_construction_equals_oracle():
    cfg = dict(embed_dim=16, num_layers=2, noise_embed_dim=8, noise_type="isotropic", affine_norms=True, normalize_big_skip=True)
    sel = ace_b200.ModuleSelector(type="B200NoiseConditionedSFNO", config=cfg)
    assert sel.config["mlp_ratio"] == 2.0 and sel.config["data_grid"] == "legendre-gauss"  # defaults normalised into the config
    torch.manual_seed(3)
    m = sel.build(5, 4, ace_b200.DatasetInfo(img_shape=(12, 24))).torch_module
    torch.manual_seed(3)
    o = oc.SphericalFourierNeuralOperatorNet((12, 24), 5, 4, oc.ContextConfig(embed_dim_noise=8), embed_dim=16, num_layers=2,
                                             affine_norms=True, normalize_big_skip=True, data_grid="legendre-gauss")
    sm, so = m.conditional_model.state_dict(), o.state_dict()
    assert list(sm.keys()) == list(so.keys())
    for k in so:
        assert torch.equal(sm[k], so[k]), k
    assert list(m.state_dict().keys()) == ["conditional_model." + k for k in so]  # NoiseConditionedModel prefix (stochastic_sfno.py:86)
    assert m._inverse_sht is m.conditional_model.itrans_up and m._lmax == 12 and m._mmax == 13


@pytest.mark.parametrize("name", CSFNO_GOLDENS)
def test_reference_checkpoints_load(name):
    kwargs, dims, state, x, ctx, y = load_csfno_case(name)
    net = _b200_from_case(kwargs, dims)
    res = net.load_state_dict(state)  # incl. the [G, I, O, L, 2] layout of the stored checkpoint golden
    assert not res.missing_keys and not res.unexpected_keys
    conv = net.blocks[0].filter.filter
    cg = conv.spectral_channels // conv.num_groups
    assert tuple(conv.weight.shape) == (conv.num_groups, net.modes_lat, cg, cg, 2)
    dense = dict((k, b) for k, _, b in net.device_parameters())["blocks.0.filter.filter.weight"]
    if dense is not None:  # what the device library receives is always the dense per-degree operator
        assert tuple(dense().shape) == (1, net.modes_lat, net.embed_dim, net.embed_dim, 2)


FOLDED_CASES = [
    dict(filter_num_groups=4),
    dict(filter_num_groups=2, filter_preserves_global_mean=True),
    dict(spectral_lora_rank=3, spectral_lora_alpha=5.0, lora_rank=2),
    dict(filter_num_groups=2, spectral_ratio=0.5, spectral_lora_rank=2, lora_rank=3, lora_alpha=1.5, filter_preserves_global_mean=True),
    dict(filter_residual=True, filter_output=True, filter_num_groups=2),
]


@pytest.mark.parametrize("extra", FOLDED_CASES)
def test_folded_parameters_reproduce_the_adapted_network(extra):
    """Grouped / LoRA / mean-preserving / bottlenecked layers are handed to the device library as plain dense parameters
    (csfno.py ``device_parameters``).  The fold is checked here without a GPU: the oracle network WITH the options (bit-identical
    to the reference, tests/test_oracle_csfno.py) against the oracle network WITHOUT them carrying the folded parameters."""
    cc = dict(embed_dim_noise=5, embed_dim_pos=2, embed_dim_labels=2)
    kw = dict(embed_dim=16, num_layers=2, affine_norms=True, normalize_big_skip=True, **extra)
    torch.manual_seed(11)
    full = oc.SphericalFourierNeuralOperatorNet((10, 20), 3, 2, oc.ContextConfig(**cc), data_grid="legendre-gauss", **kw).eval()
    torch.manual_seed(11)
    mod = bc.get_lat_lon_sfnonet(bc.SFNONetConfig(**kw), 3, 2, (10, 20), "legendre-gauss", bc.ContextConfig(**cc))
    sm, so = mod.state_dict(), full.state_dict()
    assert list(sm.keys()) == list(so.keys())
    for k in so:
        assert torch.equal(sm[k], so[k]), k  # same construction order -> same seeded draws
    with torch.no_grad():
        for prm in full.parameters():
            prm.add_(0.2 * torch.randn_like(prm))  # lora_B / lora_up start at zero
    mod.load_state_dict(full.state_dict())
    plain_kw = {k: v for k, v in kw.items() if k in ("embed_dim", "num_layers", "affine_norms", "normalize_big_skip", "filter_residual", "filter_output")}
    plain = oc.SphericalFourierNeuralOperatorNet((10, 20), 3, 2, oc.ContextConfig(**cc), data_grid="legendre-gauss", **plain_kw).eval()
    folded = {}
    for key, sources, build in mod.device_parameters():
        folded[key] = sources[0].detach().clone() if build is None else build()
        assert folded[key].dtype == torch.float32 and folded[key].is_contiguous()
    assert sorted(folded) == sorted(plain.state_dict())  # exactly the keys the un-adapted network (and the device library) takes
    plain.load_state_dict(folded)
    x = torch.randn(2, 3, 10, 20)
    ctx = dict(labels=torch.randn(2, 2), noise=torch.randn(2, 5, 10, 20), embedding_pos=torch.randn(2, 2, 10, 20))
    with torch.no_grad():
        want, got = full(x, oc.Context(**ctx)), plain(x, oc.Context(**ctx))
    assert float((want - got).abs().max() / want.abs().max()) < 2e-6


def test_spectral_ratio_validation_is_the_references():
    with pytest.raises(ValueError, match="must be in"):
        bc.get_lat_lon_sfnonet(bc.SFNONetConfig(embed_dim=8, num_layers=1, spectral_ratio=0.0), 2, 2, (8, 16), "legendre-gauss")
    with pytest.raises(ValueError, match="not divisible"):
        bc.get_lat_lon_sfnonet(bc.SFNONetConfig(embed_dim=8, num_layers=1, spectral_ratio=0.75, filter_num_groups=4), 2, 2, (8, 16), "legendre-gauss")
    with pytest.raises(NotImplementedError, match="round-trip"):  # the fold cannot express post_proj(isht(sht(pre_proj(x)))) as a residual
        bc.get_lat_lon_sfnonet(bc.SFNONetConfig(embed_dim=8, num_layers=1, spectral_ratio=0.5, filter_residual=True), 2, 2, (8, 16), "legendre-gauss")


def test_clip_latent_global_means_buffers_follow_the_checkpoint():
    kw = dict(embed_dim=8, num_layers=1, clip_latent_global_means=True)
    o = oc.SphericalFourierNeuralOperatorNet((8, 16), 2, 2, data_grid="legendre-gauss", **kw)
    m = bc.get_lat_lon_sfnonet(bc.SFNONetConfig(**kw), 2, 2, (8, 16), "legendre-gauss")
    assert list(m.state_dict().keys()) == list(o.state_dict().keys())
    assert torch.isinf(m._gm_min).all() and torch.isinf(m._gm_max).all()  # fresh envelope: the device kernel stays a no-op
    with torch.no_grad():
        o._gm_min.fill_(-0.5)
        o._gm_max.fill_(0.25)
    m.load_state_dict(o.state_dict())
    keys = [k for k, _, _ in m.device_parameters()]
    assert keys[-2:] == ["_gm_min", "_gm_max"] and float(m._gm_max.max()) == 0.25
    m.request_latent_global_mean_envelope_reset()  # API of the reference; inference never updates the envelope


def test_unsupported_options_raise_and_cpu_input_is_rejected():
    for kw in (dict(global_layer_norm=True), dict(filter_type="makani-linear"), dict(local_blocks=[0]),
               dict(use_mlp=False), dict(activation_function="relu")):
        with pytest.raises(NotImplementedError):
            bc.get_lat_lon_sfnonet(bc.SFNONetConfig(embed_dim=8, num_layers=1, **kw), 2, 2, (8, 16))
    # context wider than 64 channels is supported (the reference builder's default noise_embed_dim = 256 builds)
    bc.get_lat_lon_sfnonet(bc.SFNONetConfig(embed_dim=8, num_layers=1), 2, 2, (8, 16), context_config=bc.ContextConfig(embed_dim_noise=65))
    sel = ace_b200.ModuleSelector(type="B200NoiseConditionedSFNO", config=dict(embed_dim=8, num_layers=1))
    assert sel.build(2, 2, ace_b200.DatasetInfo(img_shape=(8, 16))).torch_module.embed_dim == 256
    with pytest.raises(ValueError):
        ace_b200.ModuleSelector(type="B200NoiseConditionedSFNO", config=dict(context_pos_embed_dim=4))  # with pos_embed=True
    net = bc.get_lat_lon_sfnonet(bc.SFNONetConfig(embed_dim=8, num_layers=1), 2, 2, (8, 16)).requires_grad_(False)
    with pytest.raises(ace_b200.AceError):
        net(torch.zeros(1, 2, 8, 16), bc.Context())


def test_folded_group_weights_reproduce_the_references_stored_8_group_block_golden():
    """The dense per-degree operator the device library receives for a grouped filter (``_SpectralConvS2.effective_weight``),
    loaded into an UNGROUPED oracle block, reproduces the reference's stored 8-group block golden
    (fme/core/benchmark/testdata/csfno_block_8_groups-regression.pt)."""
    import os

    import numpy as np

    from oracle.sht import InverseRealSHT, RealSHT
    from tests.util import GOLDEN_DIR

    d = np.load(os.path.join(GOLDEN_DIR, "ref_stored_csfno_block_8_groups.npz"))
    state = {k[2:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("p:")}
    ctx = {k[4:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("ctx:")}
    holder = bc._SpectralConvS2(16, 8, num_groups=8)
    holder.load_state_dict({"weight": state["filter.filter.weight"], "bias": state["filter.filter.bias"]})
    dense = holder.effective_weight()
    assert tuple(dense.shape) == (1, 8, 16, 16, 2)
    blk = oc.FourierNeuralOperatorBlock(RealSHT(9, 18), InverseRealSHT(9, 18), 16, (9, 18),
                                        oc.ContextConfig(embed_dim_noise=4, embed_dim_labels=3, embed_dim_pos=2),
                                        filter_num_groups=1, outer_skip=None).eval()
    blk.load_state_dict({**state, "filter.filter.weight": dense})
    with torch.no_grad():
        out = blk(torch.from_numpy(d["x"]), oc.Context(**ctx))
    torch.testing.assert_close(out, torch.from_numpy(d["y"]), rtol=1e-5, atol=2e-6)
