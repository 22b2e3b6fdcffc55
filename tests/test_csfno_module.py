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
    assert tuple(net.blocks[0].filter.filter.weight.shape) == (1, net.modes_lat, net.embed_dim, net.embed_dim, 2)


def test_unsupported_options_raise_and_cpu_input_is_rejected():
    for kw in (dict(filter_num_groups=2), dict(global_layer_norm=True), dict(filter_type="makani-linear"), dict(spectral_ratio=0.5),
               dict(filter_residual=True), dict(lora_rank=2), dict(activation_function="relu")):
        with pytest.raises(NotImplementedError):
            bc.get_lat_lon_sfnonet(bc.SFNONetConfig(embed_dim=8, num_layers=1, **kw), 2, 2, (8, 16))
    with pytest.raises(NotImplementedError):
        bc.get_lat_lon_sfnonet(bc.SFNONetConfig(embed_dim=8, num_layers=1), 2, 2, (8, 16), context_config=bc.ContextConfig(embed_dim_noise=65))
    with pytest.raises(ValueError):
        ace_b200.ModuleSelector(type="B200NoiseConditionedSFNO", config=dict(context_pos_embed_dim=4))  # with pos_embed=True
    net = bc.get_lat_lon_sfnonet(bc.SFNONetConfig(embed_dim=8, num_layers=1), 2, 2, (8, 16)).requires_grad_(False)
    with pytest.raises(ace_b200.AceError):
        net(torch.zeros(1, 2, 8, 16), bc.Context())
