"""CPU: the plugin boundary -- builder fields, registry semantics, state-dict / seeded-init parity with the oracle."""
import dataclasses

import pytest
import torch

import ace_b200
from ace_b200.registry import B200SphericalFourierNeuralOperatorBuilder
from oracle import sfno as osfno

# field names and defaults of fme/ace/registry/sfno.py:21-42
REFERENCE_BUILDER_FIELDS = dict(
    spectral_transform="sht", filter_type="linear", operator_type="diagonal", scale_factor=1,
    residual_filter_factor=1, embed_dim=256, num_layers=12, hard_thresholding_fraction=1.0,
    normalization_layer="instance_norm", use_mlp=True, activation_function="gelu", encoder_layers=1,
    pos_embed=True, big_skip=True, rank=1.0, factorization=None, separable=False, complex_network=True,
    complex_activation="real", spectral_layers=1, checkpointing=0, data_grid="legendre-gauss",
)


def test_builder_has_reference_fields_and_defaults():
    got = {f.name: f.default for f in dataclasses.fields(B200SphericalFourierNeuralOperatorBuilder)}
    assert got == REFERENCE_BUILDER_FIELDS


def test_selector_round_trip_like_reference():
    # fme/ace/registry/test_sfno.py:20-62 pattern: build from a {type, config} selector, defaults normalised
    sel = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config={"embed_dim": 16, "num_layers": 2})
    assert sel.config["operator_type"] == "diagonal" and sel.config["embed_dim"] == 16
    assert "B200SphericalFourierNeuralOperatorNet" in ace_b200.ModuleSelector.get_available_types()
    mod = sel.build(3, 4, ace_b200.DatasetInfo(img_shape=(9, 18)))
    assert isinstance(mod.torch_module, ace_b200.SphericalFourierNeuralOperatorNet)
    with pytest.raises(TypeError):
        mod(torch.zeros(1, 3, 9, 18), labels=object())
    with pytest.raises(ValueError):  # strict: unknown key (dacite strict=True in the reference)
        ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config={"embed_dims": 16})
    with pytest.raises(ValueError):
        sel.build(3, 4, ace_b200.DatasetInfo(img_shape=(9, 18), all_labels=frozenset({"x"})))
    with pytest.raises(KeyError):
        ace_b200.ModuleSelector(type="NoSuchNet", config={})


@pytest.mark.parametrize("operator_type", ["dhconv", "diagonal"])
@pytest.mark.parametrize("norm", ["instance_norm", "none"])
def test_state_dict_and_seeded_init_identical_to_oracle(operator_type, norm):
    kw = dict(embed_dim=12, num_layers=3, operator_type=operator_type, normalization_layer=norm)
    torch.manual_seed(11)
    a = ace_b200.SphericalFourierNeuralOperatorNet(B200SphericalFourierNeuralOperatorBuilder(**kw), img_shape=(10, 20), in_chans=3, out_chans=5)
    torch.manual_seed(11)
    b = osfno.SphericalFourierNeuralOperatorNet((10, 20), 3, 5, **kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]), k
    # wrapper prefix used in checkpoints (fme/core/step/single_module.py:503-507) round-trips
    a.load_state_dict({k: v for k, v in sb.items()})


def test_params_object_overrides_kwargs_like_reference():
    class P:
        embed_dim = 8
        num_layers = 1
        operator_type = "dhconv"
        data_grid = "legendre-gauss"

    net = ace_b200.SphericalFourierNeuralOperatorNet(P(), img_shape=(8, 16), in_chans=2, out_chans=2, embed_dim=999, num_layers=7)
    assert net.embed_dim == 8 and net.num_layers == 1 and len(net.blocks) == 1
    assert tuple(net.blocks[0].filter.filter.weight.shape) == (8, 8, 8, 2)
    # params=None -> reference defaults: diagonal operator on an equiangular data grid (sfnonet.py:461)
    net = ace_b200.SphericalFourierNeuralOperatorNet(None, img_shape=(8, 16), in_chans=2, out_chans=2, embed_dim=4, num_layers=1)
    assert net.data_grid == "equiangular" and tuple(net.blocks[0].filter.filter.weight.shape) == (4, 4, 8, 9, 2)


@pytest.mark.parametrize(
    "kw", [dict(filter_type="non-linear"), dict(scale_factor=2), dict(spectral_transform="fft"), dict(use_mlp=False),
           dict(normalization_layer="layer_norm"), dict(factorization="cp"), dict(residual_filter_factor=2)]
)
def test_unsupported_options_raise_loudly(kw):
    with pytest.raises(NotImplementedError):
        ace_b200.SphericalFourierNeuralOperatorNet(B200SphericalFourierNeuralOperatorBuilder(**kw), img_shape=(8, 16), in_chans=2, out_chans=2)


def test_sht_module_contract_on_cpu():
    s = ace_b200.RealSHT(12, 24)
    assert (s.lmax, s.mmax, s.grid) == (11, 13, "lobatto")  # class default grid is lobatto (fme/sht_fix.py:70,95)
    s = ace_b200.InverseRealSHT(12, 24, grid="equiangular", lmax=5)
    assert (s.lmax, s.mmax) == (5, 13)
    with pytest.raises(NotImplementedError):
        ace_b200.RealSHT(12, 24, grid="healpix")
    with pytest.raises(ValueError):
        ace_b200.RealSHT(12, 24, grid="nope")
    with pytest.raises(ace_b200.AceError):
        ace_b200.RealSHT(12, 24)(torch.zeros(1, 12, 24))  # no CPU path
