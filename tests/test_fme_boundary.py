"""CPU: the fme-side installation hooks executed end to end against a minimal ``fme`` package (tests/fake_fme.py: the
reference's own registry / StepArgs / StepperState source files + small stand-ins for what this image cannot import).

* ``install_into_fme(override=True)``: the B200 builders registered in the REAL ``ModuleSelector`` registry (under their own
  names and over ``SphericalFourierNeuralOperatorNet`` / ``NoiseConditionedSFNO``), strict ``from_state``, config normalised
  with defaults, labels rejected -- ``fme/core/registry/module.py:121-210``, ``fme/ace/registry/sfno.py:14-61``;
* ``patch_torch_harmonics()``: the monkey-patch of ``fme/sht_fix.py:228-229``;
* ``install_step_into_fme(override=True)``: a ``StepSelector``-registered step config with ``SingleModuleStepConfig``'s fields
  whose step object owns a ``FusedStepper`` -- driven through ``StepSelector(type=..., config=...).get_step(...).step(StepArgs)``
  with the one library call replaced by the oracle (no GPU here), against the reference step sequence.
"""
import dataclasses

import pytest
import torch

import ace_b200
from ace_b200 import fme_step
from tests import fake_fme

pytestmark = pytest.mark.skipif(not fake_fme.available(), reason="/root/reference is not present on this box")

IMG = (8, 16)
IN_NAMES = ["a", "b", "f1", "c", "f2"]
OUT_NAMES = ["c", "d1", "a", "b", "d2"]
NAMES = sorted(set(IN_NAMES + OUT_NAMES))
MEANS = {n: 0.1 * (i - 3) for i, n in enumerate(NAMES)}
STDS = {n: 0.5 + 0.25 * i for i, n in enumerate(NAMES)}
NET = dict(embed_dim=8, num_layers=1, operator_type="dhconv")


def test_install_into_fme_registers_over_the_reference_names():
    with fake_fme.installed() as mods:
        FmeSelector = mods["fme.ace.registry.registry"].ModuleSelector
        FmeConfig = mods["fme.ace.registry.registry"].ModuleConfig
        DatasetInfo = mods["fme.core.dataset_info"].DatasetInfo
        cls = ace_b200.install_into_fme(override=True)
        assert issubclass(cls, FmeConfig) and dataclasses.is_dataclass(cls)
        types = set(FmeSelector.get_available_types())
        assert {"B200SphericalFourierNeuralOperatorNet", "SphericalFourierNeuralOperatorNet", "B200NoiseConditionedSFNO",
                "NoiseConditionedSFNO"} <= types
        for name in ("SphericalFourierNeuralOperatorNet", "B200SphericalFourierNeuralOperatorNet"):
            sel = FmeSelector(type=name, config=dict(NET))
            # defaults captured in the serialised config (module.py:150-162), field for field the reference builder's
            assert sel.config["data_grid"] == "legendre-gauss" and sel.config["spectral_transform"] == "sht" and len(sel.config) == 22
            torch.manual_seed(0)
            module = sel.build(n_in_channels=5, n_out_channels=4, dataset_info=DatasetInfo(img_shape=IMG))
            net = module.torch_module
            assert isinstance(net, ace_b200.SphericalFourierNeuralOperatorNet)
            assert net.blocks[0].filter.filter.weight.shape == (8, 8, 8, 2)
            with pytest.raises(TypeError):
                module(torch.zeros(1, 5, *IMG), labels=object())  # unconditional builder (module.py:80-86)
        with pytest.raises(ValueError):
            FmeSelector(type="SphericalFourierNeuralOperatorNet", config=dict(NET, not_a_field=1))  # dacite strict
        with pytest.raises(ValueError):
            FmeSelector(type="SphericalFourierNeuralOperatorNet", config=dict(NET)).build(
                5, 4, DatasetInfo(img_shape=IMG, all_labels=frozenset({"x"})))  # fme/ace/registry/sfno.py:50-53
        # the instances are independent dataclasses (fields were not shared between the generated classes)
        a = FmeSelector(type="SphericalFourierNeuralOperatorNet", config=dict(NET, embed_dim=16))
        b = FmeSelector(type="B200NoiseConditionedSFNO", config=dict(embed_dim=8, num_layers=1, noise_embed_dim=4, noise_type="isotropic"),
                        conditional=False)
        assert a.config["embed_dim"] == 16 and b.config["embed_dim"] == 8 and b.config["noise_embed_dim"] == 4
        nb = FmeSelector(type="NoiseConditionedSFNO", config=dict(embed_dim=8, num_layers=1, noise_embed_dim=4, noise_type="isotropic"))
        m = nb.build(n_in_channels=3, n_out_channels=2, dataset_info=DatasetInfo(img_shape=IMG)).torch_module
        assert isinstance(m, ace_b200.NoiseConditionedModel)
        with pytest.raises(ValueError):
            FmeSelector(type="NoiseConditionedSFNO", config=dict(operator_type="diagonal"))  # stochastic_sfno.py:300-318


def test_patch_torch_harmonics():
    with fake_fme.installed() as mods:
        th = mods["torch_harmonics"]
        ace_b200.patch_torch_harmonics()
        assert th.RealSHT is ace_b200.RealSHT and th.InverseRealSHT is ace_b200.InverseRealSHT
        sht = th.RealSHT(9, 18)  # class default grid is lobatto: lmax = nlat - 1 (fme/sht_fix.py:91-104)
        assert (sht.nlat, sht.nlon, sht.lmax, sht.mmax, sht.grid) == (9, 18, 8, 10, "lobatto")
        isht = th.InverseRealSHT(9, 18, lmax=8, mmax=10, grid="legendre-gauss")
        assert isinstance(isht, torch.nn.Module) and isht.grid == "legendre-gauss"
        with pytest.raises(ace_b200.AceError):
            sht(torch.zeros(1, 9, 18))  # CPU tensor: no fallback


def test_kill_switch_leaves_the_reference_in_place(monkeypatch):
    """ACE_B200_DISABLE=1 (A/B against the reference with an unchanged config): the install functions register the ``B200...`` names as
    aliases of whatever the reference registered and override nothing."""
    monkeypatch.setenv("ACE_B200_DISABLE", "1")
    with fake_fme.installed() as mods:
        FmeSelector = mods["fme.ace.registry.registry"].ModuleSelector
        FmeConfig = mods["fme.ace.registry.registry"].ModuleConfig

        @dataclasses.dataclass
        class RefNet(FmeConfig):
            embed_dim: int = 4

            def build(self, n_in_channels, n_out_channels, dataset_info):
                return torch.nn.Conv2d(n_in_channels, n_out_channels, 1)

        @dataclasses.dataclass
        class RefNoise(RefNet):
            pass

        FmeSelector.register("SphericalFourierNeuralOperatorNet")(RefNet)
        FmeSelector.register("NoiseConditionedSFNO")(RefNoise)
        assert ace_b200.install_into_fme(override=True) is None
        reg = FmeSelector.registry._types
        assert reg["SphericalFourierNeuralOperatorNet"] is RefNet and reg["B200SphericalFourierNeuralOperatorNet"] is RefNet
        assert reg["NoiseConditionedSFNO"] is RefNoise and reg["B200NoiseConditionedSFNO"] is RefNoise
        th = mods["torch_harmonics"]
        before = (th.RealSHT, th.InverseRealSHT)
        ace_b200.patch_torch_harmonics()
        assert (th.RealSHT, th.InverseRealSHT) == before
        ref_step = mods["fme.core.step.single_module"].SingleModuleStepConfig
        StepSelector = mods["fme.core.step.step"].StepSelector
        names_before = {n: c for n, c in StepSelector.registry._types.items()}
        assert ace_b200.install_step_into_fme(override=True) is ref_step
        after = StepSelector.registry._types
        assert after["b200_single_module"] is ref_step and all(after[n] is c for n, c in names_before.items())


def _oracle_like(net):
    from oracle import sfno as osfno

    onet = osfno.SphericalFourierNeuralOperatorNet(IMG, len(IN_NAMES), len(OUT_NAMES), **NET).eval()
    onet.load_state_dict(net.state_dict())
    return onet


@pytest.mark.parametrize("residual", [False, True])
def test_install_step_into_fme_drives_the_fused_step(residual, monkeypatch):
    with fake_fme.installed() as mods:
        StepSelector = mods["fme.core.step.step"].StepSelector
        StepArgs = mods["fme.core.step.args"].StepArgs
        DatasetInfo = mods["fme.core.dataset_info"].DatasetInfo
        ace_b200.install_into_fme(override=True)
        cfg_cls = fme_step.install_step_into_fme(override=True)
        assert issubclass(cfg_cls, mods["fme.core.step.single_module"].SingleModuleStepConfig)
        config = dict(builder=dict(type="SphericalFourierNeuralOperatorNet", config=dict(NET)), in_names=IN_NAMES, out_names=OUT_NAMES,
                      normalization=dict(means=MEANS, stds=STDS), residual_prediction=residual)
        steps = {}
        for type_name in ("b200_single_module", "single_module"):  # the new name and the overridden reference name
            torch.manual_seed(0)
            steps[type_name] = StepSelector(type=type_name, config=config).get_step(DatasetInfo(img_shape=IMG), lambda mods_: None)
            assert isinstance(steps[type_name], cfg_cls.fused_step_class)
        step = steps["b200_single_module"]
        fused = step.fused_stepper
        assert fused.prognostic_names == ["c", "a", "b"] and fused.forcing_names == ["f1", "f2"]
        # the reference wraps the module (DummyWrapper / DDP, fme/core/step/single_module.py:345); the fused stepper holds the net itself
        assert isinstance(fused.module, ace_b200.SphericalFourierNeuralOperatorNet) and step.modules[0].module is fused.module
        assert step.config.residual_prediction is residual and step.eval() is step
        # the one library call replaced by the oracle network (CPU): everything else is the product's host code
        onet = _oracle_like(fused.module)

        def native(prog, forcing, ocean, corrector_next, noise, out, next_prog):
            state = {n: prog[:, i] for i, n in enumerate(fused.prognostic_names)}
            state.update({n: forcing[:, i] for i, n in enumerate(fused.forcing_names)})
            norm = {n: (state[n] - MEANS[n]) / STDS[n] for n in IN_NAMES}
            with torch.no_grad():
                y = onet(torch.stack([norm[n] for n in IN_NAMES], dim=1))
            for i, n in enumerate(OUT_NAMES):
                v = y[:, i] + (norm[n] if residual and n in IN_NAMES else 0.0)
                out[:, i] = v * STDS[n] + MEANS[n]
            for i, n in enumerate(fused.prognostic_names):
                next_prog[:, i] = out[:, OUT_NAMES.index(n)]

        monkeypatch.setattr(fused, "_native_step", native)
        g = torch.Generator().manual_seed(3)
        inp = {n: torch.randn(2, *IMG, generator=g) * STDS[n] + MEANS[n] for n in IN_NAMES}
        with torch.no_grad():
            res = step.step(StepArgs(input=inp, next_step_input_data={}))
        # the reference sequence (fake SingleModuleStep = normalise -> pack -> module -> unpack -> [residual] -> denormalise) with
        # the oracle network standing in for the module
        ref_cfg = mods["fme.core.step.single_module"].SingleModuleStepConfig.from_state(
            dict(config, builder=mods["fme.ace.registry.registry"].ModuleSelector(type="SphericalFourierNeuralOperatorNet", config=dict(NET))))
        ref_step = mods["fme.core.step.single_module"].SingleModuleStep.__new__(mods["fme.core.step.single_module"].SingleModuleStep)
        ref_step._config, ref_step._normalizer = ref_cfg, step.normalizer
        ref_step.in_names, ref_step.out_names = IN_NAMES, OUT_NAMES
        ref_step.module = lambda x, labels=None: onet(x)
        with torch.no_grad():
            ref = ref_step.step(StepArgs(input=inp, next_step_input_data={}))
        assert list(res.output) == OUT_NAMES
        for n in OUT_NAMES:
            torch.testing.assert_close(res.output[n], ref.output[n], rtol=1e-6, atol=1e-6)
        assert res.stepper_state is None  # nothing to carry without a corrector
        with pytest.raises(NotImplementedError):
            step.step(StepArgs(input=inp, next_step_input_data={}, labels=object()))
        # state dict round trip through the wrapped reference step
        st = step.get_state()
        step.load_state(st)


def test_fused_step_refuses_what_it_does_not_fuse():
    with fake_fme.installed() as mods:
        StepSelector = mods["fme.core.step.step"].StepSelector
        DatasetInfo = mods["fme.core.dataset_info"].DatasetInfo
        ace_b200.install_into_fme(override=True)
        fme_step.install_step_into_fme()
        base = dict(builder=dict(type="SphericalFourierNeuralOperatorNet", config=dict(NET)), in_names=IN_NAMES, out_names=OUT_NAMES,
                    normalization=dict(means=MEANS, stds=STDS))
        with pytest.raises(NotImplementedError):
            StepSelector(type="b200_single_module", config=dict(base, include_channel_mask_inputs=True)).get_step(DatasetInfo(img_shape=IMG))
        with pytest.raises(NotImplementedError):
            StepSelector(type="b200_single_module", config=dict(base, secondary_decoder=object())).get_step(DatasetInfo(img_shape=IMG))
        with pytest.raises(NotImplementedError):  # conservation corrector without vertical coordinate / area weights
            StepSelector(type="b200_single_module", config=dict(base, corrector=dict(conserve_dry_air=True))).get_step(DatasetInfo(img_shape=IMG))
        with pytest.raises(ValueError):
            StepSelector(type="b200_single_module", config=dict(base, unknown_field=1))
