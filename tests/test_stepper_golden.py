"""The reference's STORED stepper-level golden (fme/ace/stepper/testdata/stepper_predict_regression.pt, pinned by the reference's
own test_stepper_predict_regression, fme/ace/stepper/test_single_module.py:2403-2418): Stepper.predict of a single_module step
around SphericalFourierNeuralOperatorNet(embed_dim=16, num_layers=2), 3 samples, 2 steps, 9x18.

CPU: (1) this repository's builder under torch.manual_seed(0) draws exactly the parameters that produced the stored tensors;
(2) the oracle step chain reproduces them; (3) FusedStepper.predict's window / feedback logic reproduces them with the oracle
standing in for the one library call.  GPU: FusedStepper.predict through the C ABI reproduces them.
"""
import os

import numpy as np
import pytest
import torch

import ace_b200
from oracle import make_golden_stepper as mgs
from tests.util import GOLDEN_DIR, field_rel_err, golden_net_fields

IN_NAMES, OUT_NAMES, IMG, MEAN, STD = mgs.IN_NAMES, mgs.OUT_NAMES, mgs.IMG, mgs.MEAN, mgs.STD


def _load():
    g = np.load(os.path.join(GOLDEN_DIR, "ref_stored_stepper_predict.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    return g, sd, t


def _b200_module(fields=None):
    sel = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=fields or dict(embed_dim=16, num_layers=2))
    return sel, sel.build(len(IN_NAMES), len(OUT_NAMES), ace_b200.DatasetInfo(img_shape=IMG)).torch_module


def test_seeded_builder_draws_the_parameters_behind_the_stored_golden():
    g, sd, t = _load()
    torch.manual_seed(0)
    sel, mod = _b200_module()
    a, b, c = (torch.randn(3, 3, *IMG) for _ in range(3))  # the reference draws its data right after building the network
    assert list(mod.state_dict().keys()) == list(sd.keys())
    for k, v in mod.state_dict().items():
        assert torch.equal(v, sd[k]), k
    assert torch.equal(a, t("a")) and torch.equal(b, t("b")) and torch.equal(c, t("c"))
    assert sel.config["operator_type"] == "diagonal" and sel.config["data_grid"] == "legendre-gauss"  # the builder's defaults


def test_oracle_chain_reproduces_the_stored_golden():
    g, sd, t = _load()
    outs, final = mgs.replay(sd, golden_net_fields(g), t("a"), t("b"), 2)
    for n in OUT_NAMES:
        torch.testing.assert_close(outs[n], t(f"output.{n}"), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(final[:, None], t("next_state.b"), rtol=1e-5, atol=1e-7)


def _stepper(mod, cls=ace_b200.FusedStepper, **kw):
    stats = {n: MEAN for n in ("a", "b", "c")}, {n: STD for n in ("a", "b", "c")}
    return cls(*kw.pop("lead", ()), mod, IN_NAMES, OUT_NAMES, *stats, residual_prediction=False, **kw)


def test_predict_host_logic_reproduces_the_stored_golden():
    from tests.test_stepper_host_logic import _HostStepper  # the native step replaced by an oracle chain

    g, sd, t = _load()
    _, mod = _b200_module()
    mod.load_state_dict(sd)
    import inspect

    from oracle import sfno as osfno

    ok = set(inspect.signature(osfno.SphericalFourierNeuralOperatorNet.__init__).parameters)
    onet = osfno.SphericalFourierNeuralOperatorNet(IMG, 2, 2, **{k: v for k, v in golden_net_fields(g).items() if k in ok}).eval()
    onet.load_state_dict(sd)

    class _Host(_HostStepper):
        def _native_step(self, prog, forcing, ocean, corrector_next, noise, out, next_prog):
            with torch.no_grad():
                y = onet(torch.stack([(forcing[:, 0] - MEAN) / STD, (prog[:, 0] - MEAN) / STD], dim=1))
            out.copy_(y * STD + MEAN)
            next_prog[:, 0] = out[:, 0]

    st = _stepper(mod, cls=_Host, lead=(onet,))
    assert st.prognostic_names == ["b"] and st.forcing_names == ["a"] and st.diagnostic_names == ["c"]
    data, new_ic = st.predict({"b": t("b")[:, :1]}, {"a": t("a")}, use_cuda_graph=False)
    for n in OUT_NAMES:
        torch.testing.assert_close(data[n], t(f"output.{n}"), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(new_ic["b"], t("next_state.b"), rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_fused_stepper_predict_reproduces_the_stored_golden_on_device():
    g, sd, t = _load()
    _, mod = _b200_module()
    mod.load_state_dict(sd)
    mod = mod.cuda().eval().requires_grad_(False)
    st = _stepper(mod)
    # eager launches: graph replay == eager is asserted bit-for-bit in tests/test_gpu_stepper.py
    data, new_ic = st.predict({"b": t("b")[:, :1].cuda()}, {"a": t("a").cuda()}, use_cuda_graph=False)
    for n in OUT_NAMES:
        # per (sample, time) relative error in NORMALISED units (the offset 0.1 would otherwise hide errors)
        got, ref = (data[n].cpu() - MEAN) / STD, (t(f"output.{n}") - MEAN) / STD
        assert field_rel_err(got, ref) < 1e-4, (n, field_rel_err(got, ref))
    got, ref = (new_ic["b"].cpu() - MEAN) / STD, (t("next_state.b") - MEAN) / STD
    assert field_rel_err(got, ref) < 1e-4
