"""CPU: the host-side logic of ``FusedStepper.predict`` / ``predict_generator`` / ``rollout`` (forcing windows, next-step forcing,
prescribed prognostic overwrite, feedback of the prognostic subset) against a loop restating the reference's
``Stepper.predict_generator`` (fme/ace/stepper/single_module.py:1124-1167) + ``step_with_adjustments``
(fme/core/step/single_module.py:648-714).

No kernel runs here: the one call into the library (``FusedStepper._native_step``) is replaced by the oracle network so that
only the Python plumbing is under test; tests/test_gpu_stepper.py runs the same scenario through the C ABI on the GPU.
"""
import pytest
import torch

import ace_b200
from oracle import sfno as osfno

IMG = (8, 16)
IN_NAMES = ["a", "b", "f1", "c", "f2"]
OUT_NAMES = ["c", "d1", "a", "b", "d2"]
NAMES = sorted(set(IN_NAMES + OUT_NAMES))
MEANS = {n: 0.1 * (i - 3) for i, n in enumerate(NAMES)}
STDS = {n: 0.5 + 0.25 * i for i, n in enumerate(NAMES)}


def _oracle_step(onet, residual, state):
    norm = {n: (state[n] - MEANS[n]) / STDS[n] for n in IN_NAMES}
    with torch.no_grad():
        y = onet(torch.stack([norm[n] for n in IN_NAMES], dim=1))
    out = {n: y[:, i] for i, n in enumerate(OUT_NAMES)}
    if residual:
        for n in OUT_NAMES:
            if n in IN_NAMES:
                out[n] = out[n] + norm[n]
    return {n: out[n] * STDS[n] + MEANS[n] for n in OUT_NAMES}


class _HostStepper(ace_b200.FusedStepper):
    """FusedStepper whose native step is the oracle chain (normalise -> net -> residual -> denormalise) on CPU tensors."""

    def __init__(self, onet, *args, **kw):
        super().__init__(*args, **kw)
        self._onet = onet

    def _native_step(self, prog, forcing, ocean, corrector_next, noise, out, next_prog):
        state = {n: prog[:, i] for i, n in enumerate(self.prognostic_names)}
        state.update({n: forcing[:, i] for i, n in enumerate(self.forcing_names)})
        res = _oracle_step(self._onet, self.residual_prediction, state)
        for i, n in enumerate(self.out_names):
            out[:, i] = res[n]
        for i, n in enumerate(self.prognostic_names):
            next_prog[:, i] = res[n]


def _make(**kw):
    torch.manual_seed(0)
    onet = osfno.SphericalFourierNeuralOperatorNet(IMG, len(IN_NAMES), len(OUT_NAMES), embed_dim=8, num_layers=1, operator_type="dhconv").eval()
    sel = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=dict(embed_dim=8, num_layers=1, operator_type="dhconv"))
    net = sel.build(len(IN_NAMES), len(OUT_NAMES), ace_b200.DatasetInfo(img_shape=IMG)).torch_module
    return onet, _HostStepper(onet, net, IN_NAMES, OUT_NAMES, MEANS, STDS, residual_prediction=True, **kw)


def _reference_loop(onet, ic, forcing, n_steps, next_step_forcing, prescribed):
    """predict_generator: inputs of step t = state + forcing[t] (forcing[t + 1] for next_step_forcing_names); after the step the
    prescribed names are overwritten from time t + 1; the prognostic subset of the output is the next state."""
    prog = [n for n in OUT_NAMES if n in IN_NAMES]
    state = {n: ic[n][:, 0] for n in prog}
    outs = []
    for t in range(n_steps):
        full = dict(state)
        for n in IN_NAMES:
            if n not in OUT_NAMES:
                full[n] = forcing[n][:, t + 1 if n in next_step_forcing else t]
        out = _oracle_step(onet, True, full)
        for n in prescribed:
            out[n] = forcing[n][:, t + 1]
        outs.append(out)
        state = {n: out[n] for n in prog}
    return outs


@pytest.mark.parametrize("next_step_forcing,prescribed", [((), ()), (("f2",), ()), (("f2",), ("b", "d1"))])
def test_predict_follows_the_reference_loop(next_step_forcing, prescribed):
    onet, st = _make(next_step_forcing_names=next_step_forcing, prescribed_prognostic_names=prescribed)
    assert st.prognostic_names == ["c", "a", "b"] and st.forcing_names == ["f1", "f2"]
    assert st.next_step_input_names == ["f1", "f2", *prescribed]
    T, B = 3, 2
    torch.manual_seed(1)
    ic = {n: torch.randn(B, 1, *IMG) for n in st.prognostic_names}
    forcing = {n: torch.randn(B, T + 1, *IMG) for n in st.next_step_input_names}
    ref = _reference_loop(onet, ic, forcing, T, next_step_forcing, prescribed)
    data, new_ic = st.predict(ic, forcing, use_cuda_graph=False)
    assert list(data.keys()) == OUT_NAMES and list(new_ic.keys()) == st.prognostic_names
    for n in OUT_NAMES:
        assert data[n].shape == (B, T, *IMG)
        for t in range(T):
            torch.testing.assert_close(data[n][:, t], ref[t][n], rtol=1e-6, atol=1e-6)
    for n in st.prognostic_names:
        assert new_ic[n].shape == (B, 1, *IMG)
        torch.testing.assert_close(new_ic[n][:, 0], ref[-1][n], rtol=1e-6, atol=1e-6)
    for n in prescribed:  # the overwrite is exact
        assert torch.equal(data[n], forcing[n][:, 1:])
    (pred, reference), new_ic2 = st.predict_paired(ic, forcing, use_cuda_graph=False)
    assert all(torch.equal(pred[n], data[n]) for n in OUT_NAMES) and all(torch.equal(new_ic2[n], new_ic[n]) for n in new_ic)
    assert list(reference) == list(forcing) and all(torch.equal(reference[n], forcing[n][:, 1:]) for n in forcing)
    # the generator yields the same steps; a second window continues from the returned state
    gen = list(st.predict_generator(ic, forcing, T, use_cuda_graph=False))
    assert len(gen) == T
    for t in range(T):
        for n in OUT_NAMES:
            assert torch.equal(gen[t][n], data[n][:, t])
    forcing2 = {n: torch.randn(B, 2, *IMG) for n in st.next_step_input_names}
    data2, _ = st.predict(new_ic, forcing2, use_cuda_graph=False)
    ref2 = _reference_loop(onet, {n: new_ic[n] for n in new_ic}, forcing2, 1, next_step_forcing, prescribed)
    for n in OUT_NAMES:
        torch.testing.assert_close(data2[n][:, 0], ref2[0][n], rtol=1e-6, atol=1e-6)


def test_step_and_argument_errors():
    onet, st = _make(prescribed_prognostic_names=("b",))
    torch.manual_seed(2)
    state = {n: torch.randn(2, *IMG) for n in IN_NAMES}
    with pytest.raises(ValueError, match="not in next_step_input_data"):
        st.step(state, {})
    nxt = {"b": torch.randn(2, *IMG)}
    out = st.step(state, nxt)
    assert torch.equal(out["b"], nxt["b"])
    ref = _oracle_step(onet, True, state)
    torch.testing.assert_close(out["a"], ref["a"], rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError, match="prescribed"):
        st.step_packed(torch.zeros(2, 3, *IMG), torch.zeros(2, 2, *IMG))
    with pytest.raises(ValueError, match="prescribed_seq"):
        st.rollout(torch.zeros(2, 3, *IMG), torch.zeros(2, 2, 2, *IMG), 2, use_cuda_graph=False)
    with pytest.raises(ValueError, match="times"):
        st.predict({n: torch.zeros(2, 1, *IMG) for n in st.prognostic_names},
                   {"f1": torch.zeros(2, 3, *IMG), "f2": torch.zeros(2, 2, *IMG), "b": torch.zeros(2, 3, *IMG)}, use_cuda_graph=False)
    with pytest.raises(ValueError, match="must be in out_names"):
        _make(prescribed_prognostic_names=("f1",))
    with pytest.raises(ValueError, match="not in in_names"):
        _make(next_step_forcing_names=("d1",))
    with pytest.raises(ValueError, match="is an output variable"):
        _make(next_step_forcing_names=("a",))
    # the real stepper refuses CPU tensors (no fallback)
    sel = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=dict(embed_dim=8, num_layers=1, operator_type="dhconv"))
    net = sel.build(len(IN_NAMES), len(OUT_NAMES), ace_b200.DatasetInfo(img_shape=IMG)).torch_module
    real = ace_b200.FusedStepper(net, IN_NAMES, OUT_NAMES, MEANS, STDS)
    with pytest.raises(ace_b200.AceError):
        real.step(state)
