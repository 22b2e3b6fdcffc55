"""Generate tests/golden/ref_stored_stepper_predict.npz from the reference's stored stepper-level golden (build container only).

TEST INFRASTRUCTURE.  ``fme/ace/stepper/testdata/stepper_predict_regression.pt`` is what the reference's own
``test_stepper_predict_regression`` (fme/ace/stepper/test_single_module.py:2403-2418) pins: ``Stepper.predict`` of a
``single_module`` step around ``SphericalFourierNeuralOperatorNet(embed_dim=16, num_layers=2)`` (builder defaults: diagonal
operator, InstanceNorm, Legendre-Gauss grid), in_names [a, b], out_names [b, c], every variable normalised with mean 0.1 / std 1.1,
3 samples, 2 forward steps on a 9x18 grid, all under ``torch.manual_seed(0)``: first the network is built (parameter draws), then
a, b, c ~ randn(3, 3, 9, 18) in that order (``get_regression_stepper_and_data``, :2293-2358).  ``fme`` cannot be imported here
(xarray etc.), so the scenario is replayed with this repository's builder (whose seeded initialisation equals the reference's,
tests/test_registry.py) and the oracle network; the replay must hit the STORED tensors before anything is written.

    python -m oracle.make_golden_stepper
"""
import inspect
import os

import numpy as np
import torch

from . import refload
from . import sfno as osfno

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
IN_NAMES, OUT_NAMES, IMG, MEAN, STD = ["a", "b"], ["b", "c"], (9, 18), 0.1, 1.1


def replay(state_dict, fields, a, b, n_steps):
    """Stepper.predict at the tensor level (fme/ace/stepper/single_module.py:1136-1167 around fme/core/step/single_module.py:648-665):
    normalise -> pack [a, b] -> net -> unpack [b, c] -> denormalise; b is fed back.  Returns ({name: [B, T, H, W]}, final b)."""
    ok = set(inspect.signature(osfno.SphericalFourierNeuralOperatorNet.__init__).parameters)
    net = osfno.SphericalFourierNeuralOperatorNet(IMG, len(IN_NAMES), len(OUT_NAMES), **{k: v for k, v in fields.items() if k in ok}).eval()
    net.load_state_dict(state_dict)
    state = b[:, 0]
    outs = {n: [] for n in OUT_NAMES}
    with torch.no_grad():
        for t in range(n_steps):
            y = net(torch.stack([(a[:, t] - MEAN) / STD, (state - MEAN) / STD], dim=1))
            for i, n in enumerate(OUT_NAMES):
                outs[n].append(y[:, i] * STD + MEAN)
            state = outs["b"][-1]
    return {n: torch.stack(v, dim=1) for n, v in outs.items()}, state


def main():
    import ace_b200

    stored = torch.load(os.path.join(refload.REFERENCE_ROOT, "fme", "ace", "stepper", "testdata", "stepper_predict_regression.pt"),
                        map_location="cpu", weights_only=False)
    torch.manual_seed(0)
    sel = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=dict(embed_dim=16, num_layers=2))
    mod = sel.build(len(IN_NAMES), len(OUT_NAMES), ace_b200.DatasetInfo(img_shape=IMG)).torch_module
    a, b, c = (torch.randn(3, 3, *IMG) for _ in range(3))
    outs, final = replay(mod.state_dict(), sel.config, a, b, 2)
    for n in OUT_NAMES:
        torch.testing.assert_close(outs[n], stored[f"output.{n}"].detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(final[:, None], stored["next_state.b"].detach(), rtol=1e-5, atol=1e-7)
    d = {f"sd.{k}": v.detach().numpy() for k, v in mod.state_dict().items()}
    d.update(a=a.numpy(), b=b.numpy(), c=c.numpy(), fields=np.asarray(repr(dict(sel.config))))
    d.update({k: v.detach().numpy() for k, v in stored.items()})
    np.savez(os.path.join(OUT, "ref_stored_stepper_predict.npz"), **d)
    print("wrote", os.path.join(OUT, "ref_stored_stepper_predict.npz"))


if __name__ == "__main__":
    main()
