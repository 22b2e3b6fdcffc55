"""Shared helpers for the parity tests (golden loading, oracle net construction, tolerances)."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_state_dict(g):
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}


def golden_net_fields(g):
    if "fields" in g.files:
        return ast.literal_eval(str(g["fields"]))
    return dict(
        embed_dim=int(g["embed_dim"]), num_layers=int(g["num_layers"]),
        operator_type=str(g["operator_type"]), data_grid=str(g["data_grid"]),
    )


def oracle_net_from_golden(g):
    from oracle import sfno as osfno

    net = osfno.SphericalFourierNeuralOperatorNet(
        tuple(int(v) for v in g["img_shape"]), int(g["in_chans"]), int(g["out_chans"]), **golden_net_fields(g)
    ).eval()
    net.load_state_dict(golden_state_dict(g))
    return net


def field_rel_err(y, ref):
    """max over (batch, field) of max|y - ref| / max|ref|: the 'rtol on prognostic fields' figure."""
    y = y.double().flatten(2)
    ref = ref.double().flatten(2)
    return float(((y - ref).abs().amax(-1) / ref.abs().amax(-1).clamp_min(1e-30)).max())


NET_GOLDENS = [
    "ref_live_net_dhconv_9x18.npz",
    "ref_live_net_dhconv_16x32_eq.npz",
    "ref_live_net_diag_12x24.npz",
    "ref_live_net_dhconv_32x64_nonorm.npz",
    "ref_live_net_ace2like_48x96.npz",
    "ref_stored_sfnonet_output_is_unchanged.npz",
]
