"""Shared helpers for the parity tests (golden loading, oracle net construction, tolerances)."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_DIR = GOLDEN


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


CSFNO_GOLDENS = [
    "ref_stored_csfno_output_is_unchanged.npz",
    "ref_stored_csfno_checkpoint.npz",
    "ref_live_csfno_era5like_24x48.npz",
    "ref_live_csfno_noaffine_pos_17x32.npz",
    "ref_live_csfno_nonoise_12x24.npz",
    "ref_live_csfno_grouped_lora_bottleneck_16x32.npz",
    "ref_live_csfno_filtered_20x40.npz",
]


def load_csfno_case(name):
    """A conditional-SFNO golden: returns (net kwargs, context dims, state_dict, x, context dict, y)."""
    import torch

    d = np.load(os.path.join(GOLDEN_DIR, name))
    state = {k[2:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("p:")}
    ctx = {k: None for k in ("embedding_scalar", "embedding_pos", "labels", "noise")}
    ctx.update({k[4:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("ctx:")})
    meta = {k[5:]: d[k].item() for k in d.files if k.startswith("meta:")}
    x, y = torch.from_numpy(d["x"]), torch.from_numpy(d["y"])
    dims = dict(embed_dim_scalar=0 if ctx["embedding_scalar"] is None else ctx["embedding_scalar"].shape[-1],
                embed_dim_labels=0 if ctx["labels"] is None else ctx["labels"].shape[-1],
                embed_dim_noise=0 if ctx["noise"] is None else ctx["noise"].shape[-3],
                embed_dim_pos=0 if ctx["embedding_pos"] is None else ctx["embedding_pos"].shape[-3])
    kwargs = dict(img_shape=tuple(x.shape[-2:]), in_chans=x.shape[1], out_chans=y.shape[1], **meta)
    return kwargs, dims, state, x, ctx, y
