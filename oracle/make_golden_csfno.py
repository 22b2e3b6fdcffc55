"""Generate tests/golden/ref_*csfno*.npz from the reference's conditional SFNO (run in the build container only).

TEST INFRASTRUCTURE.  Stored goldens of the reference (fme/core/models/conditional_sfno/testdata/*.pt) are converted to .npz
together with the exact parameters / inputs that reproduce them; live cases run the reference modules imported from
/root/reference (oracle/refload.py:load_csfno) on seeded inputs.  Parameters are stored under "p:<state_dict key>".

    python -m oracle.make_golden_csfno
"""
import os

import numpy as np
import torch

from . import refload

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TESTDATA = os.path.join(refload.REFERENCE_ROOT, "fme", "core", "models", "conditional_sfno", "testdata")


def _np(t):
    return t.detach().cpu().numpy()


def _pack(state, x, ctx, y, **meta):
    d = {f"p:{k}": _np(v) for k, v in state.items()}
    d["x"], d["y"] = _np(x), _np(y)
    for k, v in ctx.items():
        if v is not None:
            d[f"ctx:{k}"] = _np(v)
    for k, v in meta.items():
        d[f"meta:{k}"] = np.asarray(v)
    return d


def _reference_setup(r):
    """test_sfnonet.py:109-148 (setup_sfnonet) under manual_seed(0)."""
    torch.manual_seed(0)
    model = r.get_lat_lon_sfnonet(
        params=r.SFNONetConfig(embed_dim=16, num_layers=2, filter_type="linear"), img_shape=(9, 18), in_chans=2, out_chans=3,
        context_config=r.ContextConfig(embed_dim_scalar=8, embed_dim_labels=4, embed_dim_noise=16, embed_dim_pos=0))
    x = torch.randn(4, 2, 9, 18)
    ctx = dict(embedding_scalar=torch.randn(4, 8), labels=torch.randn(4, 4), noise=torch.randn(4, 16, 9, 18), embedding_pos=None)
    return model, x, ctx


def main(only=()):
    """``only``: names of live cases to (re)generate; empty = everything, including the conversions of the stored goldens."""
    r = refload.load_csfno()
    os.makedirs(OUT, exist_ok=True)
    if only == ("blocks",):
        return _block_cases(r)
    if only:
        return _live_cases(r, only)
    _stored_cases(r)
    _block_cases(r)
    _live_cases(r, only)
    _noise_case(r)
    print("wrote csfno goldens to", OUT)


def _stored_cases(r):
    # 1. test_sfnonet_output_is_unchanged.pt (test_sfnonet.py:151-159)
    model, x, ctx = _reference_setup(r)
    with torch.no_grad():
        y = model(x, r.Context(**ctx))
    stored = torch.load(os.path.join(TESTDATA, "test_sfnonet_output_is_unchanged.pt"), map_location="cpu")
    assert torch.allclose(y, stored, rtol=1e-5, atol=1e-6), float((y - stored).abs().max())
    np.savez(os.path.join(OUT, "ref_stored_csfno_output_is_unchanged.npz"),
             **_pack(model.state_dict(), x, ctx, stored, embed_dim=16, num_layers=2, data_grid="equiangular"))
    # 2. test_sfnonet_checkpoint_{input,output}.pt (test_sfnonet.py:181-192): parameters in the OLD dhconv layout
    ck = torch.load(os.path.join(TESTDATA, "test_sfnonet_checkpoint_input.pt"), map_location="cpu")
    x2, ctx2 = ck.pop("x"), ck.pop("context")
    stored2 = torch.load(os.path.join(TESTDATA, "test_sfnonet_checkpoint_output.pt"), map_location="cpu")
    model.load_state_dict(ck)
    with torch.no_grad():
        y2 = model(x2, r.Context(**ctx2))
    assert torch.allclose(y2, stored2, rtol=1e-5, atol=1e-6), float((y2 - stored2).abs().max())
    np.savez(os.path.join(OUT, "ref_stored_csfno_checkpoint.npz"), **_pack(ck, x2, ctx2, stored2, embed_dim=16, num_layers=2, data_grid="equiangular"))


def _block_cases(r):
    """fme/core/benchmark/testdata/csfno_block{,_8_groups}-regression.pt: ONE FourierNeuralOperatorBlock (no outer skip, default
    lobatto transforms, noise + labels + positional conditioning; 1 and 8 filter groups) as
    fme/core/models/conditional_sfno/benchmark.py:106-121 builds it under fme.core.rand.set_seed(0)
    (fme/core/benchmark/test_benchmark.py:44-60): numpy seed 1, random seed 2, torch seed 3."""
    import random

    base = refload.load()
    for name, groups in (("csfno_block", 1), ("csfno_block_8_groups", 8)):
        np.random.seed(1)
        random.seed(2)
        torch.manual_seed(3)
        B, C, H, L = 1, 16, 9, 18
        noise, labels, pos = torch.randn(B, 4, H, L), torch.randn(B, 3), torch.randn(B, 2, H, L)
        x = torch.randn(B, C, H, L)
        blk = r.FourierNeuralOperatorBlock(
            forward_transform=base.RealSHT(nlat=H, nlon=L), inverse_transform=base.InverseRealSHT(nlat=H, nlon=L), img_shape=(H, L),
            embed_dim=C, filter_type="linear", use_mlp=True, filter_num_groups=groups,
            context_config=r.ContextConfig(embed_dim_scalar=0, embed_dim_noise=4, embed_dim_labels=3, embed_dim_pos=2))
        ctx = dict(embedding_scalar=None, embedding_pos=pos, noise=noise, labels=labels)
        with torch.no_grad():
            y = blk(x, r.Context(**ctx))
        stored = torch.load(os.path.join(refload.REFERENCE_ROOT, "fme", "core", "benchmark", "testdata", f"{name}-regression.pt"),
                            map_location="cpu")["output"]
        assert torch.allclose(y, stored, rtol=1e-5, atol=1e-6), float((y - stored).abs().max())
        np.savez(os.path.join(OUT, f"ref_stored_{name}.npz"), **_pack(blk.state_dict(), x, ctx, stored, embed_dim=C, filter_num_groups=groups))


def _live_cases(r, only):
    # 3. live: the ERA5 baseline's option set at toy size (configs/baselines/era5/ace-train-config-1-step-pretrain.yaml:94-108:
    #    noise conditioning only, affine_norms, normalize_big_skip, legendre-gauss data grid), non-trivial conditioning weights
    for name, kw, cc, shape in [
        ("era5like_24x48", dict(embed_dim=32, num_layers=3, affine_norms=True, normalize_big_skip=True), dict(embed_dim_noise=8), (24, 48)),
        ("noaffine_pos_17x32", dict(embed_dim=24, num_layers=2, mlp_ratio=1.5), dict(embed_dim_noise=4, embed_dim_pos=2), (17, 32)),
        ("nonoise_12x24", dict(embed_dim=16, num_layers=2, affine_norms=True, big_skip=False, pos_embed=False), dict(), (12, 24)),
        # grouped filter + spectral / conv LoRA + bottlenecked spectral width + l = 0 pass-through, every parameter randomised
        ("grouped_lora_bottleneck_16x32", dict(embed_dim=32, num_layers=2, affine_norms=True, normalize_big_skip=True, filter_num_groups=2,
                                               spectral_ratio=0.5, spectral_lora_rank=2, spectral_lora_alpha=3.0, lora_rank=2,
                                               filter_preserves_global_mean=True), dict(embed_dim_noise=6, embed_dim_labels=2), (16, 32)),
        # SHT round trips of every residual, the big skip and the output, on the equiangular data grid
        ("filtered_20x40", dict(embed_dim=24, num_layers=2, affine_norms=True, normalize_big_skip=True, filter_num_groups=4,
                                filter_residual=True, filter_output=True), dict(embed_dim_noise=4, embed_dim_pos=3), (20, 40)),
    ]:
        if only and name not in only:
            continue
        torch.manual_seed(7)
        cfg = dict(embed_dim_scalar=0, embed_dim_labels=0, embed_dim_noise=0, embed_dim_pos=0)
        cfg.update(cc)
        grid = "legendre-gauss" if name.startswith(("era5", "grouped")) else "equiangular"
        m = r.get_lat_lon_sfnonet(params=r.SFNONetConfig(filter_type="linear", **kw), img_shape=shape, in_chans=5, out_chans=4,
                                  data_grid=grid, context_config=r.ContextConfig(**cfg))
        with torch.no_grad():
            for k, p in m.named_parameters():  # conditioning and affine parameters start at identity: randomise them
                if "W_scale" in k or "W_bias" in k or ".norm." in k or k.endswith("filter.filter.bias"):
                    p.add_(0.3 * torch.randn_like(p))
                elif "lora" in k:  # lora_B / lora_up start at zero
                    p.add_(0.2 * torch.randn_like(p))
        B = 2
        x = torch.randn(B, 5, *shape)
        ctx = dict(embedding_scalar=None, labels=torch.randn(B, cfg["embed_dim_labels"]) if cfg["embed_dim_labels"] else None,
                   noise=torch.randn(B, cfg["embed_dim_noise"], *shape) if cfg["embed_dim_noise"] else None,
                   embedding_pos=torch.randn(B, cfg["embed_dim_pos"], *shape) if cfg["embed_dim_pos"] else None)
        with torch.no_grad():
            y = m(x, r.Context(**ctx))
        np.savez(os.path.join(OUT, f"ref_live_csfno_{name}.npz"), **_pack(m.state_dict(), x, ctx, y, data_grid=grid, **kw))


def _noise_case(r):
    # 4. isotropic noise (stochastic_sfno.py:21-47) through the reference's InverseRealSHT, seeded draws stored
    torch.manual_seed(3)
    isht = r.InverseRealSHT(12, 24, lmax=12, mmax=13, grid="legendre-gauss")
    state = torch.get_rng_state()
    noise = r.isotropic_noise((2, 3), 12, 13, isht, torch.device("cpu"))
    torch.set_rng_state(state)
    real, imag = torch.randn(2, 3, 12, 13), torch.randn(2, 3, 12, 13)
    np.savez(os.path.join(OUT, "ref_live_isotropic_noise.npz"), real=_np(real), imag=_np(imag), noise=_np(noise))


if __name__ == "__main__":
    import sys

    main(tuple(sys.argv[1:]))
