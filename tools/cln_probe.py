"""ConditionalLayerNorm timing at the ERA5 baseline's width / resolution (C = 512, 180x360, 32 noise channels): per-kernel profile of a
2-layer noise-conditioned SFNO with the streaming kernel (option cln_gemm = 0) and the tensor-core path (default).  One JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ace_b200

sel = ace_b200.ModuleSelector(type="B200NoiseConditionedSFNO", config=dict(embed_dim=512, num_layers=2, noise_embed_dim=32, noise_type="isotropic",
                                                                          affine_norms=True, normalize_big_skip=True))
m = sel.build(40, 54, ace_b200.DatasetInfo(img_shape=(180, 360))).torch_module.cuda().eval().requires_grad_(False)
with torch.no_grad():
    for k, p in m.named_parameters():
        if "W_scale" in k or "W_bias" in k:
            p.add_(0.1 * torch.randn_like(p))
x = torch.randn(1, 40, 180, 360, device="cuda")
out = {}
ys = {}
for label, opt in (("streaming", 0), ("cln_gemm", 1)):
    ace_b200.set_option("cln_gemm", opt)
    torch.manual_seed(0)
    for _ in range(3):
        y = m(x)
    ace_b200.set_option("profile", 1)
    ace_b200._lib.profile_report()
    torch.manual_seed(0)
    for _ in range(5):
        y = m(x)
    rep = ace_b200._lib.profile_report()
    ace_b200.set_option("profile", 0)
    out[label] = {k: round(t / n * 1e3, 1) for k, (n, t) in rep.items() if k.startswith(("cond_layer", "cln"))}
    ys[label] = y
ace_b200.set_option("cln_gemm", 1)
out["paths_rel_diff"] = float((ys["streaming"] - ys["cln_gemm"]).abs().max() / ys["streaming"].abs().max())
print(json.dumps(out))
