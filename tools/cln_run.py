"""One forward of a 1-layer noise-conditioned SFNO at the ERA5 baseline's width / resolution (for ncu captures of cond_layer_norm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ace_b200
sel = ace_b200.ModuleSelector(type="B200NoiseConditionedSFNO", config=dict(embed_dim=512, num_layers=1, noise_embed_dim=32, noise_type="isotropic",
                                                                          affine_norms=True, normalize_big_skip=True))
m = sel.build(40, 54, ace_b200.DatasetInfo(img_shape=(180, 360))).torch_module.cuda().eval().requires_grad_(False)
x = torch.randn(1, 40, 180, 360, device="cuda")
for _ in range(2):
    y = m(x)
torch.cuda.synchronize()
print(float(y.abs().mean()))
