"""A/B of the two ConditionalLayerNorm paths on the whole ERA5-baseline network (embed 512, 8 layers, 32 noise channels, 40 -> 54
channels, 180x360, B = 1): CUDA-graph replays, the two settings alternated to cancel clock / power drift.  One JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ace_b200

sel = ace_b200.ModuleSelector(type="B200NoiseConditionedSFNO", config=dict(embed_dim=512, num_layers=8, noise_embed_dim=32, noise_type="isotropic",
                                                                          affine_norms=True, normalize_big_skip=True))
m = sel.build(40, 54, ace_b200.DatasetInfo(img_shape=(180, 360))).torch_module.cuda().eval().requires_grad_(False)
x = torch.randn(1, 40, 180, 360, device="cuda")
graphs = {}
for opt in (0, 1):
    ace_b200.set_option("cln_gemm", opt)
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y = m(x)
    graphs[opt] = g
ace_b200.set_option("cln_gemm", 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {0: [], 1: []}
for rnd in range(6):
    for opt in (0, 1):
        g = graphs[opt]
        for _ in range(3):
            g.replay()
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[opt].append(round(e0.elapsed_time(e1) / 20, 3))
print(json.dumps({"ms_per_forward": {"streaming": res[0], "cln_gemm": res[1]},
                  "median": {"streaming": sorted(res[0])[3], "cln_gemm": sorted(res[1])[3]}}))
