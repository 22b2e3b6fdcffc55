"""Noise-conditioned SFNO (row f1) at the ERA5 baseline's size (configs/baselines/era5/ace-train-config-1-step-pretrain.yaml:94-108:
embed 512, 8 layers, 32 isotropic noise channels, affine_norms, normalize_big_skip; 40 inputs, 54 outputs, 180x360):
forward time (CUDA events over graph replays incl. the per-step noise draw), per-kernel profile, and a reduced-depth full-resolution
parity check against the oracle.  Prints JSON lines; run on the GPU box:  python tools/csfno_probe.py > gpurun_out/csfno_probe.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ace_b200
from ace_b200 import csfno as bc

IMG, CIN, COUT = (180, 360), 40, 54
CFG = dict(embed_dim=512, num_layers=8, noise_embed_dim=32, noise_type="isotropic", affine_norms=True, normalize_big_skip=True)


def build(layers, seed=0):
    torch.manual_seed(seed)
    sel = ace_b200.ModuleSelector(type="B200NoiseConditionedSFNO", config={**CFG, "num_layers": layers})
    m = sel.build(CIN, COUT, ace_b200.DatasetInfo(img_shape=IMG)).torch_module
    with torch.no_grad():
        for k, p in m.named_parameters():
            if "W_scale" in k or "W_bias" in k:
                p.add_(0.1 * torch.randn_like(p))
    return m


def main():
    dev = torch.device("cuda")
    # 1. parity at full resolution and width, 2 layers (the CPU oracle needs ~10 s per layer here)
    from oracle import csfno as oc
    from tests.util import field_rel_err
    m = build(2)
    onet = oc.SphericalFourierNeuralOperatorNet(IMG, CIN, COUT, oc.ContextConfig(embed_dim_noise=32), embed_dim=512, num_layers=2, affine_norms=True,
                                                normalize_big_skip=True, data_grid="legendre-gauss").eval()
    onet.load_state_dict(m.conditional_model.state_dict())
    x = torch.randn(1, CIN, *IMG)
    noise = oc.NoiseConditionedModel(onet, IMG, embed_dim_noise=32, isotropic=True).draw_noise(1)
    torch.set_num_threads(16)
    t0 = time.time()
    with torch.no_grad():
        ref = onet(x, oc.Context(noise=noise))
    t_cpu = time.time() - t0
    md = m.to(dev).eval().requires_grad_(False)
    u0, s0 = ace_b200.get_option("count_umma"), ace_b200.get_option("count_simt")
    out = md(x.to(dev), noise=noise.to(dev)).cpu()
    print(json.dumps({"parity_full_res_2_layers": {"max_field_rel_err": field_rel_err(out, ref), "tol": 1e-4, "oracle_cpu_s": round(t_cpu, 1),
                                                   "gemms_tcgen05": ace_b200.get_option("count_umma") - u0,
                                                   "gemms_simt": ace_b200.get_option("count_simt") - s0}}), flush=True)
    del md, m, onet
    # 2. timing, full depth
    for B in (1, 2):
        m = build(8).to(dev).eval().requires_grad_(False)
        x = torch.randn(B, CIN, *IMG, device=dev)
        y = m(x)  # warm-up: allocations, weight upload
        torch.cuda.synchronize()
        static_x = x.clone()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            static_y = m(static_x)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        ace_b200.set_option("profile", 1)
        ace_b200._lib.profile_report()
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        rep = ace_b200._lib.profile_report()
        ace_b200.set_option("profile", 0)
        kern = {k: {"n_per_step": c // 3, "ms_per_step": round(t / 3, 4)} for k, (c, t) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
        print(json.dumps({"B": B, "ms_per_step_graph": round(ms, 3), "sim_years_per_day": round(B * 6 / 24 / 365.25 / (ms / 1e3) * 86400, 1),
                          "profiled_sum_ms": round(sum(v["ms_per_step"] for v in kern.values()), 3), "kernels": kern}), flush=True)
        del m, g


if __name__ == "__main__":
    main()
