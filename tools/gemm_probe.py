"""Per-kernel CUDA-event times of one ACE2-size forward under different GEMM options (development probe)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ace_b200
from ace_b200 import _lib

IMG = tuple(int(v) for v in os.environ.get("ACE_PROBE_IMG", "180x360").split("x"))
def build(embed=384, layers=8, cin=44, cout=50):
    torch.manual_seed(0)
    fields = dict(embed_dim=embed, num_layers=layers, operator_type="dhconv", data_grid="legendre-gauss")
    net = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=fields).build(
        cin, cout, ace_b200.DatasetInfo(img_shape=IMG)).torch_module
    return net.cuda().eval().requires_grad_(False)

def run(net, x, n=3):
    with torch.no_grad():
        net(x); torch.cuda.synchronize()
        _lib.set_option("profile", 1)
        for _ in range(n): net(x)
        rep = _lib.profile_report()
        _lib.set_option("profile", 0)
    return {k: round(ms / c * 1e3, 1) for k, (c, ms) in rep.items()}

if __name__ == "__main__":
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    net = build(layers=layers)
    x = torch.randn(1, 44, *IMG, device="cuda")
    variants = [dict()]

    if len(sys.argv) > 2:
        variants = [json.loads(a) for a in sys.argv[2:]]
    base = dict(split_terms=3, umma_bn=0, umma_bk=0, dbg=0, conv_bn=0, pair=-1, dhconv_t=0, inv2=1, tile_list=1, l2_persist=0, group_order=1, trace=0, mma_batch=1, sp=1, sp_tma=1, bfly_pair=1, tile_serpentine=1, sp_tmx=1)
    for k, val in base.items(): _lib.set_option(k, val)
    with torch.no_grad():
        y_ref = net(x).clone()  # default options: the configuration the GPU test-suite validates against the oracle
    for v in variants:
        for k, val in {**base, **v}.items(): _lib.set_option(k, val)
        r = run(net, x)
        with torch.no_grad():
            y = net(x)
        diff = None if v.get("dbg") else float((y - y_ref).abs().max() / y_ref.abs().max())
        print(json.dumps({"opts": v, "us": r, "total_ms": round(sum(r.values())/1e3*1.0, 3), "max_rel_diff_vs_default": diff}), flush=True)
