"""HEALPix SHT (row f4) at BASELINE configs[4]'s shape (nside 64, lmax = mmax = 127, 4 x 50 fields): time per transform (CUDA events)
and the per-kernel split.  Run on the GPU box:  python tools/hpx_probe.py > gpurun_out/hpx_probe.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ace_b200

nside, lmax = 64, 127
fwd = ace_b200.HealpixSHT(nside, lmax=lmax, mmax=lmax, quad_weights="none")
inv = ace_b200.HealpixISHT(nside, lmax=lmax, mmax=lmax)
x = torch.randn(4, 50, 12 * nside**2, device="cuda")
c = fwd(x)
res = {}
for name, fn in [("healpix_sht", lambda: fwd(x)), ("healpix_isht", lambda: inv(c))]:
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    nbytes = x.numel() * 4 + c.numel() * 8
    ace_b200.set_option("profile", 1)
    ace_b200._lib.profile_report()
    for _ in range(3):
        fn()
    rep = ace_b200._lib.profile_report()
    ace_b200.set_option("profile", 0)
    res[name] = {"us": round(us, 1), "GBps": round(nbytes / us / 1e3, 1), "kernels_us": {k: round(t / n * 1e3, 1) for k, (n, t) in rep.items()}}
print(json.dumps({"healpix": "nside 64, 4 x 50 fields, lmax = mmax = 127", **res}))
