"""Timing of the aggregator reductions (row f3) at the ACE2 output shape: 50 fields of 180x360, B = 1 and 8."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ace_b200.metrics as am

lat = torch.linspace(-89.5, 89.5, 180)
w = torch.cos(torch.deg2rad(lat))[:, None].expand(180, 360).contiguous()
ops = am.LatLonOperations(w)
sht = ops.get_real_sht()
for B in (1, 8):
    x, t = torch.randn(B, 50, 180, 360, device="cuda"), torch.randn(B, 50, 180, 360, device="cuda")
    res = {}
    for name, fn, nbytes in [("weighted_moments(bias,rmse,mean,std)", lambda: ops.area_weighted_statistics(x, t), 2 * x.numel() * 4),
                             ("zonal_mean", lambda: ops.zonal_mean(x), x.numel() * 4),
                             ("power_spectrum(incl. SHT)", lambda: am.spherical_power_spectrum(x, sht), x.numel() * 4)]:
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        res[name] = {"us": round(us, 1), "GBps": round(nbytes / us / 1e3, 1)}
    print(json.dumps({"B": B, "fields": 50, **res}))

# HEALPix SHT (row f4) at BASELINE configs[4]'s shape: nside 64, lmax = mmax = 127, batch 4 x 50 fields
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
    res[name] = {"us": round(us, 1), "GBps": round(nbytes / us / 1e3, 1)}
print(json.dumps({"healpix": "nside 64, 4 x 50 fields, lmax = mmax = 127", **res}))
