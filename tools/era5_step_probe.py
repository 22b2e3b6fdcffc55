"""The complete ERA5-baseline step on the device (configs/baselines/era5/ace-train-config-1-step-pretrain.yaml:94-133 and its
in_names / out_names): NoiseConditionedSFNO (embed 512, 8 layers, 32 isotropic noise channels) inside the fused step with
ForcePositive, dry-air / moisture / energy correctors, frozen-precipitation clip and the prescribed-SST ocean; 40 inputs, 54 outputs,
180x360.  Times the CUDA-graph rollout (one replay per 6-hour step, fresh noise drawn inside the graph) with device-resident forcing
and the host-pipelined variant (forcing in / all outputs out through pinned memory every step).  Synthetic weights and data.

    python tools/era5_step_probe.py > gpurun_out/era5_step_probe.json
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ace_b200

NZ = 8
lev = lambda p: [f"{p}_{k}" for k in range(NZ)]  # noqa: E731
IN = (["land_fraction", "ocean_fraction", "sea_ice_fraction", "DSWRFtoa", "HGTsfc", "global_mean_co2", "PRESsfc", "surface_temperature"]
      + lev("air_temperature") + lev("specific_total_water") + lev("eastward_wind") + lev("northward_wind"))
OUT = (["PRESsfc", "surface_temperature"] + lev("air_temperature") + lev("specific_total_water") + lev("eastward_wind") + lev("northward_wind")
       + ["TMP2m", "Q2m", "UGRD10m", "VGRD10m", "LHTFLsfc", "SHTFLsfc", "PRATEsfc", "ULWRFsfc", "ULWRFtoa", "DLWRFsfc", "DSWRFsfc", "USWRFsfc",
          "USWRFtoa", "tendency_of_total_water_path_due_to_advection", "TMP850", "h500", "total_frozen_precipitation_rate", "PRMSL",
          "eastward_surface_stress", "northward_surface_stress"])
FORCE_POS = lev("specific_total_water") + ["Q2m", "PRATEsfc", "total_frozen_precipitation_rate", "ULWRFsfc", "ULWRFtoa", "DLWRFsfc", "DSWRFsfc",
                                           "USWRFsfc", "USWRFtoa"]
IMG = (180, 360)


def stats():
    m, s = {n: 0.0 for n in set(IN + OUT)}, {n: 1.0 for n in set(IN + OUT)}
    m["PRESsfc"], s["PRESsfc"] = 9.85e4, 1.0e3
    m["surface_temperature"], s["surface_temperature"] = 288.0, 12.0
    for k in range(NZ):
        m[f"air_temperature_{k}"], s[f"air_temperature_{k}"] = 215.0 + 9.0 * k, 6.0
        m[f"specific_total_water_{k}"], s[f"specific_total_water_{k}"] = 2e-6 * 4.0 ** k, 1e-6 * 4.0 ** k
        s[f"eastward_wind_{k}"] = s[f"northward_wind_{k}"] = 8.0
    for n, v in dict(LHTFLsfc=(85, 40), SHTFLsfc=(18, 25), PRATEsfc=(3e-5, 4e-5), ULWRFsfc=(395, 60), ULWRFtoa=(238, 35), DLWRFsfc=(340, 65),
                     DSWRFsfc=(185, 90), USWRFsfc=(25, 30), USWRFtoa=(100, 45), DSWRFtoa=(340, 120), HGTsfc=(380, 850),
                     total_frozen_precipitation_rate=(4e-6, 1e-5), tendency_of_total_water_path_due_to_advection=(0, 4e-5), TMP2m=(287, 14),
                     Q2m=(8e-3, 5e-3), TMP850=(280, 12), h500=(5600, 250), PRMSL=(1.011e5, 900)).items():
        m[n], s[n] = float(v[0]), float(v[1])
    return m, s


def main():
    dev = torch.device("cuda")
    torch.manual_seed(0)
    sel = ace_b200.ModuleSelector(type="B200NoiseConditionedSFNO", config=dict(embed_dim=512, num_layers=8, noise_embed_dim=32, noise_type="isotropic",
                                                                                affine_norms=True, normalize_big_skip=True))
    model = sel.build(len(IN), len(OUT), ace_b200.DatasetInfo(img_shape=IMG)).torch_module
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "W_scale" in k or "W_bias" in k:
                p.add_(0.05 * torch.randn_like(p))
            if k.endswith("decoder.2.weight"):
                p.mul_(0.05)  # small residual updates: the synthetic state stays physical over the rollout
    model = model.to(dev).eval().requires_grad_(False)
    means, stds = stats()
    lat = torch.linspace(-89.5, 89.5, IMG[0])
    w = torch.cos(torch.deg2rad(lat))[:, None].expand(*IMG).contiguous()
    ak = torch.tensor([0.0, 5000.0, 11000.0, 16000.0, 16500.0, 12000.0, 6000.0, 1500.0, 0.0], dtype=torch.float64)
    bk = torch.tensor([0.0, 0.0, 0.01, 0.08, 0.24, 0.48, 0.74, 0.93, 1.0], dtype=torch.float64)
    st = ace_b200.FusedStepper(
        model, IN, OUT, means, stds, residual_prediction=True, force_positive_names=FORCE_POS,
        ocean=dict(surface_temperature_name="surface_temperature", ocean_fraction_name="ocean_fraction", interpolate=False),
        corrector=dict(conserve_dry_air=True, moisture_budget_correction="advection_and_precipitation", clip_frozen_precipitation=True,
                       total_energy_budget_correction=dict(method="constant_temperature"), ak=ak, bk=bk, area_weights=w, timestep_seconds=21600.0))
    for B in (1, 2):
        T = 12
        g = torch.Generator().manual_seed(1)
        pm = torch.tensor([means[n] for n in st.prognostic_names])[None, :, None, None]
        ps = torch.tensor([stds[n] for n in st.prognostic_names])[None, :, None, None]
        prog0 = (pm + 0.3 * ps * torch.randn(B, len(st.prognostic_names), *IMG, generator=g))
        for i, n in enumerate(st.prognostic_names):
            if n.startswith("specific_total_water"):
                prog0[:, i].clamp_(min=0)
        prog0 = prog0.to(dev)
        fm = torch.tensor([means[n] for n in st.forcing_names])[None, None, :, None, None]
        fs = torch.tensor([stds[n] for n in st.forcing_names])[None, None, :, None, None]
        forcing = fm + 0.3 * fs * torch.randn(T + 1, B, len(st.forcing_names), *IMG, generator=g)
        forcing[:, :, st.forcing_names.index("ocean_fraction")] = torch.rand(T + 1, B, *IMG, generator=g)
        ocean = torch.stack([torch.rand(T, B, *IMG, generator=g), 288.0 + 10.0 * torch.randn(T, B, *IMG, generator=g)], dim=2)
        fdev, odev = forcing.to(dev), ocean.to(dev)
        st.rollout(prog0, fdev, 2, keep_outputs=False, ocean_seq=odev)  # capture + warm-up
        l0 = ace_b200.launch_count()
        st.rollout(prog0, fdev, 1, use_cuda_graph=False, keep_outputs=False, ocean_seq=odev)  # one eager step: count the launches
        launches = ace_b200.launch_count() - l0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        outs, fin = st.rollout(prog0, fdev, T, keep_outputs=True, ocean_seq=odev)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / T
        finite = bool(torch.isfinite(outs).all())
        # invariants on the last step's outputs: dry-air mass pinned, budgets closed (fp64 on the host)
        from oracle import corrector as oc
        from oracle import metrics as om
        vc = oc.VerticalCoordinate(ak.float(), bk.float())
        o = {n: outs[-1, :, i].cpu() for i, n in enumerate(OUT)}
        ic = {n: prog0[:, i].cpu() for i, n in enumerate(st.prognostic_names)}
        wat = lambda d: torch.stack([d[f"specific_total_water_{k}"] for k in range(NZ)], dim=-1)  # noqa: E731
        dry = lambda d: om.weighted_mean(oc.dry_air(d["PRESsfc"], wat(d), vc).double(), w.double())  # noqa: E731
        fh, oh = forcing.pin_memory(), ocean.pin_memory()
        out_host = torch.empty(T, B, len(OUT), *IMG).pin_memory()
        st.rollout_host(prog0, fh, 2, out_host[:2], ocean_host=oh)
        torch.cuda.synchronize()
        e0.record()
        st.rollout_host(prog0, fh, T, out_host, ocean_host=oh)
        e1.record()
        torch.cuda.synchronize()
        ms_host = e0.elapsed_time(e1) / T
        print(json.dumps({
            "workload": "ERA5 baseline step: NoiseConditionedSFNO 512x8 + ForcePositive + dry-air/moisture/energy correctors + frozen clip + prescribed-SST ocean, 40in/54out, 180x360",
            "B": B, "steps": T, "ms_per_step_device_resident": round(ms, 3), "sim_years_per_day": round(B * 6 / 24 / 365.25 / (ms / 1e3) * 86400, 1),
            "ms_per_step_host_io": round(ms_host, 3), "sim_years_per_day_host_io": round(B * 6 / 24 / 365.25 / (ms_host / 1e3) * 86400, 1),
            "h2d_bytes_per_step": int(fh[0].numel() * 4 * 2 + oh[0].numel() * 4), "d2h_bytes_per_step": int(out_host[0].numel() * 4),
            "kernel_launches_per_step": launches,
            "outputs_finite": finite, "dry_air_drift_Pa_after_%d_steps" % T: float((dry(o) - dry(ic)).abs().max()),
            "stochastic": bool(B == 1 or not torch.equal(outs[:, 0], outs[:, 1]))}), flush=True)


if __name__ == "__main__":
    main()
