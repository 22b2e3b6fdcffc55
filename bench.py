#!/usr/bin/env python
"""Benchmark of the ACE2 1-degree autoregressive rollout hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one 6-hour step of the ACE2 1-degree configuration (configs[1] of BASELINE.json:
180x360 grid, 44 input / 50 output channels, embed 384, 8 SFNO blocks, dhconv, instance norm;
38 prognostic variables fed back, 6 forcing-only inputs, 12 diagnostics) on synthetic data with
random-init weights: normalise -> pack -> SFNO -> unpack -> denormalise -> feed back.
Metric: simulated-years/day = steps/s * 86400 / 1460.

N > 1 (torchrun, one rank per GPU): every rank advances its own ensemble member(s) -- the rollout is
embarrassingly parallel -- and the only collective is one all_gather of the
per-member area-weighted global-mean diagnostics per 40-step window,
INSIDE the timed region; "scaling": "weak".

The timed region is repeated (--repeats, default 3; `ms_per_step` = median, all samples in `repeats`), a `sustained`
leg times one simulated year (1460 steps) with the SM clock sampled, `step_roofline` gives the whole-step fractions
of the measured HBM / bf16 peaks, and `gpu_eager_baseline` times the reference algorithm as PyTorch eager on the SAME
GPU (cuFFT + cuBLAS + ATen: the oracle port moved to CUDA, TF32 off and on) -- SURVEY.md section 2.3's bar.

--workload sht / inverse_sht: the reference's own micro-benchmark shapes (fme/sht_fix.py:232-327: 1024 fields of
180x360, default lobatto grid) through ace_b200.RealSHT / InverseRealSHT, reported as HBM GB/s against the roofline.
--workload csfno_block / csfno_block_8_groups: the reference's conditional-SFNO block micro-benchmarks
(fme/core/models/conditional_sfno/benchmark.py:29-46).
--workload healpix: BASELINE configs[4] as SURVEY.md maps it (HEALPix SHT pair, nside 64, 4 x 50 fields split over the ranks).
--workload quarter_degree: BASELINE configs[3] (721x1440, 44 in / 50 out, embed 384, 8 blocks) network forward on one GPU.

--impl reference times the reference algorithm's CPU path (the oracle port of the reference modules;
the reference is pure Python and `import fme` is impossible in this image) on the host
cores of the box, on a bounded number of steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STEPS_PER_YEAR = 1460  # 6-hourly
WORKLOAD = "ACE2 1deg rollout (BASELINE configs[1]): 180x360, 44in/50out, embed 384, 8 SFNO blocks (dhconv, instance_norm), 6h steps, B=1 per GPU"
IMG = (180, 360)
N_PROG, N_FORCING, N_DIAG = 38, 6, 12
EMBED, LAYERS = 384, 8
L_MODES, M_MODES = 180, 181


def names():
    prog = [f"p{i:02d}" for i in range(N_PROG)]
    forcing = [f"f{i}" for i in range(N_FORCING)]
    diag = [f"d{i:02d}" for i in range(N_DIAG)]
    in_names = forcing + prog           # 44 inputs
    out_names = prog + diag             # 50 outputs
    return in_names, out_names, prog, forcing, diag


def norm_stats(in_names, out_names):
    allnames = sorted(set(in_names) | set(out_names))
    means = {n: 0.05 * ((i % 7) - 3) for i, n in enumerate(allnames)}
    stds = {n: 1.0 + 0.1 * (i % 5) for i, n in enumerate(allnames)}
    return means, stds


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]),
                    bf16_tflops_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- algorithmic work
def algorithmic(B):
    """Per-launch algorithmic FLOPs / bytes of each kernel name (SURVEY.md section 8d, fp32 sizes)."""
    C, HW, L, M, K, W = EMBED, IMG[0] * IMG[1], L_MODES, M_MODES, IMG[0], IMG[1]
    act = B * C * HW * 4
    spec = B * C * L * M * 8
    leg = M * L * K * 4
    conv = lambda ci, co, extra_in=0: dict(  # noqa: E731
        flops=2.0 * B * HW * ci * co, bytes=B * HW * 4.0 * (ci + co) + ci * co * 4 + extra_in, bound="tensor")
    return {
        "encoder.0": conv(44, C), "encoder.2": conv(C, C, C * HW * 4),
        "inner_skip": conv(C, C, act), "mlp.fc1": conv(C, 2 * C), "mlp.fc2": conv(2 * C, C, act),
        "decoder.0": conv(C + 44, C), "decoder.2": conv(C, 50),
        "dhconv": dict(flops=8.0 * B * C * C * L * M, bytes=2.0 * spec + C * C * L * 8, bound="hbm"),
        # forward SHT = dft_fwd + legendre_fwd, inverse = legendre_inv + dft_inv; 9.01 GFLOP / 223.1 MB per transform
        "sht.dft_fwd": dict(flops=0.0, bytes=act + spec, bound="hbm"),
        "sht.legendre_fwd": dict(flops=4.0 * B * C * M * L * K, bytes=2.0 * spec + leg, bound="hbm"),
        "sht.legendre_inv": dict(flops=4.0 * B * C * M * L * K, bytes=2.0 * spec + leg, bound="hbm"),
        "sht.dft_inv": dict(flops=0.0, bytes=act + spec, bound="hbm"),
        "norm_split": dict(flops=0.0, bytes=2.0 * B * 44 * HW * 4, bound="hbm"),  # 44 input channels: fp32 read + split-plane write
    }


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture."""
    for name in ("r02_ncu_traffic.json", "ncu_traffic.json"):  # this round's capture first
        path = os.path.join(ROOT, "profiles", name)
        try:
            with open(path) as f:
                v = json.load(f).get(kernel)
            if v is not None:
                return v, name
        except Exception:  # noqa: BLE001
            continue
    return None, None


def sht_transform_bytes(B):
    C, HW, L, M, K = EMBED, IMG[0] * IMG[1], L_MODES, M_MODES, IMG[0]
    return B * C * (HW * 4 + L * M * 8) + M * L * K * 4


# --------------------------------------------------------------------------------------------- reference arm
def build_oracle_net():
    import torch

    from oracle import sfno as osfno

    torch.manual_seed(0)
    return osfno.SphericalFourierNeuralOperatorNet(IMG, 44, 50, embed_dim=EMBED, num_layers=LAYERS, operator_type="dhconv").eval()


def cpu_step_fn(onet, in_names, out_names, means, stds):
    """Reference algorithm of one step on CPU (oracle port): normalise, pack, net, unpack, denormalise."""
    import torch

    mi = torch.tensor([means[n] for n in in_names]).view(1, -1, 1, 1)
    si = torch.tensor([stds[n] for n in in_names]).view(1, -1, 1, 1)
    mo = torch.tensor([means[n] for n in out_names]).view(1, -1, 1, 1)
    so = torch.tensor([stds[n] for n in out_names]).view(1, -1, 1, 1)

    def step(x_in):
        with torch.no_grad():
            return onet((x_in - mi) / si) * so + mo

    return step


def log(msg):
    sys.stderr.write(f"[bench {time.strftime('%H:%M:%S')}] {msg}\n")
    sys.stderr.flush()


def pick_cpu_threads(onet):
    """Fastest torch thread count for this model on this host (one SFNO block as the probe).

    Using every logical CPU is not the fastest choice on a 100+ thread host (the per-degree bmm / 1x1 convs
    of one sample do not scale that far), and an honest CPU baseline is the best the host can do."""
    import torch

    cores = os.cpu_count() or 1
    cands = sorted({c for c in (cores, cores // 2, 64, 32, 16, 8) if 1 <= c <= cores}, reverse=True)
    x = torch.randn(1, EMBED, *IMG)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            onet.blocks[1](x)  # warm-up (thread pool, mkldnn primitives)
            t0 = time.perf_counter()
            onet.blocks[1](x)
            dt = time.perf_counter() - t0
        log(f"cpu baseline probe: {c} threads -> {dt:.2f} s per SFNO block")
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best, cores


def time_cpu_reference(max_seconds, max_steps, warmup=1):
    import torch

    in_names, out_names, prog, forcing, _ = names()
    means, stds = norm_stats(in_names, out_names)
    log("cpu baseline: building the oracle net (reference algorithm, torch CPU)")
    onet = build_oracle_net()
    threads, cores = pick_cpu_threads(onet)
    step = cpu_step_fn(onet, in_names, out_names, means, stds)
    torch.manual_seed(1)
    x = torch.randn(1, 44, *IMG)
    t0 = time.perf_counter()
    for _ in range(warmup):
        y = step(x)
    t_warm = (time.perf_counter() - t0) / max(warmup, 1)
    log(f"cpu baseline: warm-up step {t_warm:.2f} s with {threads} threads")
    n = int(max(1, min(max_steps, max_seconds // max(t_warm, 1e-3))))
    times = []
    for _ in range(n):
        t0 = time.perf_counter()
        y = step(x)
        x = torch.cat([x[:, :N_FORCING], y[:, :N_PROG]], dim=1)  # feed prognostic outputs back
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return dict(sec_per_step=sec, steps=n, cores=threads, host_cpus=cores, min_sec=min(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_cpu_reference(max_seconds=120.0, max_steps=max(1, args.steps), warmup=1)
    sypd = 86400.0 / r["sec_per_step"] / STEPS_PER_YEAR
    sample = (f"{r['steps']} full ACE2 1-degree steps (B=1) after 1 warm-up, {r['cores']} torch threads (fastest of a sweep; "
              f"host has {r['host_cpus']} logical CPUs), torch CPU fp32")
    line = {
        "impl": "reference", "metric": "simulated_years_per_day", "value": sypd, "unit": "sim-years/day",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": 1, "ms_per_step": r["sec_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "impl_note": "reference algorithm on the host CPUs (oracle port: same torch CPU ops as the fme modules)"},
        "cpu_baseline": {"value": sypd, "unit": "sim-years/day", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": sypd, "unit": "sim-years/day", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def time_gpu_eager(dev, in_names, out_names, means, stds, n_steps=5):
    """The reference algorithm as PyTorch eager on the same GPU: the oracle port's modules moved to CUDA, i.e. exactly the torch
    ops of the fme modules (torch.fft.rfft/irfft -> cuFFT, einsum / Conv2d -> cuBLAS / cuDNN, InstanceNorm2d, GELU -> ATen), with
    TF32 off (fp32 FFMA GEMMs: the parity reference's arithmetic) and on (TORCH_ALLOW_TF32_CUBLAS_OVERRIDE=1 is what the
    reference's own image runs, docker/Dockerfile:5).  A baseline measurement: none of this repository's kernels run here."""
    import torch

    log("gpu eager baseline: building the oracle net on the GPU")
    onet = build_oracle_net().to(dev)
    mi = torch.tensor([means[n] for n in in_names], device=dev).view(1, -1, 1, 1)
    si = torch.tensor([stds[n] for n in in_names], device=dev).view(1, -1, 1, 1)
    mo = torch.tensor([means[n] for n in out_names], device=dev).view(1, -1, 1, 1)
    so = torch.tensor([stds[n] for n in out_names], device=dev).view(1, -1, 1, 1)
    torch.manual_seed(1)
    x0 = torch.randn(1, 44, *IMG, device=dev)
    out = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for label, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            x = x0.clone()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.no_grad():
                for it in range(2 + n_steps):
                    if it == 2:
                        ev0.record()
                    y = onet((x - mi) / si) * so + mo
                    x = torch.cat([x[:, :N_FORCING], y[:, :N_PROG]], dim=1)
                ev1.record()
            torch.cuda.synchronize(dev)
            ms = ev0.elapsed_time(ev1) / n_steps
            out[label] = {"ms_per_step": ms, "value": 86400.0 / (ms * 1e-3) / STEPS_PER_YEAR, "unit": "sim-years/day"}
            log(f"gpu eager baseline ({label}): {ms:.2f} ms/step")
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    out["what"] = (f"reference algorithm (oracle port of the fme modules) as PyTorch eager on this GPU, B=1, {n_steps} steps after 2 warm-up, "
                   "CUDA events; cuFFT + cuBLAS/cuDNN + ATen kernels only")
    del onet
    torch.cuda.empty_cache()
    return out


def run_sht_workload(args):
    """The reference's `sht` / `inverse_sht` micro-benchmarks (fme/sht_fix.py:232-327: RealSHT(180, 360) -- default lobatto grid,
    lmax 179, mmax 181 -- on randn(1024, 180, 360); 10 iterations after 1 warm-up in the reference, CUDA events) through
    ace_b200.RealSHT / InverseRealSHT (C ABI ace_sht_forward / ace_sht_inverse), plus the same op as PyTorch eager on this GPU."""
    import torch

    import ace_b200
    from ace_b200 import _lib

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    nb, nlat, nlon = 1024, IMG[0], IMG[1]
    K, Wm = args.steps, max(args.warmup, 3)
    sht = ace_b200.RealSHT(nlat, nlon)
    isht = ace_b200.InverseRealSHT(nlat, nlon)
    L, M = sht.lmax, sht.mmax
    torch.manual_seed(0)
    x = torch.randn(nb, nlat, nlon, device=dev)
    xh = sht(x)
    inverse = args.workload == "inverse_sht"
    fn, arg = (isht, xh) if inverse else (sht, x)
    for _ in range(Wm):
        fn(arg)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    fn(arg)
    launches = _lib.launch_count() - l0
    sampler = ClockSampler(dev.index)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    samples = []
    for r in range(max(1, args.repeats)):
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(K):
            y = fn(arg)
        ev1.record()
        torch.cuda.synchronize()
        samples.append(ev0.elapsed_time(ev1) / K)
    clocks = sampler.stop()
    ms = sorted(samples)[len(samples) // 2]
    by = nb * (nlat * nlon * 4 + L * M * 8) + M * L * nlat * 4  # fields in + coefficients out + one table
    pk = peaks()
    # per-kernel split
    _lib.set_option("profile", 1)
    _lib.profile_report()
    for _ in range(3):
        fn(arg)
    rep = _lib.profile_report()
    _lib.set_option("profile", 0)
    kernels = {k: round(t / c * 1e3, 1) for k, (c, t) in rep.items()}
    # the same op as PyTorch eager on this GPU (oracle port moved to CUDA)
    from oracle import sht as osht

    o = osht.InverseRealSHT(nlat, nlon) if inverse else osht.RealSHT(nlat, nlon)
    eager = {}
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        for label, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                for _ in range(2):
                    o(arg)
                ev0.record()
                for _ in range(5):
                    o(arg)
                ev1.record()
            torch.cuda.synchronize()
            t = ev0.elapsed_time(ev1) / 5
            eager[label] = {"ms": t, "GBps": by / (t * 1e-3) / 1e9}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    # e2e: host fields in, host coefficients out
    xh_host = (xh.cpu() if inverse else x.cpu()).pin_memory()
    out_host = torch.empty_like(y, device="cpu").pin_memory()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(3):
        out_host.copy_(fn(xh_host.to(dev, non_blocking=True)), non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    e2e_ms = ev0.elapsed_time(ev1) / 3
    line = {
        "metric": "sht_hbm_gbps", "value": by / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": 1, "steps": K, "warmup": Wm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split-bf16 3-term products, fp32 accumulate; fp32 / complex64 I/O)", "data": "synthetic",
        "config": {"workload": f"reference micro-benchmark `{args.workload}` (fme/sht_fix.py:232-327): 1024 fields of 180x360, lobatto grid, "
                               f"lmax {L}, mmax {M}", "l2": "inputs larger than L2 (265 MB of fields per call)"},
        "e2e": {"value": by / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(xh_host.numel() * xh_host.element_size()),
                "d2h_bytes_per_step": int(out_host.numel() * out_host.element_size())},
        "gpu_launches": int(launches * K), "launches_per_step": int(launches), "clocks": clocks,
        "repeats": {"n": len(samples), "ms_per_step": [round(v, 4) for v in samples]},
        "roofline": {"bound": "hbm", "achieved": by / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": by / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None, "algorithmic_bytes": by,
                     "note": "whole transform (layout conversion + DFT + Legendre kernels); peak of " + pk["source"]},
        "kernels_us": kernels, "gpu_eager_baseline": eager, "cpu_baseline": None,
    }
    print(json.dumps(line), flush=True)


def run_quarter_degree(args):
    """BASELINE configs[3]: ACE 0.25 degree (721x1440, 44 in / 50 out, embed 384, 8 blocks, dhconv), one member on one B200,
    device-resident forward passes of the network (the step wrapper adds three elementwise kernels).  No CPU baseline leg: the
    reference algorithm needs minutes per Legendre table at L = 721 on the host."""
    import torch

    import ace_b200
    from ace_b200 import _lib

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    img = (721, 1440)
    fields = dict(embed_dim=384, num_layers=8, operator_type="dhconv", data_grid="legendre-gauss")
    torch.manual_seed(0)
    with torch.device(dev):
        net = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=fields).build(
            44, 50, ace_b200.DatasetInfo(img_shape=img)).torch_module
    net = net.eval().requires_grad_(False)
    x = torch.randn(1, 44, *img, device=dev)
    K, Wm = args.steps, max(args.warmup, 3)
    s0 = _lib.get_option("count_simt")
    with torch.no_grad():
        for _ in range(Wm):
            y = net(x)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        net(x)
        launches = _lib.launch_count() - l0
        sampler = ClockSampler(dev.index)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        samples = []
        for _ in range(max(1, args.repeats)):
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(K):
                y = net(x)
            ev1.record()
            torch.cuda.synchronize()
            samples.append(ev0.elapsed_time(ev1) / K)
        clocks = sampler.stop()
        _lib.set_option("profile", 1)
        _lib.profile_report()
        net(x)
        rep = _lib.profile_report()
        _lib.set_option("profile", 0)
        # e2e: host input in, host output out
        xh = x.cpu().pin_memory()
        yh = torch.empty_like(y, device="cpu").pin_memory()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(3):
            yh.copy_(net(xh.to(dev, non_blocking=True)), non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        e2e_ms = ev0.elapsed_time(ev1) / 3
    assert _lib.get_option("count_simt") == s0, "a GEMM fell back to the SIMT kernel"
    ms = sorted(samples)[len(samples) // 2]
    C, Kl, W, L, M = 384, img[0], img[1], img[0], img[0]
    sht_bytes = C * (Kl * W * 4 + L * M * 8) + M * L * Kl * 4
    kus = {k: round(t / c * 1e3, 1) for k, (c, t) in rep.items()}
    fwd, inv = (kus["sht.dft_fwd"] + kus["sht.legendre_fwd"]) * 1e-6, (kus["sht.dft_inv"] + kus["sht.legendre_inv"]) * 1e-6
    pk = peaks()
    syd = lambda t_ms: 86400.0 / (t_ms * 1e-3) / STEPS_PER_YEAR
    print(json.dumps({
        "metric": "simulated_years_per_day", "value": syd(ms), "unit": "sim-years/day", "n_gpus": 1, "steps": K, "warmup": Wm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split-bf16 3-term products, fp32 accumulate; fp32 I/O)", "data": "synthetic",
        "config": {"workload": "ACE 0.25deg (BASELINE configs[3]): 721x1440, 44in/50out, embed 384, 8 SFNO blocks (dhconv, instance_norm), B=1, network forward",
                   "weights": "random init (reference initialisation, seed 0)", "l2": "inputs larger than L2 (1.6 GB per activation tensor)"},
        "e2e": {"value": syd(e2e_ms), "unit": "sim-years/day", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(xh.numel() * 4), "d2h_bytes_per_step": int(yh.numel() * 4)},
        "gpu_launches": int(launches * K), "launches_per_step": int(launches), "clocks": clocks,
        "repeats": {"n": len(samples), "ms_per_step": [round(v, 3) for v in samples]},
        "roofline": {"kernel": "sht (forward / inverse)", "bound": "hbm", "achieved": sht_bytes / max(fwd, inv) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": sht_bytes / max(fwd, inv) / 1e9 / pk["hbm_gbs"], "traffic": None, "algorithmic_bytes": sht_bytes,
                     "forward_us": fwd * 1e6, "inverse_us": inv * 1e6, "note": "slower of the two transforms; peak of " + pk["source"]},
        "kernels_us": kus, "mem_allocated_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1), "outputs_finite": bool(torch.isfinite(y).all()),
        "cpu_baseline": None,
    }), flush=True)


def run_csfno_block(args):
    """The reference's `csfno_block` / `csfno_block_8_groups` micro-benchmarks (fme/core/models/conditional_sfno/benchmark.py:29-46:
    one conditional FourierNeuralOperatorBlock, B = 2, C = 512, 180x360, 64 noise + 32 positional context channels + 3 labels,
    filter_num_groups 1 or 8).  The C ABI exposes networks, not blocks: the block runs as the single block of a network with an
    8-channel encoder / decoder, and `value` is the block's own kernels (norm0, SHT, dhconv, inverse SHT, inner_skip + GELU, norm1,
    MLP, plus the per-forward context preparation) summed from the per-kernel CUDA-event profile; the whole forward is reported too."""
    import torch

    import ace_b200
    from ace_b200 import _lib
    from ace_b200 import csfno as bc

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    G = 8 if args.workload.endswith("8_groups") else 1
    B, C, img, cio = 2, 512, IMG, 8
    dims = dict(embed_dim_noise=64, embed_dim_pos=32, embed_dim_labels=3)
    torch.manual_seed(0)
    net = bc.get_lat_lon_sfnonet(bc.SFNONetConfig(embed_dim=C, num_layers=1, filter_num_groups=G), in_chans=cio, out_chans=cio, img_shape=img,
                                 data_grid="legendre-gauss", context_config=bc.ContextConfig(**dims)).to(dev).eval().requires_grad_(False)
    with torch.no_grad():
        for k, p in net.named_parameters():
            if "W_scale" in k or "W_bias" in k:
                p.add_(0.1 * torch.randn_like(p))
    x = torch.randn(B, cio, *img, device=dev)
    ctx = bc.Context(noise=torch.randn(B, 64, *img, device=dev), embedding_pos=torch.randn(B, 32, *img, device=dev),
                     labels=torch.randn(B, 3, device=dev))
    K, Wm = args.steps, max(args.warmup, 3)
    for _ in range(Wm):
        y = net(x, ctx)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    net(x, ctx)
    launches = _lib.launch_count() - l0
    sampler = ClockSampler(dev.index)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    samples = []
    for _ in range(max(1, args.repeats)):
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(K):
            y = net(x, ctx)
        ev1.record()
        torch.cuda.synchronize()
        samples.append(ev0.elapsed_time(ev1) / K)
    clocks = sampler.stop()
    _lib.set_option("profile", 1)
    _lib.profile_report()
    for _ in range(3):
        net(x, ctx)
    rep = _lib.profile_report()
    _lib.set_option("profile", 0)
    kus = {k: {"n": c // 3, "us": round(t / c * 1e3, 1)} for k, (c, t) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
    outside = ("encoder", "decoder", "norm_split", "split_input", "big_skip")
    block_ms = sum(t for k, (c, t) in rep.items() if not k.startswith(outside)) / 3
    # e2e: host tensors in (input + per-pixel context), host output out
    host = [t.cpu().pin_memory() for t in (x, ctx.noise, ctx.embedding_pos, ctx.labels)]
    yh = torch.empty_like(y, device="cpu").pin_memory()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(3):
        xd, nd, pd, ld = (t.to(dev, non_blocking=True) for t in host)
        yh.copy_(net(xd, bc.Context(noise=nd, embedding_pos=pd, labels=ld)), non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    e2e_ms = ev0.elapsed_time(ev1) / 3
    # the reference block (oracle port of the fme modules) as PyTorch eager on this GPU
    eager = None
    try:
        from oracle import csfno as oc
        from oracle import sht as osht

        blk = oc.FourierNeuralOperatorBlock(osht.RealSHT(*img), osht.InverseRealSHT(*img), C, img, oc.ContextConfig(**dims), filter_num_groups=G,
                                            outer_skip=None).to(dev).eval()
        xb = torch.randn(B, C, *img, device=dev)
        octx = oc.Context(noise=ctx.noise, embedding_pos=ctx.embedding_pos, labels=ctx.labels)
        eager = {}
        old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
        try:
            for label, tf32 in (("fp32", False), ("tf32", True)):
                torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = tf32
                with torch.no_grad():
                    for _ in range(2):
                        blk(xb, octx)
                    torch.cuda.synchronize()
                    ev0.record()
                    for _ in range(5):
                        blk(xb, octx)
                    ev1.record()
                torch.cuda.synchronize()
                eager[label] = {"ms": ev0.elapsed_time(ev1) / 5}
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    except Exception as e:  # the eager leg is a reported baseline, not the measurement
        eager = {"error": repr(e)[:200]}
    ms = sorted(samples)[len(samples) // 2]
    print(json.dumps({
        "metric": "csfno_block_ms", "value": block_ms, "unit": "ms", "n_gpus": 1, "steps": K, "warmup": Wm, "ms_per_step": ms,
        "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split-bf16 3-term products, fp32 accumulate; fp32 I/O)", "data": "synthetic",
        "config": {"workload": f"reference micro-benchmark `{args.workload}` (fme/core/models/conditional_sfno/benchmark.py:29-46): one conditional "
                               f"SFNO block, B=2, C=512, 180x360, context 64 noise + 32 positional + 3 labels, filter_num_groups={G}",
                   "value_is": "sum of the block's kernels (CUDA-event profile); ms_per_step = the whole one-block network forward",
                   "l2": "inputs larger than L2 (265 MB per activation tensor)"},
        "e2e": {"value": e2e_ms, "unit": "ms", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in host)),
                "d2h_bytes_per_step": int(yh.numel() * 4), "path": "one-block network forward, host tensors in / out"},
        "gpu_launches": int(launches * K), "launches_per_step": int(launches), "clocks": clocks,
        "repeats": {"n": len(samples), "ms_per_step": [round(v, 4) for v in samples]},
        "kernels_us": kus, "gpu_eager_baseline": eager, "roofline": None, "cpu_baseline": None,
        "outputs_finite": bool(torch.isfinite(y).all()),
    }), flush=True)


def run_healpix(args):
    """BASELINE configs[4] as SURVEY.md section 0 / 8(d) maps it: the HEALPix SHT pair of fme/core/cuhpx/sht.py:32-153 at nside 64
    (lmax = mmax = 127, 49 152 ring-ordered pixels), input randn(4, 50, npix) (seed 0), the 4 samples split over the ranks
    (1 sample per GPU on 4 GPUs; every rank keeps at least one).  A step = one forward + one inverse transform of the rank's fields
    through ace_b200.HealpixSHT / HealpixISHT (C ABI ace_hpx_forward / ace_hpx_inverse)."""
    import torch

    import ace_b200
    from ace_b200 import _lib, parallel

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    rank, world, local_rank = parallel.init_from_env(backend="nccl", device=dev)
    nside, lmax, nch, nsamp = 64, 127, 50, 4
    npix = 12 * nside**2
    per_rank = max(1, nsamp // world)
    K, Wm = args.steps, max(args.warmup, 3)
    fwd = ace_b200.HealpixSHT(nside, lmax=lmax, mmax=lmax, quad_weights="none")
    inv = ace_b200.HealpixISHT(nside, lmax=lmax, mmax=lmax)
    torch.manual_seed(0)
    x_all = torch.randn(nsamp, nch, npix)
    x = x_all[(rank * per_rank) % nsamp:][:per_rank].to(dev)

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(Wm):
        c = fwd(x)
        y = inv(c)
    torch.cuda.synchronize(dev)
    l0 = _lib.launch_count()
    inv(fwd(x))
    launches = _lib.launch_count() - l0
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    samples = []
    for _ in range(max(1, args.repeats)):
        barrier()
        ev0.record()
        for _ in range(K):
            y = inv(fwd(x))
        ev1.record()
        barrier()
        samples.append(parallel.max_over_ranks(ev0.elapsed_time(ev1), device=dev) / K)
    clocks = sampler.stop()
    ms = sorted(samples)[len(samples) // 2]
    # end to end: host pixels in, host pixels out
    xh = x.cpu().pin_memory()
    yh = torch.empty_like(xh).pin_memory()
    barrier()
    ev0.record()
    for _ in range(5):
        yh.copy_(inv(fwd(xh.to(dev, non_blocking=True))), non_blocking=True)
    ev1.record()
    barrier()
    e2e_ms = parallel.max_over_ranks(ev0.elapsed_time(ev1), device=dev) / 5
    if rank != 0:
        return
    _lib.set_option("profile", 1)
    _lib.profile_report()
    for _ in range(3):
        inv(fwd(x))
    rep = _lib.profile_report()
    _lib.set_option("profile", 0)
    kus = {k: round(t / n * 1e3, 1) for k, (n, t) in rep.items()}
    nf = per_rank * nch
    L = M = lmax
    T = 4 * nside - 1
    by = 2 * (nf * (npix * 4 + L * M * 8) + M * L * T * 4)  # per rank and step: pixels + coefficients + one Legendre table, both directions
    gbps = lambda t_ms: world * by / (t_ms * 1e-3) / 1e9
    pk = peaks()
    # reference algorithm on the host (oracle port of the cuhpx classes: a Python loop of 255 torch.fft calls per direction)
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import healpix as oh

        o_f, o_i = oh.SHT(nside, lmax, lmax, oh.uniform_weights(nside)), oh.iSHT(nside, lmax, lmax)
        xs = x_all[:1]
        o_i(o_f(xs))
        t0 = time.time()
        o_i(o_f(xs))
        dt = time.time() - t0
        cpu = {"value": (2 * (nch * (npix * 4 + L * M * 8) + M * L * T * 4)) / dt / 1e9, "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"one forward + inverse pair of 1 x {nch} fields, oracle port of fme/core/cuhpx (torch CPU), {dt * 1e3:.0f} ms"}
    print(json.dumps({
        "metric": "healpix_sht_hbm_gbps", "value": gbps(ms), "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong" if world <= nsamp else "weak", "vs_baseline": None,
        "dtype": "bf16x3 Legendre stage (split-bf16, fp32 accumulate), fp32 ring DFT; fp32 / complex64 I/O", "data": "synthetic",
        "config": {"workload": f"HEALPix SHT pair (BASELINE configs[4] per SURVEY section 0): nside {nside}, lmax = mmax = {lmax}, {nsamp} x {nch} fields "
                               f"split {per_rank} sample(s) per GPU; step = forward + inverse transform", "l2": "no flush: 39 MB of pixels + 26 MB of "
                               "coefficients per rank fit the L2; the transform is compute / latency bound (see kernels_us)"},
        "e2e": {"value": gbps(e2e_ms), "unit": "GB/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(xh.numel() * 4), "d2h_bytes_per_step": int(yh.numel() * 4)},
        "gpu_launches": int(launches * K), "launches_per_step": int(launches), "clocks": clocks,
        "repeats": {"n": len(samples), "ms_per_step": [round(v, 4) for v in samples]},
        "roofline": {"bound": "hbm", "achieved": by / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": by / (ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                     "traffic": None, "algorithmic_bytes": by, "note": "per GPU, whole pair (8 kernels); peak of " + pk["source"]},
        "kernels_us": kus, "cpu_baseline": cpu, "outputs_finite": bool(torch.isfinite(y).all()),
    }), flush=True)


# --------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import ace_b200
    from ace_b200 import _lib

    from ace_b200 import parallel

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    rank, world, local_rank = parallel.init_from_env(backend="nccl", device=dev)
    B = args.batch
    K, Wm = args.steps, max(args.warmup, 3)
    H, Wd = IMG
    HW = H * Wd

    log(f"rank {rank}/{world}: building the B200 net")
    in_names, out_names, prog, forcing, diag = names()
    means, stds = norm_stats(in_names, out_names)
    fields = dict(embed_dim=EMBED, num_layers=LAYERS, operator_type="dhconv", data_grid="legendre-gauss")
    torch.manual_seed(0)
    net = ace_b200.ModuleSelector(type="B200SphericalFourierNeuralOperatorNet", config=fields).build(
        len(in_names), len(out_names), ace_b200.DatasetInfo(img_shape=IMG)).torch_module
    net = net.to(dev).eval().requires_grad_(False)
    stepper = ace_b200.FusedStepper(net, in_names, out_names, means, stds, residual_prediction=False)

    # ensemble sharding: this rank advances members member_slice(world * B) of the global ensemble
    msl = parallel.member_slice(world * B, rank, world)
    g = torch.Generator().manual_seed(1 + msl.start)
    prog0 = (torch.randn(B, N_PROG, H, Wd, generator=g)).to(dev)
    n_forc_steps = min(K, 64)  # forcing window resident on device, cycled
    forcing_dev = torch.randn(n_forc_steps, B, N_FORCING, H, Wd, generator=g).to(dev)

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize(dev)

    # ---- eager warm-up: allocations, parameter upload, launch count per step
    log("eager warm-up step (parameter upload, workspaces)")
    l0 = _lib.launch_count()
    out_buf, nxt = stepper.step_packed(prog0, forcing_dev[0])
    torch.cuda.synchronize(dev)
    l1 = _lib.launch_count()
    stepper.step_packed(prog0, forcing_dev[0], out_buf, nxt)
    torch.cuda.synchronize(dev)
    launches_per_step = _lib.launch_count() - l1
    del l0
    diag_launches_per_window = 1  # ace_weighted_moments (the all_gather is NCCL's kernel, not counted)

    # ---- device-resident timed region (CUDA graph replay per step, forcing already in HBM)
    log("graph capture + warm-up")
    st = stepper
    st.rollout(prog0, forcing_dev, min(Wm, n_forc_steps), use_cuda_graph=True, keep_outputs=False)  # capture + warm-up
    static = st._static
    graph = st._graph
    static["prog"].copy_(prog0)
    for t in range(Wm):
        static["forcing"].copy_(forcing_dev[t % n_forc_steps])
        graph.replay()
    # per-window diagnostics: area-weighted global means of every output field of the window's
    # last step (one ace_weighted_moments launch) and ONE all_gather of [B_local, n_out] across the ranks, every WINDOW steps
    from ace_b200 import legendre as _leg
    from ace_b200.metrics import LatLonOperations

    _, wq, _ = _leg.grid_nodes("legendre-gauss", H)
    area = torch.as_tensor(wq, dtype=torch.float32)[:, None].expand(H, Wd).contiguous()
    ops = LatLonOperations(area)
    WINDOW = 40
    diag = {"n": 0, "last": None}

    def window_diagnostics():
        gm_local = ops.area_weighted_mean(static["out"])       # [B, n_out], one reduction kernel
        diag["last"] = parallel.gather_members(gm_local)        # the rollout's only collective
        diag["n"] += 1

    def timed_steps(n):
        """n steps on the current stream: forcing copy + CUDA-graph replay per step, diagnostics gather per window."""
        for t in range(n):
            static["forcing"].copy_(forcing_dev[t % n_forc_steps])
            graph.replay()
            if (t + 1) % WINDOW == 0 or t + 1 == n:
                window_diagnostics()

    window_diagnostics()  # warm-up of the reduction kernel / NCCL communicator, outside every timed region
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    samples = []
    log(f"timed region: {args.repeats} x {K} steps")
    sampler.start()
    for r in range(max(1, args.repeats)):
        barrier()
        ev0.record()
        timed_steps(K)
        ev1.record()
        barrier()
        samples.append(parallel.max_over_ranks(ev0.elapsed_time(ev1), device=dev) / K)
    clocks = sampler.stop()
    ms_sorted = sorted(samples)
    ms_per_step = ms_sorted[len(ms_sorted) // 2]  # median of the repeats; every repeat times exactly K steps
    steps_per_s = world * B * 1e3 / ms_per_step  # every rank advances B members per step
    value = steps_per_s * 86400.0 / STEPS_PER_YEAR
    gm = diag["last"]
    finite = bool(torch.isfinite(gm).all().item()) and gm.shape[0] == world * B

    # ---- sustained leg: one simulated year (1460 steps) back to back, the clock the power cap settles at
    sustained = None
    if args.sustained_steps > 0:
        log(f"sustained leg: {args.sustained_steps} steps")
        s2 = ClockSampler(local_rank)
        barrier()
        s2.start()
        ev0.record()
        timed_steps(args.sustained_steps)
        ev1.record()
        barrier()
        c2 = s2.stop()
        ms_s = parallel.max_over_ranks(ev0.elapsed_time(ev1), device=dev) / args.sustained_steps
        sustained = {"steps": args.sustained_steps, "ms_per_step": ms_s, "value": world * B * 1e3 / ms_s * 86400.0 / STEPS_PER_YEAR,
                     "unit": "sim-years/day", "clocks": c2, "windows_gathered": args.sustained_steps // WINDOW}

    log(f"device-resident: {ms_per_step:.3f} ms/step; end-to-end loop")
    # ---- end-to-end through the public API with HOST buffers: H2D forcing + D2H outputs every step
    forcing_host = torch.randn(n_forc_steps, B, N_FORCING, H, Wd, generator=g).pin_memory()
    n_e2e = min(K, 64)  # pinned output window (cycled for longer runs)
    out_host = torch.empty(n_e2e, B, len(out_names), H, Wd).pin_memory()
    state = prog0.clone()

    def e2e_run(n):
        nonlocal state
        done = 0
        while done < n:
            m = min(n_e2e, n - done)
            state = stepper.rollout_host(state, forcing_host, m, out_host[:m])
            done += m

    e2e_run(Wm)
    e2e_samples = []
    for r in range(max(1, args.repeats)):
        barrier()
        ev0.record()
        e2e_run(K)
        ev1.record()
        barrier()
        e2e_samples.append(parallel.max_over_ranks(ev0.elapsed_time(ev1), device=dev) / K)
    e2e_ms_per_step = sorted(e2e_samples)[len(e2e_samples) // 2]
    e2e_value = world * B * 1e3 / e2e_ms_per_step * 86400.0 / STEPS_PER_YEAR
    h2d = B * N_FORCING * HW * 4
    d2h = B * len(out_names) * HW * 4

    # ---- per-kernel CUDA-event timing (eager, same steps) for the roofline of the dominant kernel
    roofline, roofline_sht, shares, kernels = None, None, None, None
    log("per-kernel event timing")
    if rank == 0:
        pk = peaks()
        _lib.set_option("profile", 1)
        n_prof = min(K, 10)
        for t in range(n_prof):
            stepper.step_packed(prog0, forcing_dev[t % n_forc_steps], out_buf, nxt)
        rep = _lib.profile_report()
        _lib.set_option("profile", 0)
        total_ms = sum(ms for _, ms in rep.values())
        shares = {k: round(ms / total_ms, 4) for k, (_, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])}
        alg = algorithmic(B)
        kernels = {}
        for name, (cnt_k, ms_k) in rep.items():
            a_k = alg.get(name)
            us_k = ms_k / cnt_k * 1e3
            ent = {"us": round(us_k, 1), "launches_per_step": cnt_k // n_prof}
            if a_k is not None:
                if a_k["bound"] == "tensor":
                    ent.update(bound="tensor", frac=round(a_k["flops"] / (us_k * 1e-6) / 1e12 / pk["bf16_tflops_sustained"], 3))
                else:
                    ent.update(bound="hbm", frac=round(a_k["bytes"] / (us_k * 1e-6) / 1e9 / pk["hbm_gbs"], 3))
            kernels[name] = ent
        dom = max(rep.items(), key=lambda kv: kv[1][1])[0]
        cnt, ms = rep[dom]
        avg_s = ms / cnt * 1e-3
        a = alg.get(dom)
        if a is not None:
            if a["bound"] == "tensor":
                ach = a["flops"] / avg_s / 1e12
                peak = pk["bf16_tflops_sustained"]
                roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                            "frac": ach / peak, "traffic": ncu_traffic(dom)[0], "traffic_source": ncu_traffic(dom)[1],
                            "note": f"algorithmic fp32-equivalent FLOPs (2MNK); the kernel issues 3 bf16 MMAs per product, "
                                    f"so the tensor pipe runs at 3x this; peak = bf16 sustained of {pk['source']} "
                                    f"MEASURED_PEAKS.json; avg of {cnt} launches {avg_s*1e6:.1f} us (CUDA events)"}
            else:
                ach = a["bytes"] / avg_s / 1e9
                peak = pk["hbm_gbs"]
                roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                            "traffic": ncu_traffic(dom)[0], "traffic_source": ncu_traffic(dom)[1], "note": f"algorithmic bytes / avg of {cnt} launches ({avg_s*1e6:.1f} us); peak of {pk['source']}"}
        # BASELINE.json second metric: SHT achieved HBM GB/s (forward transform = DFT + Legendre kernels)
        if "sht.dft_fwd" in rep and "sht.legendre_fwd" in rep:
            t_f = (rep["sht.dft_fwd"][1] + rep["sht.legendre_fwd"][1]) / rep["sht.dft_fwd"][0] * 1e-3
            t_i = (rep["sht.dft_inv"][1] + rep["sht.legendre_inv"][1]) / rep["sht.dft_inv"][0] * 1e-3
            by = sht_transform_bytes(B)
            roofline_sht = {"bound": "hbm", "unit": "GB/s", "peak": pk["hbm_gbs"], "algorithmic_bytes": by,
                            "forward": {"achieved": by / t_f / 1e9, "frac": by / t_f / 1e9 / pk["hbm_gbs"], "us": t_f * 1e6},
                            "inverse": {"achieved": by / t_i / 1e9, "frac": by / t_i / 1e9 / pk["hbm_gbs"], "us": t_i * 1e6}}

    # ---- whole-step roofline: the stable headline (3.70 GB and 1.26 TFLOP per B = 1 step)
    step_roofline = None
    if rank == 0:
        pk = peaks()
        step_bytes, step_flops = 3.70e9 * B, 1.26e12 * B
        t_s = ms_per_step * 1e-3
        step_roofline = {
            "algorithmic_bytes": step_bytes, "algorithmic_flops": step_flops,
            "hbm_frac": step_bytes / t_s / 1e9 / pk["hbm_gbs"], "hbm_peak_gbs": pk["hbm_gbs"],
            "tensor_frac_fp32_equivalent": step_flops / t_s / 1e12 / pk["bf16_tflops_sustained"],
            "tensor_frac_issued_mmas": 3.0 * step_flops / t_s / 1e12 / pk["bf16_tflops_sustained"],
            "bf16_peak_tflops_sustained": pk["bf16_tflops_sustained"], "peaks": pk["source"],
            "note": "per-GPU; issued MMAs = 3 bf16 products per fp32-equivalent product (split-bf16), DFT-as-GEMM work not counted",
        }

    # ---- PyTorch eager on the same GPU (rank 0, N = 1 only): the reference algorithm through cuFFT + cuBLAS + ATen
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager_baseline:
        try:
            gpu_eager = time_gpu_eager(dev, in_names, out_names, means, stds)
        except Exception as e:  # noqa: BLE001  (a baseline must not take the bench line down)
            gpu_eager = {"error": repr(e)[:300]}

    # ---- CPU baseline (rank 0, N = 1 only): reference algorithm on the host cores, bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = time_cpu_reference(max_seconds=30.0, max_steps=3, warmup=1)
        cpu_baseline = {"value": 86400.0 / r["sec_per_step"] / STEPS_PER_YEAR, "unit": "sim-years/day", "cores": r["cores"],
                        "kind": "port", "sample": f"{r['steps']} ACE2 1-degree steps (B=1) after 1 warm-up, oracle port of the "
                                                  f"reference modules, torch CPU fp32, {r['cores']} threads (fastest of a sweep, "
                                                  f"host has {r['host_cpus']} logical CPUs), {r['sec_per_step']:.2f} s/step"}

    if rank == 0:
        line = {
            "metric": "simulated_years_per_day", "value": value, "unit": "sim-years/day", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 (split-bf16 3-term products, fp32 accumulate; fp32 I/O)", "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "members_per_gpu": B, "global_members": B * world, "parallelism": f"ensemble-dp{world}",
                "weights": "random init (reference initialisation, seed 0)",
                "l2": "inputs larger than L2: one step streams 1.7 GB of dhconv weights + ~100 MB activation tensors per kernel (L2 = 126 MB); no explicit flush",
                "timed": "CUDA-graph replay per step, forcing window resident in HBM; per 40-step window one area-weighted-mean "
                         "reduction of the outputs + one all_gather of [members, 50] diagnostics inside the timed region; "
                         "value = median of the repeats",
            },
            "e2e": {"value": e2e_value, "unit": "sim-years/day", "ms_per_step": e2e_ms_per_step, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "path": "FusedStepper.rollout_host (C ABI ace_stepper_step per step): pinned host forcing in and all 50 output fields out every step, copies on side streams overlapping compute"},
            "gpu_launches": int((launches_per_step + 0) * K + diag_launches_per_window * ((K + WINDOW - 1) // WINDOW)),
            "launches_per_step": int(launches_per_step),
            "repeats": {"n": len(samples), "ms_per_step": [round(v, 4) for v in samples], "min": ms_sorted[0], "median": ms_per_step,
                        "e2e_ms_per_step": [round(v, 4) for v in e2e_samples]},
            "sustained": sustained, "step_roofline": step_roofline,
            "clocks": clocks, "roofline": roofline, "roofline_sht": roofline_sht, "kernel_time_shares": shares,
            "kernels": kernels,
            "gpu_eager_baseline": gpu_eager,
            "cpu_baseline": cpu_baseline, "outputs_finite": finite, "diagnostic_windows": diag["n"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="ensemble members per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--repeats", type=int, default=3, help="repeats of the K-step timed region (median reported)")
    ap.add_argument("--sustained-steps", type=int, default=STEPS_PER_YEAR, help="length of the sustained leg (0 = skip)")
    ap.add_argument("--workload", default="rollout", choices=["rollout", "sht", "inverse_sht", "quarter_degree", "csfno_block", "csfno_block_8_groups", "healpix"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("sht", "inverse_sht"):
        run_sht_workload(args)
    elif args.workload == "quarter_degree":
        run_quarter_degree(args)
    elif args.workload.startswith("csfno_block"):
        run_csfno_block(args)
    elif args.workload == "healpix":
        run_healpix(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
