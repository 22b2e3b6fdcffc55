"""Summarise the per-tile clock samples the tcgen05 GEMM kernel writes with option "trace" (development tool).

    ACE_B200_TRACE_FILE=gpurun_out/trace.bin python tools/gemm_probe.py 1 '{"trace":1}'
    python tools/trace_report.py gpurun_out/trace.bin [--tiles]

Record layout (gemm_umma.cu: launch()): char name[64], int32 {grid, iters, slots, stages, tile_m, bn, bk, pair}, then
int64 [grid][iters][slots].  Slots per (CTA, tile iteration):
  0 MMA warp: loop top            1 after the accumulator-stage (tempty) wait   2 cycles spent waiting for full stages
  3 after the last commit         4 producer: tile start                        5 cycles spent waiting for empty stages
  6 producer: tile end            7/8/9   epilogue warp 4:  before / after the tfull wait, after the tile's stores
  10/11/12 epilogue warp 15       13 (num_kc << 32) | n_count
Times are SM clocks (clock64) of the CTA's own SM; reported in microseconds at --mhz.
"""
import json
import struct
import sys

import numpy as np


def records(path):
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    while off < len(data):
        name = data[off:off + 64].split(b"\0")[0].decode()
        hdr = struct.unpack_from("8i", data, off + 64)
        grid, iters, slots = hdr[:3]
        n = grid * iters * slots
        arr = np.frombuffer(data, dtype=np.int64, count=n, offset=off + 96).reshape(grid, iters, slots)
        off += 96 + 8 * n
        yield name, hdr, arr


def summarise(name, hdr, a, mhz, tiles=False):
    grid, iters, slots, stages, tile_m, bn, bk, pair = hdr
    us = lambda c: float(c) / mhz
    hdr_rows = a[:, iters - 1, :]
    a = a[:, :iters - 1, :]
    iters -= 1
    lead = a[::2] if pair else a  # MMA samples exist on the leader CTA only
    g0, c0, g1, c1, g2, c2 = (hdr_rows[:, i].astype(np.float64) for i in range(6))
    t_first = g0.min()
    out_hdr = {
        "kernel_extent_us": round((g2.max() - t_first) / 1e3, 2),
        "cta_entry_skew_us": round((g0.max() - t_first) / 1e3, 2),
        "setup_us_med": round(float(np.median(g1 - g0)) / 1e3, 2),
        "body_us_med": round(float(np.median(g2 - g1)) / 1e3, 2),
        "body_us_max": round(float((g2 - g1).max()) / 1e3, 2),
        "cta_end_spread_us": round((g2.max() - g2.min()) / 1e3, 2),
        "sm_mhz_med": round(float(np.median((c2 - c1) / np.maximum(g2 - g1, 1.0))) * 1e3, 0),
    }
    mhz = out_hdr["sm_mhz_med"] if out_hdr["sm_mhz_med"] > 0 else mhz
    out = {"kernel": name, "grid": grid, "stages": stages, "tile": [tile_m, bn, bk], "pair": pair}
    out.update(out_hdr)
    ntile = (lead[:, :, 3] != 0).sum(1)
    out["tiles_per_cta"] = [int(ntile.min()), float(np.median(ntile)), int(ntile.max())]
    if ntile.max() >= iters:
        out["note"] = "tile iterations beyond %d not recorded" % iters
    span, w_acc, w_full, issue, mma_units = [], [], [], [], []
    for c in range(lead.shape[0]):
        n = ntile[c]
        if n == 0:
            continue
        r = lead[c, :n]
        span.append(r[n - 1, 3] - r[0, 0])
        w_acc.append((r[:, 1] - r[:, 0]).sum())
        w_full.append(r[:, 2].sum())
        issue.append((r[:, 3] - r[:, 1] - r[:, 2]).sum())
        kc = r[:, 13] >> 32
        ncnt = r[:, 13] & 0xffffffff
        mma_units.append(int((kc * ((ncnt + 15) // 16 * 16)).sum()))
    med = lambda v: round(us(np.median(v)), 1)
    out["mma_warp_us"] = {"span": med(span), "span_max": round(us(max(span)), 1), "wait_accumulator": med(w_acc), "wait_full_stage": med(w_full),
                          "issue": med(issue)}
    out["mma_work_units(kc*n)"] = [min(mma_units), int(np.median(mma_units)), max(mma_units)]
    pw, pspan = [], []
    ew, ework, espan = [], [], []
    for c in range(a.shape[0]):
        n = int((a[c, :, 6] != 0).sum())
        if n:
            r = a[c, :n]
            pw.append(r[:, 5].sum())
            pspan.append(r[n - 1, 6] - r[0, 4])
        for base in (7, 10):
            n = int((a[c, :, base + 2] != 0).sum())
            if n:
                r = a[c, :n]
                ew.append((r[:, base + 1] - r[:, base]).sum())
                ework.append((r[:, base + 2] - r[:, base + 1]).sum())
                espan.append(r[n - 1, base + 2] - r[0, base])
    out["producer_us"] = {"span": med(pspan), "wait_empty_stage": med(pw)}
    out["epilogue_warp_us"] = {"span": med(espan), "wait_accumulator_full": med(ew), "work": med(ework), "work_per_tile": round(us(np.median(ework)) / max(1.0, float(np.median(ntile))), 2)}
    print(json.dumps(out))
    if tiles:
        c = 0
        n = ntile[0]
        t0 = lead[c, 0, 0]
        for i in range(n):
            r = lead[c, i]
            print("   tile %2d kc %3d n %3d | mma: top %7.2f acc_wait %5.2f full_wait %5.2f end %7.2f | prod %7.2f..%7.2f wait %5.2f | epi4 wait %7.2f got %7.2f done %7.2f | epi15 got %7.2f done %7.2f" % (
                i, r[13] >> 32, r[13] & 0xffffffff, us(r[0] - t0), us(r[1] - r[0]), us(r[2]), us(r[3] - t0), us(r[4] - t0), us(r[6] - t0), us(r[5]),
                us(r[7] - t0), us(r[8] - t0), us(r[9] - t0), us(r[11] - t0), us(r[12] - t0)))


if __name__ == "__main__":
    path = sys.argv[1]
    tiles = "--tiles" in sys.argv
    mhz = 1965.0
    seen = {}
    for name, hdr, arr in records(path):
        seen[name] = (hdr, arr)  # keep the last launch of every kernel name
    for name, (hdr, arr) in seen.items():
        summarise(name, hdr, arr, mhz, tiles)
