#!/usr/bin/env python
"""Kernel microbench (BASELINE.json configs[4]): glimpse extract + fused render, forward and backward,
over cells x glimpse size x channels, reported as GB/s of ALGORITHMIC bytes against the
measured HBM peak.  Inputs follow SURVEY.md §8(d): xt,yt ~ U(0,1), xs,ys ~ U(12/I, 48/I), logits ~ N(0,1),
z_pres ~ U(0,1), z_depth ~ U(0,4).  Timing: CUDA events on the launching stream, 256 MB L2 flush between
launches, 3 warm-up launches.

    python tools/kernel_microbench.py                 # the sweep, one JSON line per case
    python tools/kernel_microbench.py --case A --once # one launch of each kernel (for ncu)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from spair_pytorch_b200 import kernels as K  # noqa: E402

CASES = {
    # name: (C, I, Hc, G, B)
    "A": (1, 128, 11, 28, 256),      # BASELINE configs[1] shape
    "C": (1, 128, 16, 14, 512),      # configs[2] shape
    "D": (3, 256, 32, 28, 32),       # configs[3] shape (32 images = 32k objects)
    "A14": (1, 128, 11, 14, 256),
    "C28": (1, 128, 16, 28, 256),
    "D14": (3, 256, 32, 14, 32),
}


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def make_inputs(C, I, Hc, G, B, dev, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    HW = Hc * Hc
    N = B * HW
    zw = torch.rand(B, HW, 4, device=dev, generator=g)
    zw[..., 2:] = (12.0 + 36.0 * zw[..., 2:]) / I
    t = dict(x=torch.rand(B, C, I, I, device=dev, generator=g), zw=zw.contiguous(),
             logits=torch.randn(N, G, G, C + 1, device=dev, generator=g), zd=4 * torch.rand(N, device=dev, generator=g),
             zp=torch.rand(N, device=dev, generator=g), recon=torch.empty(B, C, I, I, device=dev),
             denom=torch.empty(B, I, I, device=dev), partial=torch.empty(K.render_num_tiles(B, I, I), device=dev),
             gs=torch.empty(B, C + 1, I, I, device=dev), cells=torch.arange(HW, dtype=torch.int32, device=dev),
             glimpses=torch.empty(N, C * G * G, device=dev), d_gl=torch.randn(N, C * G * G, device=dev, generator=g),
             d_zw_l=torch.empty(N, 4, device=dev))
    t["d_logits"], t["d_zw"] = torch.empty_like(t["logits"]), torch.empty(N, 4, device=dev)
    t["d_zd"], t["d_zp"] = torch.empty(N, device=dev), torch.empty(N, device=dev)
    return t


def kernels_for(C, I, Hc, G, B, t):
    HW = Hc * Hc
    scales = (2.0, 0.1, 5.0)
    zw_flat = t["zw"].view(-1, 4)
    fns = {
        "glimpse_fwd": lambda: K.glimpse_fwd(t["x"], t["zw"], t["cells"], B, HW, G, G, t["glimpses"]),
        "glimpse_bwd": lambda: K.glimpse_bwd(t["x"], t["zw"], t["cells"], B, HW, G, G, t["d_gl"], t["d_zw_l"], None),
        "render_fwd": lambda: K.render_fwd(t["logits"], zw_flat, t["zd"], t["zp"], B, HW, C, G, I, I, scales, t["recon"],
                                           t["denom"], t["x"], t["partial"]),
        "render_bwd": lambda: K.render_bwd(t["logits"], zw_flat, t["zd"], t["zp"], B, HW, C, G, I, I, scales, t["recon"],
                                           t["denom"], None, t["x"], None, t["gs"], t["d_logits"], t["d_zw"], t["d_zd"],
                                           t["d_zp"]),
    }
    per_image = {
        "glimpse_fwd": 4 * C * I * I + HW * (16 + 4 * C * G * G),
        "glimpse_bwd": 4 * C * I * I + HW * (4 * C * G * G + 32),
        "render_fwd": HW * (4 * (C + 1) * G * G + 24) + 4 * C * I * I,
        "render_bwd": 4 * C * I * I + HW * (8 * (C + 1) * G * G + 48),
    }
    return fns, {k: v * B for k, v in per_image.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--once", action="store_true", help="launch each kernel once (after one warm-up) and exit")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda")
    peak = hbm_peak()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    for name in ([args.case] if args.case else list(CASES)):
        C, I, Hc, G, B = CASES[name]
        t = make_inputs(C, I, Hc, G, B, dev)
        fns, nbytes = kernels_for(C, I, Hc, G, B, t)
        fns["render_fwd"]()           # recon/denom must exist before render_bwd
        if args.once:
            for fn in fns.values():
                fn()
            torch.cuda.synchronize()
            for fn in fns.values():
                flush.fill_(1.0)
                fn()
            torch.cuda.synchronize()
            continue
        for k, fn in fns.items():
            for _ in range(3):
                fn()
            times = []
            for _ in range(args.iters):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                times.append(e0.elapsed_time(e1))
            ms = statistics.mean(times)
            gbs = nbytes[k] / (ms * 1e-3) / 1e9
            print(json.dumps({"case": name, "shape": dict(C=C, I=I, cells=Hc * Hc, G=G, B=B), "kernel": k, "ms": round(ms, 4),
                              "min_ms": round(min(times), 4), "algorithmic_bytes": nbytes[k], "GB/s": round(gbs, 1),
                              "frac_of_measured_hbm_peak": round(gbs / peak, 4), "peak_GB/s": peak}), flush=True)


if __name__ == "__main__":
    main()
