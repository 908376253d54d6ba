#!/usr/bin/env python
"""bench.py — SPAIR per-cell object pipeline, train images/sec (fwd+bwd) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

Workload (BASELINE.json configs[1]): spair/config.py defaults (1x128x128 canvas, 11x11 cells,
28x28 glimpses, 50 attributes), batch 256 PER GPU (weak scaling), procedurally generated
scattered-sprite scenes, fp32 everywhere, global_step >= 1000 so that every gradient is live.
A "step" is what reference train.py:64-67 does: zero grads, forward, backward, Adam step (plus the
NCCL gradient allreduce when N > 1).

One JSON line is printed by rank 0:
  value        whole-job images/s with the input batch already resident in HBM
  e2e          same metric through the public API with HOST inputs: every step copies the batch from
               pinned host memory to the device and reads the loss back to the host
  roofline     the dominant hand-written kernel (fused render backward), timed in isolation with CUDA
               events and an L2 flush between launches, against MEASURED_PEAKS.json's HBM copy peak
  cpu_baseline the CPU oracle (a restatement of the reference's op sequence, oracle/spair_oracle.py)
               timed on this box's host cores on a bounded sample (rank 0, N=1 only)
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train images/sec (fwd+bwd)"
UNIT = "images/s"
PER_GPU_BATCH = 256
CONFIG_NAME = "A"          # spair/config.py defaults
STEP0 = 1000               # training wheel off: all gradients live


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


class NullWriter:
    def add_scalar(self, *a, **k):
        pass

    add_image = add_figure = add_histogram = add_scalar


def make_batches(n_batches, batch, image_shape, seed):
    from spair_pytorch_b200.dataloader import scattered_sprites
    return [scattered_sprites(batch, image_shape, seed=seed + 1000 * i).pin_memory() for i in range(n_batches)]


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def kernel_roofline(net, x, steps=20):
    """Times the hand-written warp kernels in isolation on the tensors of a real step (CUDA events on
    the launching stream, L2 flushed between launches) and converts to GB/s with the ALGORITHMIC bytes
    of SURVEY.md §8(d) / DESIGN.md."""
    from spair_pytorch_b200 import kernels as K
    dev = x.device
    B, C, I, _ = x.shape
    L = net._latents
    HW = L.Hc * L.Wc
    G = net._cfg.object_shape[0]
    N = B * HW
    with torch.no_grad():
        logits = net.object_decoder(L.attr.reshape(N, -1)).contiguous()
    zw, zd, zp = L.z_where.reshape(N, 4).contiguous(), L.depth.reshape(-1).contiguous(), L.pres.reshape(-1).contiguous()
    recon = torch.empty(B, C, I, I, device=dev)
    denom = torch.empty(B, I, I, device=dev)
    partial = torch.empty(K.render_num_tiles(B, I, I), device=dev)
    gs = torch.empty(B, C + 1, I, I, device=dev)
    d_logits, d_zw, d_zd, d_zp = torch.empty_like(logits), torch.empty_like(zw), torch.empty_like(zd), torch.empty_like(zp)
    cells = torch.arange(HW, dtype=torch.int32, device=dev)
    glimpses = torch.empty(N, C * G * G, device=dev)
    d_gl = torch.randn(N, C * G * G, device=dev)
    d_zw_l = torch.empty(N, 4, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    scales = net._cfg.scales

    def t_render_fwd():
        K.render_fwd(logits, zw, zd, zp, B, HW, C, G, I, I, scales, recon, denom, x, partial)

    def t_render_bwd():
        K.render_bwd(logits, zw, zd, zp, B, HW, C, G, I, I, scales, recon, denom, None, x, None, gs, d_logits, d_zw, d_zd, d_zp)

    def t_glimpse_fwd():
        K.glimpse_fwd(x, L.z_where, cells, B, HW, G, G, glimpses)

    def t_glimpse_bwd():
        K.glimpse_bwd(x, L.z_where, cells, B, HW, G, G, d_gl, d_zw_l, None)

    # backbone stem (ZeroPad2d + conv_0 + bias + ReLU; csrc/stem.cu) on the step's own batch
    conv0 = net.backbone.net[0]
    pl, pr, pt, pb = net.backbone.padding.padding
    s0 = conv0.stride[0]
    Ho, Wo = (I + pt + pb - 4) // s0 + 1, (I + pl + pr - 4) // s0 + 1
    stem_y = torch.empty(B, conv0.out_channels, Ho, Wo, device=dev)
    stem_dy = torch.randn_like(stem_y)
    stem_ws = K.stem_bwd_workspace(C, conv0.out_channels, dev)
    stem_dw, stem_db = torch.empty_like(conv0.weight), torch.empty_like(conv0.bias)
    w0, b0 = conv0.weight.detach().contiguous(), conv0.bias.detach().contiguous()

    def t_stem_fwd():
        K.stem_conv_fwd(x, w0, b0, s0, pt, pl, Ho, Wo, stem_y)

    def t_stem_bwd():
        K.stem_conv_bwd(x, stem_y, stem_dy, tuple(w0.shape), s0, pt, pl, stem_ws, stem_dw, stem_db)

    # algorithmic bytes per image
    per_image = {
        "stem_fwd": 4 * C * I * I + 4 * conv0.out_channels * Ho * Wo,
        "stem_bwd": 4 * C * I * I + 8 * conv0.out_channels * Ho * Wo,
        "render_fwd": HW * (4 * (C + 1) * G * G + 24) + 4 * C * I * I,
        "render_bwd": 4 * C * I * I + HW * (8 * (C + 1) * G * G + 48),
        "glimpse_fwd": 4 * C * I * I + HW * (16 + 4 * C * G * G),
        "glimpse_bwd": 4 * C * I * I + HW * (4 * C * G * G + 32),
    }
    peak, peak_src = measured_hbm_peak()
    out = {}
    for name, fn in (("render_fwd", t_render_fwd), ("render_bwd", t_render_bwd), ("glimpse_fwd", t_glimpse_fwd),
                     ("glimpse_bwd", t_glimpse_bwd), ("stem_fwd", t_stem_fwd), ("stem_bwd", t_stem_bwd)):
        for _ in range(3):
            fn()
        times = []
        for _ in range(steps):
            flush.fill_(1.0)                      # 256 MB > 126 MB L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = statistics.mean(times)
        nbytes = per_image[name] * B
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "bytes": nbytes, "achieved": gbs, "frac": gbs / peak}
    # dominant of the glimpse / render kernels BASELINE's metric names (the stem kernels are the caller side of the path)
    dom = max((k for k in out if not k.startswith("stem")), key=lambda k: out[k]["ms"])
    # the fused cell sweep: ONE persistent launch per direction holding all 31 wavefronts x 4 three-layer MLPs (fp32 SIMT
    # dot products on <= 16 rows per CTA: latency / issue bound, not HBM work) — reported with its FLOP rate next to
    # the HBM-bound kernels.  Timed with CUDA events placed directly around the two launches inside the operator.
    from spair_pytorch_b200 import ops
    plan = net._plan
    sweep_ms = {"fwd": [], "bwd": []}
    feat = net.backbone(x).detach().requires_grad_(True)
    noise = net._draw_noise(B, HW, dev)
    st = net.prepare_step(STEP0, dev)
    sweep_args = (plan, x, feat, net.virtual_edge_element, *noise, st.wheel, *net._sweep_params())
    for it in range(3 + steps):
        flush.fill_(1.0)
        ops.SWEEP_EVENTS = {}
        outs = ops.CellSweepFunction.apply(*sweep_args)
        torch.autograd.backward([o.sum() for o in outs[:4]], inputs=[feat])
        torch.cuda.synchronize()
        ev, ops.SWEEP_EVENTS = ops.SWEEP_EVENTS, None
        if it >= 3:
            for k in sweep_ms:
                if k in ev:
                    sweep_ms[k].append(ev[k][0].elapsed_time(ev[k][1]))
    macs = sum(w.shape[0] * w.shape[1] for m in plan.last_mlps for w in m.W)
    fp32_peak = 148 * 128 * 2 * 1.92e9 / 1e12          # SIMT FMA peak at the 1.92 GHz boost clock, TFLOP/s
    for k, label in (("fwd", "sweep_fwd"), ("bwd", "sweep_bwd")):
        if sweep_ms[k]:
            ms = statistics.mean(sweep_ms[k])
            tf = 2 * macs * N / (ms * 1e-3) / 1e12
            out[label] = {"ms": ms, "flops": 2 * macs * N, "achieved_tflops": tf, "fp32_simt_peak_tflops": fp32_peak,
                          "frac_of_fp32_simt_peak": tf / fp32_peak,
                          "note": "persistent fused cell sweep (%s): context + 4 MLPs + heads + glimpse for all wavefronts "
                                  "in one launch; fp32 SIMT FMAs, <= 16 rows per CTA and wavefront" % k}
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the same
    # kernels at this shape (profiles/r01_warp_kernels_final_full.md; B=256, C=1, I=128, 121 cells, G=28)
    ncu_traffic = {"render_bwd": 228623872 + 159112192, "render_fwd": 201631488 + 24653056,
                   "glimpse_fwd": 17969920 + 44048128, "glimpse_bwd": 115137792 + 4571392}
    traffic = ncu_traffic.get(dom) if (B, C, I, HW, G) == (256, 1, 128, 121, 28) else None
    roof = {"bound": "hbm", "kernel": dom, "achieved": out[dom]["achieved"], "peak": peak, "unit": "GB/s",
            "frac": out[dom]["frac"], "traffic": traffic,
            "traffic_source": "profiles/r01_warp_kernels_final_full.md (ncu --set full, per launch)" if traffic else None,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": out[dom]["bytes"], "ms_per_launch": out[dom]["ms"],
            "timing": "CUDA events on the launching stream, 256 MB L2 flush between launches, mean of %d" % steps,
            "note": "dominant kernel of the HBM-bound class BASELINE's metric names (glimpse / render); the two persistent sweep "
                    "kernels are larger by time but are fp32 SIMT dot-product chains, reported under kernels.sweep_fwd / "
                    "kernels.sweep_bwd against the fp32 FMA peak; stem_fwd / stem_bwd are the caller-side HBM kernels"}
    return roof, out


def cpu_baseline(sample_batch=32, iters=2):
    """The CPU oracle (port of the reference's op sequence) on this host: forward + backward."""
    from oracle import spair_oracle as so
    from tests import helpers
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = helpers.oracle_config(CONFIG_NAME)
    net = helpers.build_model(CONFIG_NAME)
    params = so.params_from_state_dict(net.state_dict())
    x = so.scattered_sprites(sample_batch, cfg.image_shape, seed=1234)
    noise = so.random_noise(torch.Generator().manual_seed(7), sample_batch, cfg.grid, cfg.n_attr)
    so.forward_backward(params, x[:2], STEP0 + 1, so.Noise(*(t[:2] for t in (noise.eps_where, noise.eps_attr, noise.eps_depth,
                                                                               noise.u_pres))), cfg)   # warm-up
    t0 = time.time()
    for _ in range(iters):
        so.forward_backward(params, x, STEP0 + 1, noise, cfg)
    dt = (time.time() - t0) / iters
    return {"value": sample_batch / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d timed fwd+bwd iterations of batch %d (config defaults, step %d) after 1 warm-up; %.2f s/iter"
                      % (iters, sample_batch, STEP0 + 1, dt)}


def run_ours(args):
    from spair_pytorch_b200 import dp, kernels as K
    from tests import helpers
    rank, world, local_rank = dp.init_distributed("nccl" if "RANK" in os.environ else None)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.benchmark = True      # fixed shapes: let cuDNN pick its fastest fp32 algorithms for the backbone
    net = helpers.build_model(CONFIG_NAME, dev)
    ddp = dp.DataParallelSPAIR(net, world_size=world)
    ddp.broadcast_parameters()
    opt = torch.optim.Adam([ddp.bucket.flatten_parameters()], lr=1e-4, fused=True)   # one kernel over the flat parameters
    B = args.batch
    image_shape = tuple(net.image_shape)
    host_batches = make_batches(4, B, image_shape, seed=1234 + rank)
    dev_batches = [b.to(dev) for b in host_batches]
    x_stage = torch.empty_like(dev_batches[0])
    loss_host = torch.empty((), pin_memory=True)
    torch.manual_seed(7 + rank)
    launches_per_step = None
    if args.eager:
        def fwd_bwd(x, step):
            return ddp.step(x, step)                       # zero grads, forward, backward, allreduce
    else:
        from spair_pytorch_b200.graphed import GraphedTrainStep
        n0 = K.launch_count()
        gstep = GraphedTrainStep(net, dev_batches[0], bucket=ddp.bucket, global_step=STEP0, warmup=3)
        launches_per_step = (K.launch_count() - n0) // 4   # 3 warm-up steps + 1 captured step

        def fwd_bwd(x, step):
            out = gstep(x, step)                           # one CUDA-graph replay: zero grads, forward, backward
            if world > 1:
                ddp.bucket.all_reduce()
            return out

    def step_resident(i):
        fwd_bwd(dev_batches[i % len(dev_batches)], STEP0 + i)
        opt.step()

    def step_e2e(i):
        if args.eager:
            x_stage.copy_(host_batches[i % len(host_batches)], non_blocking=True)  # H2D from pinned memory
            loss = fwd_bwd(x_stage, STEP0 + i)[0]
        else:
            # input double buffering (GraphedTrainStep.prefetch): this step consumes the batch whose H2D copy was started
            # during the previous step and starts the copy of the next one, which overlaps this step's kernels.  Every
            # timed step still performs exactly one 16.8 MB pinned H2D copy and one D2H read inside the timed region.
            if not gstep._has_staged:
                gstep.prefetch(host_batches[i % len(host_batches)])
            loss = fwd_bwd(None, STEP0 + i)[0]
            gstep.prefetch(host_batches[(i + 1) % len(host_batches)])
        opt.step()
        loss_host.copy_(loss.detach(), non_blocking=False)                          # D2H read of the loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        n0 = K.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            e0.record()
            for i in range(steps):
                fn(warmup + i)
            e1.record()
            barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        n_launch = K.launch_count() - n0 if launches_per_step is None else launches_per_step * steps
        return float(ms) / steps, n_launch, clk.summary()

    if args.profile_step:       # for `ncu --profile-from-start off`: only the steps after warm-up are inside the
        for i in range(args.warmup):            # cudaProfilerStart/Stop range (no e2e pass, no microbench, no JSON)
            step_resident(i)
        barrier()
        torch.cuda.profiler.start()
        for i in range(args.steps):
            step_resident(args.warmup + i)
        barrier()
        torch.cuda.profiler.stop()
        return None
    ms_res, launches, clocks = timed(step_resident, args.steps, args.warmup)
    ms_e2e, _, _ = timed(step_e2e, args.steps, max(args.warmup, 3))
    value = world * B / (ms_res * 1e-3)
    e2e = world * B / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        roof, per_kernel = kernel_roofline(net, dev_batches[0])
        HW = net._latents.Hc * net._latents.Wc
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: spair/config.py defaults (1x128x128 canvas, 11x11 cells, 28x28 glimpses), "
                                   "batch %d per GPU, procedural scattered sprites, step = zero_grad+fwd+bwd+Adam%s"
                                   % (B, "+NCCL grad allreduce" if world > 1 else ""),
                       "per_gpu_batch": B, "global_batch": world * B, "objects_per_image": HW, "global_step": STEP0,
                       "parallelism": "dp%d" % world, "tf32": False,
                       "launch": "eager" if args.eager else "fwd+bwd replayed from one CUDA graph; allreduce + fused Adam eager",
                       "l2": "per-step working set (~2.3 KB x %d objects x fwd+bwd buffers, > 1 GB) exceeds the 126 MB L2; "
                             "kernel timings flush L2 between launches" % (B * HW)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": world * host_batches[0].numel() * 4,
                    "d2h_bytes_per_step": world * 4,
                    "note": "public API GraphedTrainStep with host batches in pinned memory: one H2D copy of a whole batch "
                            "and one blocking D2H read of the loss per step, both inside the timed region; the H2D copy of "
                            "step i+1 is issued on a copy stream while step i runs (input double buffering)"
                            if not args.eager else "eager step: H2D copy, forward, backward, Adam, blocking D2H loss read"},
            "gpu_launches": launches,
            "roofline": roof,
            "kernels": per_kernel,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return line


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own algorithm on the host CPU (oracle port; the Python reference
# itself cannot travel to the GPU box)
# --------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    from oracle import spair_oracle as so
    from tests import helpers
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = helpers.oracle_config(CONFIG_NAME)
    with contextlib.redirect_stdout(io.StringIO()):
        net = helpers.build_model(CONFIG_NAME)
    params = so.params_from_state_dict(net.state_dict())
    # bounded sample: calibrate on 2 images, then pick a batch so that steps+warmup stay within ~150 s
    probe = 2
    xs = so.scattered_sprites(32, cfg.image_shape, seed=1234)
    noise = so.random_noise(torch.Generator().manual_seed(7), 32, cfg.grid, cfg.n_attr)

    def sub(n):
        return xs[:n], so.Noise(noise.eps_where[:n], noise.eps_attr[:n], noise.eps_depth[:n], noise.u_pres[:n])

    t0 = time.time()
    so.forward_backward(params, *[sub(probe)[0]], STEP0 + 1, sub(probe)[1], cfg)
    t_probe = time.time() - t0
    budget = 150.0 / max(args.steps + args.warmup, 1)
    batch = int(max(1, min(32, probe * budget / max(t_probe, 1e-3))))
    x, nz = sub(batch)
    for _ in range(args.warmup):
        so.forward_backward(params, x, STEP0 + 1, nz, cfg)
    t0 = time.time()
    for i in range(args.steps):
        so.forward_backward(params, x, STEP0 + 1 + i, nz, cfg)
    dt = (time.time() - t0) / max(args.steps, 1)
    value = batch / dt
    sample = ("oracle port of the reference's CPU op sequence (per-cell loop, materialised render), batch %d per step, "
              "%d ATen threads; the Python reference itself cannot travel to the GPU box" % (batch, cores))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] workload (spair/config.py defaults), bounded sample of batch %d per step "
                                   "on the host CPU, step = fwd+bwd" % batch, "per_step_batch": batch, "global_step": STEP0 + 1},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return line


_REAL_STDOUT = None


def claim_stdout():
    """Route everything that writes to fd 1 (NCCL's version banner, stray prints of libraries) to stderr and
    keep the real stdout for the ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--profile-step", action="store_true", help="run warm-up + timed steps only and print nothing (for ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when asked for N > 1 from a plain `python bench.py`
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=_REAL_STDOUT))
    run_ours(args)


if __name__ == "__main__":
    main()
