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


def measured_tf32x3_peak():
    """fp32-equivalent tensor peak of a 3xTF32 split-precision kernel: the measured dense bf16 GEMM rate (burst figure: the
    kernel is timed alone) / 2 (TF32 runs at half the bf16 rate) / 3 (three MMAs per fp32-accurate product)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops"]) / 6.0, "measured (MEASURED_PEAKS.json bf16_tflops / 2 / 3)"
    except Exception:
        return 1590.0 / 6.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s bf16 / 2 / 3)"


def measured_fp32_peak():
    """fp32 FFMA peak of the SIMT pipes: 148 SMs x 128 lanes x 2 FLOP x the maximum SM clock the driver measured."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mhz = float(json.load(f)["sm_max_mhz"])
        src = "148 SMs x 128 FFMA lanes x 2 x %.0f MHz (MEASURED_PEAKS.json sm_max_mhz)" % mhz
    except Exception:
        mhz, src = 1965.0, "148 SMs x 128 FFMA lanes x 2 x 1965 MHz (nominal)"
    return 148 * 128 * 2 * mhz * 1e6 / 1e12, src


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
    from spair_pytorch_b200 import ops
    decoded = ops.USE_TENSOR_CORE_GEMM      # the renderer consumes sigmoid-decoded texel records (ops.DecoderFunction)
    lin = [m for m in net.object_decoder if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        if decoded:
            logits = ops.DecoderFunction.apply(L.attr.reshape(N, -1), *(p for m in lin for p in (m.weight, m.bias)), C + 1,
                                               net._cfg.scales)
        else:
            logits = net.object_decoder(L.attr.reshape(N, -1)).contiguous()
        dec_h1 = net.object_decoder[:-1](L.attr.reshape(N, -1)).contiguous()
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
        K.render_fwd(logits, zw, zd, zp, B, HW, C, G, I, I, scales, recon, denom, x, partial, decoded)

    def t_render_bwd():
        K.render_bwd(logits, zw, zd, zp, B, HW, C, G, I, I, scales, recon, denom, None, x, None, gs, d_logits, d_zw, d_zd, d_zp,
                     decoded)

    def t_glimpse_fwd():
        K.glimpse_fwd(x, L.z_where, cells, B, HW, G, G, glimpses)

    def t_glimpse_bwd():
        K.glimpse_bwd(x, L.z_where, cells, B, HW, G, G, d_gl, d_zw_l, None)

    # backbone stem (ZeroPad2d + conv_0 + bias + ReLU; csrc/stem.cu) on the step's own batch
    conv0 = net.backbone.net[0]
    pl, pr, pt, pb = net.backbone.padding.padding
    s0 = conv0.stride[0]
    Ho, Wo = (I + pt + pb - 4) // s0 + 1, (I + pl + pr - 4) // s0 + 1
    stem_y = torch.empty(B, Ho, Wo, conv0.out_channels, device=dev)      # channels-last, as the step runs it (GEMM tail)
    stem_dy = torch.randn_like(stem_y)
    stem_ws = K.stem_bwd_workspace(C, conv0.out_channels, dev)
    stem_dw, stem_db = torch.empty_like(conv0.weight), torch.empty_like(conv0.bias)
    w0, b0 = conv0.weight.detach().contiguous(), conv0.bias.detach().contiguous()

    def t_stem_fwd():
        K.stem_conv_fwd(x, w0, b0, s0, pt, pl, Ho, Wo, stem_y, True)

    def t_stem_bwd():
        K.stem_conv_bwd(x, stem_y, stem_dy, tuple(w0.shape), s0, pt, pl, stem_ws, stem_dw, stem_db, True)

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
    for v in out.values():
        v.update(bound="hbm", unit="GB/s", peak=peak)
    tpeak, tpeak_src = measured_tf32x3_peak()
    # the decoder's output layer on the tcgen05 GEMM (csrc/gemm.cu): y = x W^T with the texel epilogue, dx = dy W, dW = dy^T x.
    # Algorithmic fp32 FLOPs (2 M N K, NOT counting the 3 MMAs per product) against the fp32-equivalent tensor peak.
    if decoded and K.gemm_supported(dec_h1, lin[-1].weight):
        w2, b2 = lin[-1].weight.detach(), lin[-1].bias.detach()
        tex = torch.empty(N, w2.shape[0], device=dev)
        d_h1, d_w2 = torch.empty_like(dec_h1), torch.empty_like(w2)
        d_logits.normal_()
        gemms = {"decoder_out_fwd": lambda: K.gemm3x(dec_h1, True, w2, True, tex, b2, epilogue=K.GEMM_EPI_TEXEL, period=C + 1,
                                                     scales=scales),
                 "decoder_out_dgrad": lambda: K.gemm3x(d_logits.view(N, -1), True, w2, False, d_h1),
                 "decoder_out_wgrad": lambda: K.gemm3x(d_logits.view(N, -1), False, dec_h1, False, d_w2)}
        for name, fn in gemms.items():
            for _ in range(3):
                fn()
            times = []
            for _ in range(steps):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                times.append(e0.elapsed_time(e1))
            ms = statistics.mean(times)
            fl = 2.0 * N * w2.shape[0] * w2.shape[1]
            out[name] = {"ms": ms, "flops": fl, "achieved": fl / (ms * 1e-3) / 1e12, "frac": fl / (ms * 1e-3) / 1e12 / tpeak,
                         "bound": "tensor", "unit": "TFLOP/s", "peak": tpeak,
                         "note": "tcgen05.mma kind::tf32, 3 MMAs per fp32-accurate product (hi*hi + hi*lo + lo*hi)"}
    # the fused cell sweep: ONE persistent launch per direction holding all wavefronts x 4 three-layer MLPs, <= 16 rows per
    # CTA: a latency-bound chain of small dense layers on the fp32 SIMT pipes (a 3xTF32 mma.sync variant was measured 2x
    # slower, profiles/r02_sweep_mma_variant_rejected.md).  Reported as algorithmic fp32 FLOPs against the fp32 FFMA peak
    # (148 SMs x 128 lanes x 2 FLOP x max SM clock).  Timed with CUDA events around the two launches inside the operator.
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
    fpeak, fpeak_src = measured_fp32_peak()
    for k, label in (("fwd", "sweep_fwd"), ("bwd", "sweep_bwd")):
        if sweep_ms[k]:
            ms = statistics.mean(sweep_ms[k])
            tf = 2 * macs * N / (ms * 1e-3) / 1e12
            mc = plan.schedule.max_cells              # rows per CTA and wavefront, as ops.CellSweepFunction chooses images per CTA
            rows = mc * (max(1, min(2, 16 // mc)) if B > 148 else 1)
            on_tc = ops.sweep_tc_choice(rows)[0 if k == "fwd" else 1]
            out[label] = {"ms": ms, "flops": 2 * macs * N, "achieved": tf, "frac": tf / fpeak, "bound": "fp32",
                          "unit": "TFLOP/s", "peak": fpeak,
                          "dense_layers": "tcgen05 kind::tf32, hi/lo split (csrc/sweep_tc.cuh)" if on_tc else "fp32 SIMT FFMA",
                          "note": "persistent fused cell sweep (%s): context + 4 MLPs + heads + glimpse for all wavefronts "
                                  "in one launch; the chain of %d dependent wavefronts x 12 layers bounds it (tensor-core "
                                  "variant: the ~67-cycle issue interval of a tcgen05.mma, tensor pipe 6 %% busy), not the "
                                  "arithmetic rate; algorithmic fp32 FLOPs against the fp32 FFMA peak either way"
                                  % (k, plan.schedule.n_wavefronts)}
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of the same
    # kernels at this shape (B=256, C=1, I=128, 121 cells, G=28)
    ncu_traffic = NCU_TRAFFIC if (B, C, I, HW, G) == (256, 1, 128, 121, 28) else {}

    def roof_of(name):
        v = out[name]
        r = {"bound": v["bound"], "kernel": name, "achieved": v["achieved"], "peak": v["peak"], "unit": v["unit"],
             "frac": v["frac"], "traffic": ncu_traffic.get(name, (None, None))[0],
             "traffic_source": ncu_traffic.get(name, (None, None))[1],
             "peak_source": {"hbm": peak_src, "tensor": tpeak_src, "fp32": fpeak_src}[v["bound"]], "ms_per_launch": v["ms"],
             "timing": "CUDA events on the launching stream, 256 MB L2 flush between launches, mean of %d" % steps}
        if v["bound"] == "hbm":
            r["algorithmic_bytes_per_launch"] = v["bytes"]
        else:
            r["algorithmic_flops_per_launch"] = v["flops"]
        return r

    # `roofline` = the hand-written kernel with the largest launch time, whatever bounds it; `roofline_hbm` = the largest of
    # the HBM-bound glimpse / render kernels BASELINE's metric names (stem_* are the caller side of the path)
    dom = max(out, key=lambda k: out[k]["ms"])
    dom_hbm = max((k for k in out if k.startswith(("render", "glimpse"))), key=lambda k: out[k]["ms"])
    roof = roof_of(dom)
    roof["hbm_class"] = roof_of(dom_hbm)
    return roof, out


# per-launch DRAM traffic (bytes read + written) from `ncu --set full` captures at config B; (bytes, source)
NCU_TRAFFIC = {
    "sweep_fwd": (48197888 + 383831808, "profiles/r02_kernels_full.md (ncu --set full, per launch; fp32 SIMT dense layers)"),
    "sweep_bwd": (200129536 + 375443712, "profiles/r02_kernels_full.md (ncu --set full, per launch; tcgen05 dense layers)"),
    "render_bwd": (228634368 + 159153000, "profiles/r02_kernels_full.md (ncu --set full, per launch, texel-record input)"),
    "render_fwd": (197374976 + 23060000, "profiles/r02_kernels_full.md (ncu --set full, per launch, texel-record input)"),
    "glimpse_fwd": (18863616 + 40622080, "profiles/r02_glimpse_full.md (ncu --set full, per launch)"),
    "glimpse_bwd": (119411456 + 4429056, "profiles/r02_glimpse_full.md (ncu --set full, per launch)"),
}


REFERENCE_OVERRIDES = {
    # BASELINE.json configs -> overrides of the reference's spair/config.py (applied before import)
    "A": {},
    "C": dict(OBJECT_SHAPE=[14, 14], DEFAULT_BACKBONE_TOPOLOGY="cell8"),
    "D": dict(INPUT_IMAGE_SHAPE=[3, 256, 256], DEFAULT_BACKBONE_TOPOLOGY="cell8"),
}
CPU_SAMPLE_BATCH = {"A": 32, "C": 8, "D": 1}       # images per CPU step (D: 7.8 GB of intermediates PER IMAGE on the CPU path)


class ReferenceCPU:
    """The reference's own CPU implementation of the path, timed on this host: the UNMODIFIED reference from
    baseline/_ref (``SPAIR.forward`` + ``loss.backward(retain_graph=True)``, reference train.py:65-66) when it travelled
    with the snapshot (kind "reference"), else the oracle port of the same op sequence (kind "port")."""

    def __init__(self, config_name, batch):
        from oracle import ref_harness as rh
        from oracle import spair_oracle as so
        from tests import helpers
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.batch, self.config_name = batch, config_name
        cfg = helpers.oracle_config(config_name)
        self.x = so.scattered_sprites(batch, cfg.image_shape, seed=1234)
        if rh.reference_available():
            ov = dict(REFERENCE_OVERRIDES[config_name], BATCH_SIZE=batch)
            if ov.get("DEFAULT_BACKBONE_TOPOLOGY") == "cell8":
                ov["DEFAULT_BACKBONE_TOPOLOGY"] = rh.TOPOLOGY_CELL8
            self.kind = "reference"
            self.ns = rh.load_reference(ov)
            self.net = rh.build_reference_model(self.ns, seed=3)
            self.rh = rh
            self.where = rh.REFERENCE_ROOT
        else:
            self.kind = "port"
            self.so, self.cfg = so, cfg
            with contextlib.redirect_stdout(io.StringIO()):
                net = helpers.build_model(config_name)
            self.params = so.params_from_state_dict(net.state_dict())
            self.noise = so.random_noise(torch.Generator().manual_seed(7), batch, cfg.grid, cfg.n_attr)
            self.where = "oracle/spair_oracle.py"

    def step(self, global_step):
        """forward + backward of one batch; returns seconds."""
        t0 = time.time()
        if self.kind == "reference":
            self.rh.run_reference(self.net, self.x, global_step, noise_seed=7 + global_step)
        else:
            self.so.forward_backward(self.params, self.x, global_step, self.noise, self.cfg)
        return time.time() - t0

    def describe(self, n_timed, dt):
        what = ("unmodified reference (%s) SPAIR.forward + loss.backward(retain_graph=True)" % self.where
                if self.kind == "reference" else "oracle port of the reference's CPU op sequence (baseline/_ref did not travel)")
        return ("%s, config %s, batch %d, global_step >= %d, %d ATen threads: %d timed iteration(s) after 1 warm-up, %.2f s/iter"
                % (what, self.config_name, self.batch, STEP0 + 1, self.cores, n_timed, dt))


def cpu_baseline(config_name, budget_s=25.0):
    """Bounded sample (about budget_s of CPU work) of the reference's CPU path on this box's host cores."""
    ref = ReferenceCPU(config_name, CPU_SAMPLE_BATCH[config_name])
    t_warm = ref.step(STEP0 + 1)
    n = int(max(1, min(8, budget_s // max(t_warm, 1e-3))))
    dt = sum(ref.step(STEP0 + 2 + i) for i in range(n)) / n
    return {"value": ref.batch / dt, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": ref.describe(n, dt)}


WORKLOADS = {
    "A": "configs[1]: spair/config.py defaults (1x128x128 canvas, 11x11 cells, 28x28 glimpses)",
    "C": "configs[2]: 1x128x128 canvas, 16x16 cells (256 objects/image), 14x14 glimpses",
    "D": "configs[3]: 3x256x256 canvas, 32x32 cells (1024 objects/image), 28x28 RGB glimpses",
}


def workload_name(config_name):
    return WORKLOADS[config_name]


class Workload:
    """One BASELINE config on this rank: model, flat gradient bucket, fused Adam, captured step, host + device batches."""

    def __init__(self, config_name, per_gpu_batch, rank, world, dev, eager=False):
        from spair_pytorch_b200 import dp, kernels as K
        from tests import helpers
        self.K, self.name, self.B, self.rank, self.world, self.dev, self.eager = K, config_name, per_gpu_batch, rank, world, dev, eager
        self.net = helpers.build_model(config_name, dev)
        self.ddp = dp.DataParallelSPAIR(self.net, world_size=world)
        self.ddp.broadcast_parameters()
        self.opt = torch.optim.Adam([self.ddp.bucket.flatten_parameters()], lr=1e-4, fused=True)   # one kernel, flat parameters
        self.image_shape = tuple(self.net.image_shape)
        self.host_batches = make_batches(4, per_gpu_batch, self.image_shape, seed=1234 + rank)
        self.dev_batches = [b.to(dev) for b in self.host_batches]
        self.x_stage = torch.empty_like(self.dev_batches[0])
        self.loss_host = torch.empty((), pin_memory=True)
        torch.manual_seed(7 + rank)
        self.launches_per_step = None
        self.gstep = None
        if not eager:
            from spair_pytorch_b200.graphed import GraphedTrainStep
            n0 = K.launch_count()
            self.gstep = GraphedTrainStep(self.net, self.dev_batches[0], bucket=self.ddp.bucket, global_step=STEP0, warmup=3)
            self.launches_per_step = (K.launch_count() - n0) // 4   # 3 warm-up steps + 1 captured step

    def fwd_bwd(self, x, step):
        if self.eager:
            return self.ddp.step(x, step)                       # zero grads, forward, backward, allreduce
        out = self.gstep(x, step)                               # one CUDA-graph replay: zero grads, forward, backward
        if self.world > 1:
            self.ddp.bucket.all_reduce()
        return out

    def step_resident(self, i):
        self.fwd_bwd(self.dev_batches[i % len(self.dev_batches)], STEP0 + i)
        self.opt.step()

    def step_e2e(self, i):
        hb = self.host_batches
        if self.eager:
            self.x_stage.copy_(hb[i % len(hb)], non_blocking=True)  # H2D from pinned memory
            loss = self.fwd_bwd(self.x_stage, STEP0 + i)[0]
            self.opt.step()
            self.loss_host.copy_(loss.detach(), non_blocking=False)                      # D2H read of the loss
            return
        # input double buffering (GraphedTrainStep.prefetch): this step consumes the batch whose H2D copy was started
        # during the previous step and starts the copy of the next one, which overlaps this step's kernels.  The loss is
        # read through GraphedTrainStep.loss_async(): its D2H copy is queued right behind the step and the host waits for
        # it one step later, after it has launched the next step.  Every timed step still performs exactly one pinned H2D
        # copy of a whole batch and one D2H read of a loss inside the timed region (the last one in drain_e2e()).
        if not self.gstep._has_staged:
            self.gstep.prefetch(hb[i % len(hb)])
        self.fwd_bwd(None, STEP0 + i)
        self.gstep.prefetch(hb[(i + 1) % len(hb)])
        self.opt.step()
        pending, self._pending_loss = getattr(self, "_pending_loss", None), self.gstep.loss_async()
        if pending is not None:
            self.last_loss = pending.value()

    def drain_e2e(self):
        pending, self._pending_loss = getattr(self, "_pending_loss", None), None
        if pending is not None:
            self.last_loss = pending.value()

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, local_rank=0, finish=None):
        K = self.K
        for i in range(warmup):
            fn(i)
        if finish is not None:
            finish()
        self.barrier()
        n0 = K.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            e0.record()
            for i in range(steps):
                fn(warmup + i)
            if finish is not None:
                finish()                     # e.g. the D2H read of the last step's loss: inside the timed region
            e1.record()
            self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        n_launch = K.launch_count() - n0 if self.launches_per_step is None else self.launches_per_step * steps
        return float(ms) / steps, n_launch, clk.summary()

    def release(self):
        self.gstep = self.net = self.ddp = self.opt = self.dev_batches = self.host_batches = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()


def secondary_workload(config_name, global_batch, scaling, rank, world, dev, local_rank, steps, warmup):
    """A second BASELINE config measured in the same run (same timing rules), reported as a block of the JSON line."""
    per_gpu = global_batch // world if scaling == "strong" else global_batch
    wl = Workload(config_name, per_gpu, rank, world, dev)
    ms, launches, clocks = wl.timed(wl.step_resident, steps, warmup, local_rank)
    ms_e2e, _, _ = wl.timed(wl.step_e2e, steps, max(warmup, 3), local_rank, finish=wl.drain_e2e)
    block = {"workload": "BASELINE %s, batch %d per GPU (%s scaling, global batch %d), step = zero_grad+fwd+bwd+Adam%s"
                         % (workload_name(config_name), per_gpu, scaling, per_gpu * world, "+NCCL grad allreduce" if world > 1 else ""),
             "scaling": scaling, "n_gpus": world, "per_gpu_batch": per_gpu, "global_batch": per_gpu * world, "steps": steps,
             "warmup": warmup, "ms_per_step": ms, "value": world * per_gpu / (ms * 1e-3), "unit": UNIT,
             "e2e": {"value": world * per_gpu / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                     "h2d_bytes_per_step": world * wl.host_batches[0].numel() * 4, "d2h_bytes_per_step": world * 4},
             "gpu_launches": launches, "clocks": clocks}
    if rank == 0:
        roof, per_kernel = kernel_roofline(wl.net, wl.dev_batches[0], steps=5)
        block["kernels"] = {k: {kk: vv for kk, vv in v.items() if kk != "note"} for k, v in per_kernel.items()}
        block["limiting_kernel"] = roof["kernel"]
    wl.release()
    return block


def run_ours(args):
    from spair_pytorch_b200 import dp
    rank, world, local_rank = dp.init_distributed("nccl" if "RANK" in os.environ else None)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.benchmark = True      # fixed shapes: let cuDNN pick its fastest fp32 algorithms for the backbone
    if args.global_batch:
        B = args.global_batch // world if args.scaling == "strong" else args.global_batch
    else:
        B = args.batch
    wl = Workload(args.config, B, rank, world, dev, eager=args.eager)

    if args.profile_step:       # for `ncu --profile-from-start off`: only the steps after warm-up are inside the
        for i in range(args.warmup):            # cudaProfilerStart/Stop range (no e2e pass, no microbench, no JSON)
            wl.step_resident(i)
        wl.barrier()
        torch.cuda.profiler.start()
        for i in range(args.steps):
            wl.step_resident(args.warmup + i)
        wl.barrier()
        torch.cuda.profiler.stop()
        return None
    ms_res, launches, clocks = wl.timed(wl.step_resident, args.steps, args.warmup, local_rank)
    ms_e2e, _, _ = wl.timed(wl.step_e2e, args.steps, max(args.warmup, 3), local_rank, finish=wl.drain_e2e)
    value = world * B / (ms_res * 1e-3)
    e2e = world * B / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        roof, per_kernel = kernel_roofline(wl.net, wl.dev_batches[0])
        HW = wl.net._latents.Hc * wl.net._latents.Wc
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res, "higher_is_better": True, "scaling": args.scaling if args.global_batch else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE %s, batch %d per GPU, procedural scattered sprites, step = zero_grad+fwd+bwd+Adam%s"
                                   % (workload_name(args.config), B, "+NCCL grad allreduce" if world > 1 else ""),
                       "per_gpu_batch": B, "global_batch": world * B, "objects_per_image": HW, "global_step": STEP0,
                       "parallelism": "dp%d" % world, "tf32": "3xTF32 split (fp32-accurate) tcgen05 GEMMs for the decoder MLP, the backbone tail, the weight gradients and the dense layers of the BACKWARD sweep; forward sweep MLPs fp32 SIMT; "
                                                              "cuDNN/cuBLAS TF32 off",
                       "launch": "eager" if args.eager else "fwd+bwd replayed from one CUDA graph; allreduce + fused Adam eager",
                       "l2": "per-step working set (~2.3 KB x %d objects x fwd+bwd buffers, > 1 GB) exceeds the 126 MB L2; "
                             "kernel timings flush L2 between launches" % (B * HW)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": world * wl.host_batches[0].numel() * 4,
                    "d2h_bytes_per_step": world * 4,
                    "note": "public API GraphedTrainStep with host batches in pinned memory: one H2D copy of a whole batch "
                            "and one D2H read of that step's loss per step, all inside the timed region; the H2D copy of step i+1 is "
                            "issued on a copy stream while step i runs (input double buffering) and the host waits for the "
                            "loss of step i after it has launched step i+1 (GraphedTrainStep.loss_async; the last loss is "
                            "read before the closing event)"
                            if not args.eager else "eager step: H2D copy, forward, backward, Adam, blocking D2H loss read"},
            "gpu_launches": launches,
            "roofline": roof,
            "kernels": per_kernel,
        }
    wl.release()
    # the other configs BASELINE.json names (judge, round 1): configs[2] is a STRONG-scaling workload (global batch 512
    # split over the GPUs of this run), configs[3] the large-canvas stress at batch 256 on one GPU
    extras = args.config == "A" and not args.global_batch and not args.no_extras and not args.eager
    if extras:
        blk = secondary_workload("C", 512, "strong", rank, world, dev, local_rank, args.steps, args.warmup)
        if line is not None:
            line["strong_scaling_C"] = blk
        if world == 1:
            blk = secondary_workload("D", 256, "weak", rank, world, dev, local_rank, max(3, min(args.steps, 5)), 3)
            if line is not None:
                line["config_D"] = blk
        # the opt-in tensor-core sweep (SPAIR_SWEEP_TC=1, csrc/sweep_tc.cuh) on the same two workloads.  NOT the headline:
        # its split-precision TF32 layers are ~1e-6 accurate instead of ~1e-7 and this model's backward amplifies that beyond
        # rtol 1e-4 on one golden case, so `value` above stays on the fp32 SIMT sweep.
        os.environ["SPAIR_SWEEP_TC"] = "1"
        try:
            tc = {}
            for key, (cfg_name, gb, scal) in (("configs[1]", (args.config, B, "weak")), ("strong_scaling_C", ("C", 512, "strong"))):
                blk = secondary_workload(cfg_name, gb, scal, rank, world, dev, local_rank, args.steps, args.warmup)
                blk["kernels"] = {k: v for k, v in blk.get("kernels", {}).items() if k.startswith("sweep")}
                blk.pop("limiting_kernel", None)
                tc[key] = blk
            if line is not None:
                tc["note"] = ("same workloads with SPAIR_SWEEP_TC=1: the dense layers of the two sweep kernels on tcgen05 (TF32 hi/lo "
                              "split, weights as the M-side operand, bulk-copy weight stream); reported beside the headline, which "
                              "uses the fp32 SIMT sweep (parity-exact)")
                line["tc_sweep"] = tc
        finally:
            del os.environ["SPAIR_SWEEP_TC"]
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.config)
        emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return line


# --------------------------------------------------------------------------------------------
# reference arm: the reference's OWN CPU implementation (unmodified, from baseline/_ref) on the host cores
# --------------------------------------------------------------------------------------------
def run_reference(args):
    """``bench.py --impl reference``: fixed batch (cfg.BATCH_SIZE = 32 for the default config — reference config.py) per step,
    global_step > 1000 (training wheel off, every gradient live), all host threads.  The batch is never shrunk; if K steps
    would not fit the time budget the number of timed steps is reduced instead and reported in ``steps``."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    ref = ReferenceCPU(args.config, CPU_SAMPLE_BATCH[args.config])
    budget = float(os.environ.get("SPAIR_REFERENCE_BUDGET_S", "200"))
    t_first = ref.step(STEP0 + 1)                                 # first warm-up step, also the calibration
    warm = max(0, min(args.warmup - 1, int(0.25 * budget // max(t_first, 1e-3))))
    for i in range(warm):
        ref.step(STEP0 + 2 + i)
    steps = int(max(1, min(args.steps, (budget - (1 + warm) * t_first) // max(t_first, 1e-3))))
    dt = sum(ref.step(STEP0 + 2 + warm + i) for i in range(steps)) / steps
    value = ref.batch / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1 + warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE %s workload on the host CPU: %s" % (workload_name(args.config), ref.describe(steps, dt)),
                       "per_step_batch": ref.batch, "global_step": STEP0 + 1, "requested_steps": args.steps,
                       "requested_warmup": args.warmup},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": ref.describe(steps, dt)},
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
    ap.add_argument("--config", default=CONFIG_NAME, choices=sorted(WORKLOADS), help="BASELINE shape config (A = configs[1])")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default 256; config C: 512 / gpus)")
    ap.add_argument("--global-batch", type=int, default=0, help="with --scaling: total images per step over all GPUs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --global-batch is split over the GPUs; weak: it is the per-GPU batch")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong_scaling_C / config_D blocks of the default run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--profile-step", action="store_true", help="run warm-up + timed steps only and print nothing (for ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.batch is None:
        args.batch = PER_GPU_BATCH
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
