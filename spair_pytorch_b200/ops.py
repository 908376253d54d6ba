"""Autograd wrappers and the wavefront sweep over the sm_100a kernels (host orchestration).

Nothing here computes on the CPU: every numerical step is a call into ``libspair_b200.so`` (``kernels.py``) — including
the dense contractions, which run on the tcgen05 GEMM (``K.gemm3x``); ``torch.mm/addmm`` (cuBLAS) is only reached with
``SPAIR_NO_TC_GEMM=1`` or on the per-wavefront fallback path.  The backward passes are written out by hand instead of taped:

  * ``GlimpseFunction`` / ``PasteFunction`` — ``stn()`` both directions (reference modules.py:216-273);
  * ``RenderFunction``  — fused decode + inverse warp + compositing (+ BCE) (models.py:481-547);
  * ``KLFunction``      — masked Normal KLs + count-prior scan (models.py:169-262);
  * ``CellSweepFunction`` — the whole autoregressive cell loop (models.py:68-117) as Wc+2(Hc-1)
    wavefronts.  Activations of all cells live in wavefront-major row buffers
    ``[HW*B, width]`` so every per-wavefront GEMM works on a contiguous row slice, the heads write
    their results straight into the next network's input columns, and the weight gradients are
    13 large GEMMs over ALL rows at the end of backward instead of 121 x 13 small ones.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import kernels as K
from .schedule import WavefrontSchedule

# bench.py sets this to a dict to have the two persistent sweep launches bracketed with CUDA events on the launching
# stream: {"fwd": (start, end), "bwd": (start, end)}.  None (the default) records nothing.
SWEEP_EVENTS: Optional[dict] = None

# SPAIR_NO_TC_GEMM=1 keeps the decoder and the weight gradients on cuBLAS fp32 (the round-1 path; for A/B timing and as the
# reference the tensor-core path is tested against)
import os as _os
USE_TENSOR_CORE_GEMM = "SPAIR_NO_TC_GEMM" not in _os.environ
# k x k convolutions of the backbone tail as implicit GEMMs (TMA im2col loads in the GEMM's producer warp) instead of a
# materialised patch matrix; SPAIR_EXPLICIT_IM2COL=1 restores the explicit path (csrc/conv.cu im2col_nhwc)
IMPLICIT_CONV = "SPAIR_EXPLICIT_IM2COL" not in _os.environ


def sweep_tc_choice(rows_per_cta: int, mode: Optional[str] = None):
    """(forward, backward): which of the two fused sweeps run their dense layers on the tensor cores (csrc/sweep_tc.cuh).
    ``mode`` = SPAIR_SWEEP_TC: "auto" (default) -> the backward sweep when a CTA has >= 12 rows per wavefront (the cost of
    the tensor-core layers does not depend on the rows, that of the SIMT layers does: 16 rows 2.39 vs 2.65 ms, 8 rows 2.48 vs
    2.27 ms); "bwd" -> the backward always; "1" -> both (the forward is then ~1e-6 accurate, not parity-exact, DESIGN.md
    section 5); "0" -> neither."""
    if mode is None:
        mode = _os.environ.get("SPAIR_SWEEP_TC", "auto")
    if mode not in ("auto", "bwd", "1", "0"):
        raise ValueError("SPAIR_SWEEP_TC must be auto, bwd, 1 or 0, got %r" % mode)
    return mode == "1", mode in ("1", "bwd") or (mode == "auto" and rows_per_cta >= 12)


def _timed_launch(name, fn, *args, **kwargs):
    if SWEEP_EVENTS is None:
        return fn(*args, **kwargs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(*args, **kwargs)
    e1.record()
    SWEEP_EVENTS[name] = (e0, e1)


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.contiguous()


def _rows16(rows: int, width: int, device, zero=False) -> torch.Tensor:
    """[rows, width] fp32 view whose row pitch is a multiple of 4 floats (16 bytes), the unit TMA addresses rows in."""
    pitch = (width + 3) & ~3
    buf = (torch.zeros if zero else torch.empty)(rows, pitch, device=device, dtype=torch.float32)
    return buf if pitch == width else buf[:, :width]


def _tma_rows(t: torch.Tensor) -> torch.Tensor:
    """``t`` itself when the tcgen05 GEMM can load it (K.gemm_supported), else a copy with a 16-byte row pitch."""
    if K.gemm_supported(t):
        return t
    out = _rows16(t.shape[0], t.shape[1], t.device)
    out.copy_(t)
    return out


# ==========================================================================================
# stn: glimpse (forward direction) and paste (inverse direction)
# ==========================================================================================
class GlimpseFunction(torch.autograd.Function):
    """``stn(image, z_where, [Gh, Gw])``: image [n,C,Ih,Iw], z_where [n,4] -> [n,C,Gh,Gw]."""

    @staticmethod
    def forward(ctx, image, z_where, Gh, Gw):
        image, z_where = _c(image), _c(z_where)
        n, C = image.shape[0], image.shape[1]
        out = torch.empty(n, C * Gh * Gw, device=image.device, dtype=torch.float32)
        K.glimpse_fwd(image, z_where, None, n, 1, Gh, Gw, out)
        ctx.save_for_backward(image, z_where)
        ctx.dims = (Gh, Gw)
        return out.view(n, C, Gh, Gw)

    @staticmethod
    def backward(ctx, d_out):
        image, z_where = ctx.saved_tensors
        Gh, Gw = ctx.dims
        n = image.shape[0]
        d_zw = torch.empty(n, 4, device=image.device, dtype=torch.float32)
        d_image = torch.zeros_like(image) if ctx.needs_input_grad[0] else None
        K.glimpse_bwd(image, z_where, None, n, 1, Gh, Gw, _c(d_out).view(n, -1), d_zw, d_image)
        return d_image, d_zw, None, None


class PasteFunction(torch.autograd.Function):
    """``stn(image, z_where, [Oh, Ow], inverse=True)``: image [n,C,Gh,Gw] -> [n,C,Oh,Ow]."""

    @staticmethod
    def forward(ctx, image, z_where, Oh, Ow):
        image, z_where = _c(image), _c(z_where)
        n, C = image.shape[0], image.shape[1]
        out = torch.empty(n, C, Oh, Ow, device=image.device, dtype=torch.float32)
        K.paste_fwd(image, z_where, Oh, Ow, out)
        ctx.save_for_backward(image, z_where)
        ctx.dims = (Oh, Ow)
        return out

    @staticmethod
    def backward(ctx, d_out):
        image, z_where = ctx.saved_tensors
        Oh, Ow = ctx.dims
        d_zw = torch.empty(image.shape[0], 4, device=image.device, dtype=torch.float32)
        d_image = torch.zeros_like(image)
        K.paste_bwd(image, z_where, Oh, Ow, _c(d_out), d_image, d_zw)
        return d_image, d_zw, None, None


# ==========================================================================================
# renderer
# ==========================================================================================
class RenderFunction(torch.autograd.Function):
    """(logits [N,G,G,C+1], z_where [N,4], z_depth [N], z_pres [N], target [B,C,Ih,Iw] | None)
    -> (recon [B,C,Ih,Iw], bce_sum scalar).  N = B*HW, objects of one image contiguous."""

    @staticmethod
    def forward(ctx, logits, z_where, z_depth, z_pres, target, B, HW, C, G, Ih, Iw, scales, decoded=False):
        """``decoded``: ``logits`` are the texel records of ``DecoderFunction`` (already through the sigmoids).  The
        gradient returned for them is then the gradient wrt the RAW logits (the sigmoid derivative is applied here from
        the records, where it costs nothing), which is what ``DecoderFunction.backward`` expects."""
        logits, z_where, z_depth, z_pres = _c(logits), _c(z_where), _c(z_depth), _c(z_pres)
        dev = logits.device
        recon = torch.empty(B, C, Ih, Iw, device=dev, dtype=torch.float32)
        denom = torch.empty(B, Ih, Iw, device=dev, dtype=torch.float32)
        partial = None
        if target is not None:
            target = _c(target)
            partial = torch.empty(K.render_num_tiles(B, Ih, Iw), device=dev, dtype=torch.float32)
        K.render_fwd(logits, z_where, z_depth, z_pres, B, HW, C, G, Ih, Iw, scales, recon, denom, target, partial, decoded)
        bce = partial.sum() if partial is not None else torch.zeros((), device=dev)
        ctx.save_for_backward(logits, z_where, z_depth, z_pres, recon, denom, target)
        ctx.meta = (B, HW, C, G, Ih, Iw, scales, decoded)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(denom)
        return recon, bce, denom

    @staticmethod
    def backward(ctx, d_recon, d_bce, _d_denom):
        logits, z_where, z_depth, z_pres, recon, denom, target = ctx.saved_tensors
        B, HW, C, G, Ih, Iw, scales, decoded = ctx.meta
        dev = logits.device
        if d_bce is None and d_recon is None:
            return (None,) * 13
        use_target = target if d_bce is not None else None
        bce_scale = _c(d_bce.reshape(1).float()) if d_bce is not None else None
        gs = torch.empty(B, C + 1, Ih, Iw, device=dev, dtype=torch.float32)
        d_logits = torch.empty_like(logits)
        d_zw = torch.empty_like(z_where)
        d_depth = torch.empty_like(z_depth)
        d_pres = torch.empty_like(z_pres)
        K.render_bwd(logits, z_where, z_depth, z_pres, B, HW, C, G, Ih, Iw, scales, recon, denom, _c(d_recon), use_target,
                     bce_scale, gs, d_logits, d_zw, d_depth, d_pres, decoded)
        return (d_logits, d_zw, d_depth, d_pres) + (None,) * 9


# ==========================================================================================
# KL
# ==========================================================================================
class KLFunction(torch.autograd.Function):
    """(dmean, dstd [B,HW,D], pres [B,HW]) -> kl_sums [B,7] (cy, cx, height, width, attr, depth, pres)
    plus the non-differentiable kl_map [B,HW,D+1] for inspection."""

    @staticmethod
    def forward(ctx, dmean, dstd, pres, prior_mean, prior_std, count_dist0, A, side=None, sync=None):
        """``side`` (a CUDA stream) + ``sync`` (dict): the kernel — a serial scan over the cells, one CTA per image, i.e.
        pure latency on B of the 148 SMs — is launched on ``side`` so that it overlaps the decoder GEMMs and the renderer;
        ``sync['done']`` is the event the consumer of the outputs must wait for.  Outputs are allocated on the current
        stream (they belong to its allocator pool); CUDA-graph capture records the fork / join as a parallel branch."""
        dmean, dstd, pres = _c(dmean), _c(dstd), _c(pres)
        B, HW, D = dmean.shape
        dev = dmean.device
        kl_map = torch.empty(B, HW, D + 1, device=dev, dtype=torch.float32)
        p_z = torch.empty(B, HW, device=dev, dtype=torch.float32)
        sums = torch.empty(B, 7, device=dev, dtype=torch.float32)
        if side is None:
            K.kl_fwd(dmean, dstd, pres, prior_mean, prior_std, count_dist0, B, HW, A, kl_map, p_z, sums)
        else:
            fork = torch.cuda.Event()
            fork.record(torch.cuda.current_stream(dev))
            side.wait_event(fork)
            with torch.cuda.stream(side):
                K.kl_fwd(dmean, dstd, pres, prior_mean, prior_std, count_dist0, B, HW, A, kl_map, p_z, sums)
                sync["done"] = torch.cuda.Event()
                sync["done"].record(side)
        ctx.save_for_backward(dmean, dstd, pres, prior_mean, prior_std, kl_map, p_z)
        ctx.A = A
        ctx.mark_non_differentiable(kl_map)
        return sums, kl_map

    @staticmethod
    def backward(ctx, d_sums, _d_map):
        dmean, dstd, pres, prior_mean, prior_std, kl_map, p_z = ctx.saved_tensors
        B, HW, _ = dmean.shape
        d_dmean, d_dstd, d_pres = torch.empty_like(dmean), torch.empty_like(dstd), torch.empty_like(pres)
        K.kl_bwd(dmean, dstd, pres, prior_mean, prior_std, kl_map, p_z, _c(d_sums), B, HW, ctx.A, d_dmean, d_dstd, d_pres)
        return d_dmean, d_dstd, d_pres, None, None, None, None, None, None


# ==========================================================================================
# the autoregressive cell sweep
# ==========================================================================================
@dataclass
class SweepPlan:
    """Static description of one model's sweep (built once per model / device)."""
    schedule: WavefrontSchedule
    geom: "K.BoxGeom"
    B_hint: int
    F: int                 # backbone features
    A: int                 # attribute dims
    P: int                 # passthrough features
    C: int
    Ih: int
    Iw: int
    G: int
    fused_forward: bool = True         # one persistent launch for the forward sweep when the shape allows it
    fused_backward: bool = True        # same for the reverse sweep (needs the fused forward's bookkeeping)
    order_dev: torch.Tensor = None     # int32 [HW] cells in wavefront-major order
    starts_dev: torch.Tensor = None    # int32 [T+1]
    wf_pos_dev: torch.Tensor = None    # int32 [HW]
    missing_dev: torch.Tensor = None   # float [HW, n_nb] by position
    gather_index: torch.Tensor = None  # int64 [HW] wf_pos as gather index
    n_hidden: dict = field(default_factory=dict)

    @property
    def E(self):
        return self.A + 6

    @property
    def ctx_dim(self):
        return len(self.schedule.offsets) * self.E

    def side_streams(self, device, n):
        """Streams the backward pass forks the independent weight-gradient GEMMs to (created once per plan)."""
        pool = self.__dict__.setdefault("_side_streams", [])
        while len(pool) < n:
            pool.append(torch.cuda.Stream(device=device))
        return pool[:n]

    def to(self, device):
        s = self.schedule
        self.order_dev = torch.from_numpy(np.ascontiguousarray(s.order)).to(device=device, dtype=torch.int32)
        self.wf_pos_dev = torch.from_numpy(np.ascontiguousarray(s.wf_pos)).to(device=device, dtype=torch.int32)
        self.starts_dev = torch.from_numpy(np.ascontiguousarray(s.starts)).to(device=device, dtype=torch.int32)
        self.missing_dev = torch.from_numpy(s.missing.astype(np.float32)).to(device)
        self.gather_index = self.wf_pos_dev.long()
        return self


class _ManualMLP:
    """cuBLAS MLP over row slices of preallocated buffers, with a hand-written backward.
    ``weights``/``biases``: hidden layers then ONE output layer (multi-head networks pass the
    concatenation of their heads)."""

    def __init__(self, weights: List[torch.Tensor], biases: List[torch.Tensor], rows: int, device):
        self.W, self.b = weights, biases
        self.widths = [w.shape[0] for w in weights]
        self.n_in = weights[0].shape[1]
        self.X = _rows16(rows, self.n_in, device)          # 16-byte row pitch: the weight-gradient GEMM loads it with TMA
        self.H = [torch.empty(rows, w, device=device, dtype=torch.float32) for w in self.widths[:-1]]
        self.Y = torch.empty(rows, self.widths[-1], device=device, dtype=torch.float32)
        self.dX = self.dH = self.dY = None

    def forward(self, r0, r1):
        inp = self.X[r0:r1]
        for l, h in enumerate(self.H):
            # bias + ReLU fused into the GEMM epilogue (cuBLASLt) instead of a separate elementwise launch
            torch._addmm_activation(self.b[l], inp, self.W[l].t(), out=h[r0:r1])
            inp = h[r0:r1]
        torch.addmm(self.b[-1], inp, self.W[-1].t(), out=self.Y[r0:r1])

    def alloc_grads(self):
        self.dX = _rows16(self.X.shape[0], self.n_in, self.X.device)
        self.dH = [torch.empty_like(h) for h in self.H]
        self.dY = torch.empty_like(self.Y)

    def backward_dx(self, r0, r1):
        g = self.dY[r0:r1]
        for l in range(len(self.H) - 1, -1, -1):
            torch.mm(g, self.W[l + 1], out=self.dH[l][r0:r1])
            K.relu_bwd(self.dH[l][r0:r1], self.H[l][r0:r1])
            g = self.dH[l][r0:r1]
        torch.mm(g, self.W[0], out=self.dX[r0:r1])

    def alloc_weight_grads(self):
        """Outputs of ``weight_grads`` (allocated by the caller on ITS stream, filled on a side stream)."""
        dev = self.X.device
        self.dW = [torch.empty_like(w) for w in self.W]
        self.db = [torch.empty(w.shape[0], device=dev, dtype=torch.float32) for w in self.W]

    def weight_grads(self):
        """One GEMM per layer over all rows: dW = dY^T X, db = column sums."""
        inputs = [self.X] + self.H
        grads = self.dH + [self.dY]
        for g, i, dW, db in zip(grads, inputs, self.dW, self.db):
            if USE_TENSOR_CORE_GEMM:
                # dW[n_out, n_in] = g^T i, both operands MN-major as stored (a head output of 102 or 1 columns is first
                # copied to a 16-byte row pitch: 12 MB, against a 30976-row reduction)
                K.gemm3x(_tma_rows(g), False, _tma_rows(i), False, dW)
                K.relu_bwd_colsum(g, None, db)
            else:
                torch.mm(g.t(), i, out=dW)
                torch.sum(g, 0, out=db)
        return self.dW, self.db


class DecoderFunction(torch.autograd.Function):
    """``object_decoder`` (reference models.py:165,474-481: Linear(A,128) ReLU Linear(128,256) ReLU Linear(256, G*G*(C+1)))
    followed by the per-texel sigmoids of SPAIR._render (models.py:485-493), for all N = B*HW objects at once.

    The two wide layers run on the tcgen05 GEMM of csrc/gemm.cu (3xTF32, fp32 accuracy); the last one applies
    bias + scale + sigmoid in its epilogue, so the [N, G*G*(C+1)] logits (194 MB at the default config) never reach HBM:
    the output is the texel records ``RenderFunction(decoded=True)`` consumes.  Backward takes the gradient wrt the RAW
    logits (that is what ``RenderFunction.backward`` returns in decoded mode) and runs dgrad / wgrad on the same kernel.
    The first layer's K = A = 50 columns are 200-byte rows, which TMA cannot address: its input and weight are copied to
    zero-padded buffers with a 16-byte row pitch (6 MB), so all three layers run on the tensor cores."""

    @staticmethod
    def forward(ctx, attr, w0, b0, w1, b1, w2, b2, period, scales):
        attr = _c(attr)
        w0, b0, w1, b1, w2, b2 = (t.detach().contiguous() for t in (w0, b0, w1, b1, w2, b2))
        n, dev = attr.shape[0], attr.device
        # K = A = 50 columns are 200-byte rows: zero-padded copies with a 16-byte pitch (6 MB) put this layer on TMA too
        attr_p = _rows16(n, attr.shape[1], dev, zero=True)
        attr_p.copy_(attr)
        w0_p = _rows16(w0.shape[0], w0.shape[1], dev, zero=True)
        w0_p.copy_(w0)
        h0 = torch.empty(n, w0.shape[0], device=dev, dtype=torch.float32)
        K.gemm3x(attr_p, True, w0_p, True, h0, b0, epilogue=K.GEMM_EPI_RELU)
        # the two wide weights are split into their TF32 hi / lo planes once (one small launch each) and serve the forward
        # and the input-gradient GEMM: the kernels then only split the activation tiles
        w1s, w2s = K.SplitWeight(w1), K.SplitWeight(w2)
        h1 = torch.empty(n, w1.shape[0], device=dev, dtype=torch.float32)
        K.gemm3x(h0, True, w1s, True, h1, b1, epilogue=K.GEMM_EPI_RELU)
        texels = torch.empty(n, w2.shape[0], device=dev, dtype=torch.float32)
        K.gemm3x(h1, True, w2s, True, texels, b2, epilogue=K.GEMM_EPI_TEXEL, period=period, scales=scales)
        ctx.save_for_backward(attr_p, h0, h1, w0_p, w1, w2)
        ctx.split_weights = (w1s, w2s)
        return texels

    @staticmethod
    def backward(ctx, d_logits):
        attr, h0, h1, w0, w1, w2 = ctx.saved_tensors
        w1s, w2s = ctx.split_weights
        d_logits = _c(d_logits)
        n, dev = attr.shape[0], attr.device
        d_w2 = torch.empty_like(w2)
        K.gemm3x(d_logits, False, h1, False, d_w2)
        d_b2 = torch.empty(w2.shape[0], device=dev, dtype=torch.float32)
        K.relu_bwd_colsum(d_logits, None, d_b2)
        d_h1 = torch.empty_like(h1)
        K.gemm3x(d_logits, True, w2s, False, d_h1)
        d_b1 = torch.empty(w1.shape[0], device=dev, dtype=torch.float32)
        K.relu_bwd_colsum(d_h1, h1, d_b1)
        d_w1 = torch.empty_like(w1)
        K.gemm3x(d_h1, False, h0, False, d_w1)
        d_h0 = torch.empty_like(h0)
        K.gemm3x(d_h1, True, w1s, False, d_h0)
        d_b0 = torch.empty(w0.shape[0], device=dev, dtype=torch.float32)
        K.relu_bwd_colsum(d_h0, h0, d_b0)
        d_w0 = torch.empty_like(w0)
        K.gemm3x(d_h0, False, attr, False, d_w0)
        d_attr = None
        if ctx.needs_input_grad[0]:
            d_attr = torch.empty_like(attr)
            K.gemm3x(d_h0, True, w0, False, d_attr)
        return d_attr, d_w0, d_b0, d_w1, d_b1, d_w2, d_b2, None, None


class WideLinearFunction(torch.autograd.Function):
    """``F.linear(h, weight, bias)`` for the decoder's output layer (reference models.py:165,477: 256 -> G*G*(C+1), a
    [B*HW, 1568] result of 194 MB at the default config).  cuBLASLt applies the bias of an fp32 SIMT GEMM in a separate
    un-fused pass over the output (0.22 ms); here the output is pre-filled with the bias rows and the GEMM accumulates
    into it (beta = 1), which costs one 194 MB fill instead.  Same values as addmm up to the position of one rounding."""

    @staticmethod
    def forward(ctx, h, weight, bias):
        out = torch.empty(h.shape[0], weight.shape[0], device=h.device, dtype=torch.float32)
        K.broadcast_rows(bias.detach().contiguous(), h.shape[0], out)
        torch.addmm(out, h, weight.detach().t(), out=out)
        ctx.save_for_backward(h, weight)
        return out

    @staticmethod
    def backward(ctx, d_out):
        h, weight = ctx.saved_tensors
        d_out = d_out.contiguous()
        return d_out.mm(weight), d_out.t().mm(h), d_out.sum(0)


class StemConvFunction(torch.autograd.Function):
    """relu(conv2d(zero_pad(x), weight, bias, stride)) for the first backbone layer (reference modules.py:44-66,86-87) as one
    launch forward and one (+ a small fixed-order reduction) backward.  ``x`` gets no gradient (the model's image is a
    leaf without grad; callers that need d_x keep the library path)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride: int, pad_t: int, pad_l: int, Ho: int, Wo: int, channels_last: bool = False):
        x = x.contiguous()
        w = weight.detach().contiguous()
        # channels_last: [B,Ho,Wo,Cout], the layout ConvTailFunction works in (no transpose pass between the two)
        shape = (x.shape[0], Ho, Wo, w.shape[0]) if channels_last else (x.shape[0], w.shape[0], Ho, Wo)
        y = torch.empty(shape, device=x.device, dtype=torch.float32)
        K.stem_conv_fwd(x, w, bias.detach().contiguous(), stride, pad_t, pad_l, Ho, Wo, y, channels_last)
        ctx.save_for_backward(x, y)
        ctx.geom = (stride, pad_t, pad_l, tuple(w.shape), channels_last)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        stride, pad_t, pad_l, w_shape, channels_last = ctx.geom
        d_w = torch.empty(w_shape, device=x.device, dtype=torch.float32)
        d_b = torch.empty(w_shape[0], device=x.device, dtype=torch.float32)
        ws = K.stem_bwd_workspace(x.shape[1], w_shape[0], x.device)
        K.stem_conv_bwd(x, y, dy.contiguous(), w_shape, stride, pad_t, pad_l, ws, d_w, d_b, channels_last)
        return None, d_w, d_b, None, None, None, None, None, None


class ConvTailFunction(torch.autograd.Function):
    """The backbone after the stem (reference modules.py:44-66: [Conv2d(k, stride) + ReLU]* + a 1x1 ``conv_out`` without
    activation) as GEMMs on the tcgen05 kernel of csrc/gemm.cu, forward and backward.

    cuDNN has no fp32-accurate tensor-core convolution (TF32 is off for parity), so the library path runs these layers
    on the fp32 SIMT / FFT engines at ~20 TFLOP/s — 8 of the 16.4 ms of a config-B step.  Here the activations are kept
    channels-last, a 1x1 layer is a GEMM on the stored tensor, a k x k / stride s layer an implicit GEMM whose producer
    warp reads the patches with TMA im2col loads (forward and weight gradient; no patch matrix in HBM), bias + ReLU run in
    the GEMM epilogue, and the input gradient is dgrad on the same kernel plus the transposed patch gather (csrc/conv.cu).  Layers: padding 0, dilation 1, groups 1, square kernel, Cin % 4 == 0.

    forward(y0 (the stem's output: [B,C0,H0,W0] NCHW, or already channels-last [B,H0,W0,C0] with ``nhwc_in``), specs,
    nhwc_in, *weights_and_biases) -> feat [B,F,Hc,Wc] NCHW; ``specs`` = ((k, stride, relu), ...)."""

    @staticmethod
    def forward(ctx, y0, specs, nhwc_in, *params):
        if nhwc_in:
            x = y0.contiguous()                 # the stem wrote channels-last: no transpose pass
        else:
            B0, C0, H0, W0 = y0.shape
            x = torch.empty(B0, H0, W0, C0, device=y0.device, dtype=torch.float32)      # channels-last from here on
            K.transpose_batched(y0.contiguous().view(B0, C0, H0 * W0), x.view(B0, H0 * W0, C0))
        saved, shapes, split = [], [], []
        for li, (k, s, relu) in enumerate(specs):
            w, b = params[2 * li].detach(), params[2 * li + 1].detach().contiguous()
            Bn, H, W, Cin = x.shape
            Ho, Wo = (H - k) // s + 1, (W - k) // s + 1
            M = Bn * Ho * Wo
            wr = w.permute(0, 2, 3, 1).reshape(w.shape[0], k * k * Cin).contiguous()     # [Cout, (kh, kw, c)]
            wrs = K.SplitWeight(wr)           # TF32 hi / lo planes, once: forward and (1x1 layers) input gradient
            split.append(wrs)
            y = torch.empty(M, w.shape[0], device=x.device, dtype=torch.float32)
            if k == 1 and s == 1:
                a = x.view(M, Cin)
                K.gemm3x(a, True, wrs, True, y, b, epilogue=K.GEMM_EPI_RELU if relu else K.GEMM_EPI_NONE)
            elif IMPLICIT_CONV and K.conv_supported(x, k, s):
                a = x                                   # implicit GEMM: the producer warp reads the patches with TMA im2col
                K.conv_fwd(x, k, s, wrs, b, y, relu)
            else:
                a = torch.empty(M, k * k * Cin, device=x.device, dtype=torch.float32)
                K.im2col_nhwc(x, k, s, a)
                K.gemm3x(a, True, wrs, True, y, b, epilogue=K.GEMM_EPI_RELU if relu else K.GEMM_EPI_NONE)
            saved += [a, wr, y]
            shapes.append((Bn, H, W, Cin, Ho, Wo, tuple(w.shape)))
            x = y.view(Bn, Ho, Wo, w.shape[0])
        ctx.save_for_backward(*saved)
        ctx.specs, ctx.shapes, ctx.split, ctx.nhwc_in = specs, shapes, split, nhwc_in
        Bn, Ho, Wo, F = x.shape
        feat = torch.empty(Bn, F, Ho, Wo, device=x.device, dtype=torch.float32)
        K.transpose_batched(x.view(Bn, Ho * Wo, F), feat.view(Bn, F, Ho * Wo))
        return feat

    @staticmethod
    def backward(ctx, d_feat):
        specs, shapes, saved = ctx.specs, ctx.shapes, ctx.saved_tensors
        Bn, _, _, _, Ho, Wo, wshape = shapes[-1]
        dy = torch.empty(Bn * Ho * Wo, wshape[0], device=d_feat.device, dtype=torch.float32)
        K.transpose_batched(d_feat.contiguous().view(Bn, wshape[0], Ho * Wo), dy.view(Bn, Ho * Wo, wshape[0]))
        grads = [None] * (2 * len(specs))
        for li in range(len(specs) - 1, -1, -1):
            k, s, relu = specs[li]
            a, wr, y = saved[3 * li: 3 * li + 3]
            Bn, H, W, Cin, Ho, Wo, wshape = shapes[li]
            db = torch.empty(wshape[0], device=dy.device, dtype=torch.float32)
            K.relu_bwd_colsum(dy, y if relu else None, db)            # ReLU mask (in place) + bias gradient, one pass
            d_wr = torch.empty_like(wr)
            implicit = a.dim() == 4                     # forward ran as an implicit GEMM: `a` is the channels-last input
            if implicit:
                K.conv_wgrad(a, k, s, dy, d_wr)
            else:
                K.gemm3x(dy, False, a, False, d_wr)
            grads[2 * li] = d_wr.view(wshape[0], k, k, Cin).permute(0, 3, 1, 2).contiguous()
            grads[2 * li + 1] = db
            if li == 0 and not ctx.needs_input_grad[0]:
                return (None, None, None) + tuple(grads)
            if implicit and K.conv_dgrad_supported(k, s, wshape[0]):
                # transposed convolution as s*s implicit GEMMs over dy (no d_col matrix, no scatter pass)
                w4 = wr.view(wshape[0], k, k, Cin).permute(0, 3, 1, 2)
                dx = torch.empty(Bn, H, W, Cin, device=dy.device, dtype=torch.float32)
                K.conv_dgrad(dy, Bn, k, s, K.SplitWeight(K.pack_dgrad_weights(w4, s)), dx)
                dy = dx.view(Bn * H * W, Cin)
                continue
            da = torch.empty(dy.shape[0], wr.shape[1], device=dy.device, dtype=torch.float32)
            K.gemm3x(dy, True, ctx.split[li], False, da)
            if k == 1 and s == 1:
                dy = da
            else:
                dx = torch.empty(Bn, H, W, Cin, device=da.device, dtype=torch.float32)
                K.col2im_nhwc(da, k, s, dx)
                dy = dx.view(Bn * H * W, Cin)
        Bn, H, W, Cin = shapes[0][:4]
        if ctx.nhwc_in:
            return (dy.view(Bn, H, W, Cin), None, None) + tuple(grads)
        d_y0 = torch.empty(Bn, Cin, H, W, device=dy.device, dtype=torch.float32)
        K.transpose_batched(dy.view(Bn, H * W, Cin), d_y0.view(Bn, Cin, H * W))
        return (d_y0, None, None) + tuple(grads)


class CellSweepFunction(torch.autograd.Function):
    """forward(plan, x, feat, edge, eps_where, eps_attr, eps_depth, u_pres, wheel, *params)
    -> (z_where [B,HW,4], attr [B,HW,A], depth [B,HW], pres [B,HW], dmean [B,HW,D], dstd [B,HW,D], box [B,HW,4])

    ``params`` order: box_network (hidden W,b..., head0 W,b, head1 W,b), object_encoder (hidden..., out),
    z_network (hidden..., head0, head1), obj_network (hidden..., out)."""

    @staticmethod
    def forward(ctx, plan: SweepPlan, x, feat, edge, eps_where, eps_attr, eps_depth, u_pres, wheel, *params):
        s = plan.schedule
        dev = feat.device
        B = feat.shape[0]
        HW = s.Hc * s.Wc
        A, P, F, E, CTX, G = plan.A, plan.P, plan.F, plan.E, plan.ctx_dim, plan.G
        D = 4 + A + 1
        rows = HW * B
        x, feat, edge = _c(x.detach()), _c(feat.detach()), _c(edge.detach())
        params = [p.detach() for p in params]

        nh = plan.n_hidden
        it = iter(params)

        def take(n_hidden, n_heads):
            Ws, bs = [], []
            for _ in range(n_hidden):
                Ws.append(next(it)); bs.append(next(it))
            hw_, hb_ = [], []
            for _ in range(n_heads):
                hw_.append(next(it)); hb_.append(next(it))
            Ws.append(torch.cat(hw_, 0) if n_heads > 1 else hw_[0])
            bs.append(torch.cat(hb_, 0) if n_heads > 1 else hb_[0])
            return Ws, bs

        box_mlp = _ManualMLP(*take(nh["box"], 2), rows, dev)
        enc_mlp = _ManualMLP(*take(nh["enc"], 1), rows, dev)
        z_mlp = _ManualMLP(*take(nh["z"], 2), rows, dev)
        obj_mlp = _ManualMLP(*take(nh["obj"], 1), rows, dev)
        assert box_mlp.n_in == F + CTX and z_mlp.n_in == F + CTX + P + 4 + A and obj_mlp.n_in == z_mlp.n_in + 1
        assert enc_mlp.n_in == plan.C * G * G

        box = torch.empty(B, HW, 4, device=dev)
        z_where = torch.empty(B, HW, 4, device=dev)
        attr = torch.empty(B, HW, A, device=dev)
        depth = torch.empty(B, HW, device=dev)
        pres = torch.empty(B, HW, device=dev)
        dmean = torch.empty(B, HW, D, device=dev)
        dstd = torch.empty(B, HW, D, device=dev)
        c_pt, c_box, c_attr, c_depth = F + CTX, F + CTX + P, F + CTX + P + 4, F + CTX + P + 4 + A

        mlps = (box_mlp, enc_mlp, z_mlp, obj_mlp)
        max_rows = K.sweep_max_rows() if plan.fused_forward else 0
        # every precondition of spair_sweep_fwd AND spair_sweep_bwd (csrc/sweep.cu), so a shape the kernels would reject
        # takes the per-wavefront path from the start instead of failing in the middle of backward
        fused = (max_rows >= s.max_cells and all(len(m.H) == 2 and max(m.widths) <= 256 for m in mlps) and G <= 64
                 and A + 6 <= 64 and len(s.offsets) <= K.MAX_NEIGHBOURS)
        if fused:
            # ONE persistent launch: a CTA owns `ipc` images and walks all wavefronts (csrc/sweep.cu)
            # images per CTA: two share one pass over the weight stream when there are more images than SMs; below that
            # (strong scaling: 64-128 images per GPU) one image per CTA keeps twice as many SMs busy
            ipc = max(1, min(2, max_rows // s.max_cells)) if B > K.NUM_SMS else 1
            dims = K.SweepDims(B=B, HW=HW, Hc=s.Hc, Wc=s.Wc, F=F, A=A, P=P, C=plan.C, Ih=plan.Ih, Iw=plan.Iw, G=G, ipc=ipc,
                               n_wavefronts=s.n_wavefronts, max_cells=s.max_cells, n_nb=len(s.offsets))
            # Dense layers of the sweeps on the tensor cores (csrc/sweep_tc.cuh; split-precision TF32: ~1e-6 relative instead
            # of ~1e-7).  Default: the BACKWARD sweep only — gradients tolerate 1e-6, whereas a 1e-6 error in the forward
            # latents is amplified by this model's loss (BCE gradients ~ 1 / recon) beyond rtol 1e-4 on one golden case
            # (DESIGN.md section 5), so the forward stays on the fp32 SIMT layers.  The tensor-core layers cost the same for
            # 1 or 16 rows (MMA issue interval) while the SIMT layers scale with the rows, so the default ("auto") takes
            # them only when a CTA has >= 12 rows per wavefront (measured: 16 rows 2.39 vs 2.65 ms, 8 rows 2.48 vs 2.27 ms).
            # SPAIR_SWEEP_TC=1: both sweeps; =bwd: backward always; =0: neither.
            tc_fwd, tc_bwd = sweep_tc_choice(s.max_cells * ipc)
            weights = [w for m in mlps for w in m.W]
            packed = None if (tc_fwd and tc_bwd) else K.PackedSweepWeights(weights)     # one launch each; kept for backward
            packed_tc = K.PackedSweepWeightsTC(weights, forward=tc_fwd, backward=tc_bwd) if (tc_fwd or tc_bwd) else None
            pk = packed_tc if tc_fwd else packed
            descs = [K.sweep_mlp_desc(pk, 3 * i, m.b, m.X, m.H[0], m.H[1], m.Y) for i, m in enumerate(mlps)]
            _timed_launch("fwd", K.sweep_fwd, dims, plan.order_dev, plan.starts_dev, s.offsets, x, feat, edge, eps_where,
                          eps_attr, eps_depth, u_pres, plan.geom, descs, box, z_where, attr, depth, pres, dmean, dstd,
                          tc_stream=packed_tc.fwd if tc_fwd else None)
            packed = packed_tc if tc_bwd else packed

        for t in range(s.n_wavefronts if not fused else 0):
            c0, c1 = int(s.starts[t]), int(s.starts[t + 1])
            r0, r1 = c0 * B, c1 * B
            cells = plan.order_dev[c0:c1]
            Xb, Xz, Xo = box_mlp.X[r0:r1], z_mlp.X[r0:r1], obj_mlp.X[r0:r1]
            K.context_gather_fwd(feat, box, attr, depth, pres, edge, cells, s.offsets, (Xb, Xz, Xo))
            box_mlp.forward(r0, r1)
            K.box_head_fwd(box_mlp.Y[r0:r1], eps_where, cells, B, HW, s.Wc, plan.geom, box, z_where, dmean, dstd,
                           (Xz[:, c_box:], Xo[:, c_box:]), P, Xz[:, c_pt:])
            K.glimpse_fwd(x, z_where, cells, B, HW, G, G, enc_mlp.X[r0:r1])
            enc_mlp.forward(r0, r1)
            K.normal_head_fwd(enc_mlp.Y[r0:r1], A, eps_attr, cells, B, HW, 0, 1.0, attr, dmean[..., 4:], dstd[..., 4:], D,
                              (Xz[:, c_attr:], Xo[:, c_attr:]), 0, None)
            z_mlp.forward(r0, r1)
            K.normal_head_fwd(z_mlp.Y[r0:r1], 1, eps_depth, cells, B, HW, 1, 4.0, depth, dmean[..., D - 1:],
                              dstd[..., D - 1:], D, (Xo[:, c_depth:],), P, Xo[:, c_pt:])
            obj_mlp.forward(r0, r1)
            K.pres_head_fwd(obj_mlp.Y[r0:r1], u_pres, cells, B, HW, pres)

        ctx.fused_dims = dims if fused else None
        ctx.packed_weights = packed if fused else None
        ctx.plan = plan
        plan.last_mlps = (box_mlp, enc_mlp, z_mlp, obj_mlp)      # inspection hook for tests (no copy)
        ctx.mlps = (box_mlp, enc_mlp, z_mlp, obj_mlp)
        ctx.n_params = len(params)
        # the training-wheel scalar lives in a persistent device buffer that the next prepare_step() overwrites: backward
        # must see THIS forward's value, so it is snapshotted; everything backward reads goes through save_for_backward so
        # autograd's version counters catch an in-place update (optimizer step) between forward and backward
        wheel = wheel.clone()
        ctx.save_for_backward(z_where, x, eps_where, eps_attr, eps_depth, u_pres, wheel, *params)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(box)
        return z_where, attr, depth, pres, dmean, dstd, box

    @staticmethod
    def backward(ctx, d_zw, d_attr, d_depth, d_pres, d_dmean, d_dstd, _d_box):
        plan: SweepPlan = ctx.plan
        s = plan.schedule
        z_where, x, eps_where, eps_attr, eps_depth, u_pres, wheel = ctx.saved_tensors[:7]
        box_mlp, enc_mlp, z_mlp, obj_mlp = ctx.mlps
        dev = z_where.device
        B = z_where.shape[0]
        HW = s.Hc * s.Wc
        A, P, F, E, CTX, G = plan.A, plan.P, plan.F, plan.E, plan.ctx_dim, plan.G
        D = 4 + A + 1
        d_zw, d_attr, d_depth, d_pres, d_dmean, d_dstd = (_c(t) for t in (d_zw, d_attr, d_depth, d_pres, d_dmean, d_dstd))
        if (d_dmean is None) != (d_dstd is None):
            d_dmean = torch.zeros(B, HW, D, device=dev) if d_dmean is None else d_dmean
            d_dstd = torch.zeros(B, HW, D, device=dev) if d_dstd is None else d_dstd
        for m in (box_mlp, enc_mlp, z_mlp, obj_mlp):
            m.alloc_grads()
        n_max = s.max_cells * B
        d_cell = torch.empty(n_max, E, device=dev)
        d_zw_local = torch.empty(n_max, 4, device=dev)
        c_pt, c_box, c_attr, c_depth = F + CTX, F + CTX + P, F + CTX + P + 4, F + CTX + P + 4 + A
        dm_attr = None if d_dmean is None else d_dmean[..., 4:]
        ds_attr = None if d_dstd is None else d_dstd[..., 4:]
        dm_depth = None if d_dmean is None else d_dmean[..., D - 1:]
        ds_depth = None if d_dstd is None else d_dstd[..., D - 1:]

        fused = ctx.fused_dims is not None and plan.fused_backward
        if fused:
            # ONE persistent launch for the whole reverse sweep (csrc/sweep.cu: sweep_bwd_kernel)
            packed = ctx.packed_weights
            descs = [K.sweep_mlp_bwd_desc(packed, 3 * i, m.H[0], m.H[1], m.Y, m.dX, m.dH[0], m.dH[1], m.dY)
                     for i, m in enumerate((box_mlp, enc_mlp, z_mlp, obj_mlp))]
            _timed_launch("bwd", K.sweep_bwd, ctx.fused_dims, plan.order_dev, plan.starts_dev, plan.wf_pos_dev, s.offsets, x,
                          z_where, eps_where, eps_attr, eps_depth, u_pres, wheel, plan.geom, descs, d_zw, d_attr, d_depth,
                          d_pres, d_dmean, d_dstd,
                          tc_stream=packed.bwd if isinstance(packed, K.PackedSweepWeightsTC) else None)

        for t in range(s.n_wavefronts - 1 if not fused else -1, -1, -1):
            c0, c1 = int(s.starts[t]), int(s.starts[t + 1])
            r0, r1 = c0 * B, c1 * B
            n = r1 - r0
            cells = plan.order_dev[c0:c1]
            dXz, dXo = z_mlp.dX[r0:r1], obj_mlp.dX[r0:r1]
            dc = d_cell[:n]
            # gradient arriving through the lateral context of later cells (models.py:106 -> 73)
            K.context_grad_gather((box_mlp.dX, z_mlp.dX, obj_mlp.dX), F, cells, plan.wf_pos_dev, s.offsets, B, s.Hc, s.Wc,
                                  A, dc)
            # z_pres
            K.pres_head_bwd(obj_mlp.Y[r0:r1], u_pres, cells, B, HW, wheel, dc[:, E - 1:], d_pres, obj_mlp.dY[r0:r1])
            obj_mlp.backward_dx(r0, r1)
            # z_depth
            K.normal_head_bwd(z_mlp.Y[r0:r1], 1, eps_depth, cells, B, HW, 1, 4.0, wheel,
                              (dXo[:, c_depth:], dc[:, E - 2:E - 1]), d_depth, dm_depth, ds_depth, D,
                              P, dXo[:, c_pt:], z_mlp.dY[r0:r1])
            z_mlp.backward_dx(r0, r1)
            # z_what
            K.normal_head_bwd(enc_mlp.Y[r0:r1], A, eps_attr, cells, B, HW, 0, 1.0, None,
                              (dXz[:, c_attr:], dXo[:, c_attr:], dc[:, 4:4 + A]), d_attr, dm_attr, ds_attr, D,
                              0, None, enc_mlp.dY[r0:r1])
            enc_mlp.backward_dx(r0, r1)
            K.glimpse_bwd(x, z_where, cells, B, HW, G, G, enc_mlp.dX[r0:r1], d_zw_local[:n], None)
            # z_where
            K.box_head_bwd(box_mlp.Y[r0:r1], eps_where, cells, B, HW, s.Wc, plan.geom, wheel,
                           (dXz[:, c_box:], dXo[:, c_box:], dc[:, 0:4]), d_zw_local[:n], d_zw, d_dmean, d_dstd, D,
                           P, dXz[:, c_pt:], box_mlp.dY[r0:r1])
            box_mlp.backward_dx(r0, r1)

        # ---- after the sweep: the gradients of the shared inputs (what the backbone's backward waits for) on this stream, and,
        # at the same time on four side streams, the weight gradients: 12 skinny GEMMs (<= 256 x 784 outputs, reduction over
        # all HW*B rows), one stream per network (forked from / joined to the current stream with events, which CUDA-graph
        # capture records as parallel branches)
        shared = {}

        def shared_input_grads():
            d_in = box_mlp.dX[:, :F + CTX] + z_mlp.dX[:, :F + CTX] + obj_mlp.dX[:, :F + CTX]     # [rows, F+CTX]
            d_in = d_in.view(HW, B, F + CTX)
            # backbone features: wavefront-major rows -> [B,F,Hc,Wc]
            shared["feat"] = d_in[:, :, :F].index_select(0, plan.gather_index).permute(1, 2, 0).reshape(B, F, s.Hc, s.Wc)
            # virtual edge element: every (cell, slot) whose neighbour is outside the grid (models.py:316)
            n_nb = len(s.offsets)
            shared["edge"] = (d_in[:, :, F:].reshape(HW, B, n_nb, E) * plan.missing_dev[:, None, :, None]).sum((0, 1, 2))

        all_mlps = (box_mlp, enc_mlp, z_mlp, obj_mlp)
        for mlp in all_mlps:
            mlp.alloc_weight_grads()
        K.parallel_branches(x.device, plan.side_streams, [shared_input_grads] + [mlp.weight_grads for mlp in all_mlps])
        d_feat, d_edge = shared["feat"], shared["edge"]

        grads = []
        for mlp, n_heads, head_sizes in ((box_mlp, 2, (8, P)), (enc_mlp, 1, None), (z_mlp, 2, (2, P)), (obj_mlp, 1, None)):
            dWs, dbs = mlp.dW, mlp.db
            for dW, db in zip(dWs[:-1], dbs[:-1]):
                grads += [dW, db]
            if n_heads == 1:
                grads += [dWs[-1], dbs[-1]]
            else:
                a = head_sizes[0]
                grads += [dWs[-1][:a], dbs[-1][:a], dWs[-1][a:], dbs[-1][a:]]
        assert len(grads) == ctx.n_params
        # plan, x, feat, edge, eps_where, eps_attr, eps_depth, u_pres, wheel, *params
        return (None, None, d_feat, d_edge, None, None, None, None, None) + tuple(grads)
