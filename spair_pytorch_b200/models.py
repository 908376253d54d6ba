"""SPAIR model — drop-in for the reference's ``spair.models.SPAIR`` (models.py:15-604).

Same constructor, ``forward(x, global_step=0) -> (loss, recon_x, z_where, z_pres)`` and
``state_dict`` keys/shapes as the reference, so ``train.py`` and reference checkpoints work
unchanged.  Building the model under the same ``torch.manual_seed`` gives bit-identical
parameters (modules are created in the reference's order and consume the same RNG stream).

What runs where (everything behind the C-ABI of include/spair_b200.h, forward and backward, ``ops.py``):
  * the whole cell loop — lateral context, the four per-cell MLPs, latent heads (reparameterisation, relaxed
    Bernoulli, box decode), glimpse extraction — one persistent sm_100a kernel per direction (csrc/sweep.cu);
  * every other dense contraction — the decoder MLP (with the texel sigmoids in its epilogue), the weight gradients
    of all MLPs, the backbone convolutions after the stem — the tcgen05 GEMM of csrc/gemm.cu (3xTF32 = fp32 accuracy);
  * the fused renderer with BCE, the KL terms with the count-prior scan, the backbone stem — their own kernels.
  No cuDNN / cuBLAS kernel runs in a step (``SPAIR_NO_TC_GEMM=1`` restores the library path for A/B comparisons).
  There is no CPU path: tensors must be CUDA fp32.

The per-cell loop of the reference (121 strictly sequential iterations, ~27k tiny ATen ops and
~1300 device->host syncs per forward) becomes Wc + 2(Hc-1) wavefronts with no host sync.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import torch
from torch import nn
from torch.distributions import Normal

from . import config as cfg
from . import debug_tools
from . import kernels as K
from . import ops
from .modules import *  # noqa: F401,F403  (the reference re-exports modules through models)
from .modules import Backbone, build_MLP, exponential_decay
from .schedule import build_schedule

KL_NAMES = ("cy_logit", "cx_logit", "height_logit", "width_logit", "attr", "depth_logit", "pres_dist")


class SPAIR(nn.Module):
    def __init__(self, image_shape, writer, device):
        super().__init__()
        self.image_shape = image_shape
        self.writer = writer
        self.B = 1
        self.device = device
        # (2L+1)^2 // 2 neighbour cells x [box(4), attr, depth, pres]   (reference models.py:26)
        self.context_dim = (cfg.N_LOOKBACK * 2 + 1) ** 2 // 2 * (4 + cfg.N_ATTRIBUTES + 1 + 1)
        self._cfg = SimpleNamespace(
            n_attr=cfg.N_ATTRIBUTES, n_pass=cfg.N_PASSTHROUGH_FEATURES, n_lookback=cfg.N_LOOKBACK,
            object_shape=tuple(cfg.OBJECT_SHAPE), anchor=tuple(cfg.ANCHORBOX_SHAPE), image_shape=tuple(cfg.INPUT_IMAGE_SHAPE),
            max_yx=cfg.MAX_YX, min_yx=cfg.MIN_YX, max_hw=cfg.MAX_HW, min_hw=cfg.MIN_HW,
            priors={k: tuple(v) for k, v in cfg.PRIORS.items()}, beta=cfg.VAE_BETA,
            wheel=dict(cfg.LATENT_VAR_TRAINING_WHEEL_PARAM), count_prior=dict(cfg.OBJ_PRES_COUNT_LOG_PRIOR),
            scales=(cfg.OBJ_LOGIT_SCALE, cfg.ALPHA_LOGIT_SCALE, cfg.ALPHA_LOGIT_BIAS))
        assert self._cfg.max_yx > self._cfg.min_yx and self._cfg.max_hw > self._cfg.min_hw
        self._build_networks()
        self._build_edge_element()
        self._build_indep_prior()
        self.pixels_per_cell = tuple(int(i) for i in self.backbone.grid_cell_size)
        self._plan = None
        self._noise = None
        self.kl_scale = 1.0          # 1/world_size under data parallelism (dp.py), 1 otherwise
        self._step_state = None      # per-device step scalars (training wheel, count prior)
        self._static_step = False    # True while a captured CUDA graph owns the step (graphed.py)
        self.dist_param, self.dist = {}, {}
        if "SPAIR_ALLOW_TF32" not in os.environ:   # fp32 parity with the reference's CPU path
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
        if "SPAIR_DETERMINISTIC" in os.environ:
            # the hand-written kernels are bitwise reproducible; this makes the cuDNN backbone reproducible too (its
            # default backward-filter algorithms use atomics) at the price of ~2x slower convolutions
            torch.backends.cudnn.deterministic = True
        print('model initialized')

    # ------------------------------------------------------------------------------------
    # construction (same module order / RNG consumption as reference models.py:133-167,273-290)
    # ------------------------------------------------------------------------------------
    def _build_networks(self):
        c = self._cfg
        self.backbone = Backbone(self.image_shape, cfg.N_BACKBONE_FEATURES)
        self.feature_space_dim = self.backbone.compute_output_shape()
        n_feat = self.feature_space_dim[0]
        self.box_network = build_MLP(n_feat + self.context_dim, multiple_output=(8, c.n_pass))
        G = c.object_shape[0]
        C = cfg.INPUT_IMAGE_SHAPE[0]
        self.object_encoder = build_MLP(G * G * C, 2 * c.n_attr, hidden_layers=[256, 128])
        z_in = 4 + c.n_attr + c.n_pass + self.context_dim + cfg.N_BACKBONE_FEATURES
        self.z_network = build_MLP(z_in, multiple_output=(2, c.n_pass))
        self.obj_network = build_MLP(z_in + 1, 1)
        self.object_decoder = build_MLP(c.n_attr, G * G * (C + 1), hidden_layers=[128, 256])
        self.attn = Self_Attn(55)

    def _build_edge_element(self):
        """Learnable stand-in for neighbours outside the grid (reference models.py:273-290)."""
        sizes = [4, cfg.N_ATTRIBUTES, 1, 1]
        loc, attr, depth, pres = torch.split(torch.randn(sum(sizes)), sizes)
        elem = torch.cat((torch.sigmoid(loc), attr, torch.sigmoid(depth), torch.sigmoid(pres)))
        self.register_parameter('virtual_edge_element', nn.Parameter(elem))

    def _build_indep_prior(self):
        self.kl_priors = {name: Normal(mean, std) for name, (mean, std) in cfg.PRIORS.items()}

    # ------------------------------------------------------------------------------------
    # noise: the kernels take the random draws as inputs (SURVEY.md §8 RNG row)
    # ------------------------------------------------------------------------------------
    def set_noise(self, eps_where=None, eps_attr=None, eps_depth=None, u_pres=None, keep=False):
        """Inject the draws of the NEXT forward (of every following forward if ``keep``), in the
        reference's layout: eps_where [B,4,Hc,Wc] in (cy, cx, height, width) order, eps_attr [B,A,Hc,Wc],
        eps_depth and u_pres [B,1,Hc,Wc].  Without injection every forward draws fresh noise on the device."""
        self._keep_noise = keep
        if eps_where is None:
            self._noise = None
            return

        def img_major(t):
            return t.permute(0, 2, 3, 1).reshape(t.shape[0], -1, t.shape[1]).contiguous().float()

        self._noise = (img_major(eps_where), img_major(eps_attr), img_major(eps_depth).squeeze(-1).contiguous(),
                       img_major(u_pres).squeeze(-1).contiguous())

    def set_reference_noise(self, batch_size, generator=None, keep=False):
        """Draws the noise of the next forward exactly as the reference would consume the CPU RNG stream, so that
        ``torch.manual_seed(s); net.set_reference_noise(B); net(x, step)`` reproduces
        ``torch.manual_seed(s); reference_net(x, step)`` on the CPU ("same inputs and seeds").  Order of the reference's
        draws (models.py:333-336,83-85,92-97,402-403 inside the row-major cell loop models.py:68-117): per cell
        ``normal[B,1]`` x4 (cy, cx, height, width), ``normal[B,A]``, ``normal[B,1]``, ``uniform[B,1]``.
        ``generator``: a CPU ``torch.Generator`` (default: the global CPU generator, like the reference)."""
        _, Hc, Wc = (int(v) for v in self.feature_space_dim)
        A = self._cfg.n_attr
        eps_where, eps_attr = torch.empty(batch_size, 4, Hc, Wc), torch.empty(batch_size, A, Hc, Wc)
        eps_depth, u_pres = torch.empty(batch_size, 1, Hc, Wc), torch.empty(batch_size, 1, Hc, Wc)
        for h in range(Hc):
            for w in range(Wc):
                for k in range(4):
                    eps_where[:, k, h, w] = torch.empty(batch_size, 1).normal_(generator=generator)[:, 0]
                eps_attr[:, :, h, w] = torch.empty(batch_size, A).normal_(generator=generator)
                eps_depth[:, :, h, w] = torch.empty(batch_size, 1).normal_(generator=generator)
                u_pres[:, :, h, w] = torch.rand(batch_size, 1, generator=generator)
        self.set_noise(eps_where, eps_attr, eps_depth, u_pres, keep=keep)

    def _draw_noise(self, B, HW, device):
        A = self._cfg.n_attr
        if self._noise is not None:
            noise = tuple(t.to(device) for t in self._noise)
            self._noise = noise if getattr(self, "_keep_noise", False) else None
            return noise
        return (torch.randn(B, HW, 4, device=device), torch.randn(B, HW, A, device=device),
                torch.randn(B, HW, device=device), torch.rand(B, HW, device=device))

    # ------------------------------------------------------------------------------------
    def _get_plan(self, device):
        if self._plan is not None and self._plan.order_dev.device == device:
            return self._plan
        c = self._cfg
        _, Hc, Wc = (int(v) for v in self.feature_space_dim)
        C, Ih, Iw = (int(v) for v in self.image_shape)
        geom = K.BoxGeom(yx_scale=float(c.max_yx - c.min_yx), yx_min=c.min_yx, hw_scale=float(c.max_hw - c.min_hw),
                         hw_min=c.min_hw, anchor=float(c.anchor[0]), img_h=float(Ih), img_w=float(Iw),
                         cell_ratio_y=self.pixels_per_cell[0] / Ih, cell_ratio_x=self.pixels_per_cell[1] / Iw)
        plan = ops.SweepPlan(schedule=build_schedule(Hc, Wc, c.n_lookback), geom=geom, B_hint=0,
                             F=int(self.feature_space_dim[0]), A=c.n_attr, P=c.n_pass, C=C, Ih=Ih, Iw=Iw,
                             G=int(c.object_shape[0]))
        plan.fused_forward = plan.fused_backward = "SPAIR_UNFUSED_SWEEP" not in os.environ
        if plan.G * plan.G > K.RENDER_MAX_TEXELS or C > K.RENDER_MAX_CHANNELS:
            raise K.SpairKernelError("the fused renderer needs OBJECT_SHAPE[0]^2 <= %d texels and <= %d image channels, got "
                                     "G=%d, C=%d" % (K.RENDER_MAX_TEXELS, K.RENDER_MAX_CHANNELS, plan.G, C))
        plan.n_hidden = dict(box=len(self.box_network.body) // 2, enc=(len(self.object_encoder) - 1) // 2,
                             z=len(self.z_network.body) // 2, obj=(len(self.obj_network) - 1) // 2)
        plan.to(device)
        D = 4 + c.n_attr + 1
        pm, ps = torch.empty(D), torch.empty(D)
        for col, name in ((0, "cy_logit"), (1, "cx_logit"), (2, "height_logit"), (3, "width_logit"), (D - 1, "depth_logit")):
            pm[col], ps[col] = c.priors[name]
        pm[4:4 + c.n_attr], ps[4:4 + c.n_attr] = c.priors["attr"]
        plan.prior_mean, plan.prior_std = pm.to(device), ps.to(device)
        self._plan = plan
        return plan

    def _sweep_params(self):
        def seq(m):
            return [p for layer in m if isinstance(layer, nn.Linear) for p in (layer.weight, layer.bias)]

        def multi(m):
            return seq(m.body) + [p for head in m.output_layers for p in (head.weight, head.bias)]

        return multi(self.box_network) + seq(self.object_encoder) + multi(self.z_network) + seq(self.obj_network)

    def prepare_step(self, global_step, device):
        """Host side of one step: evaluates the two schedules that depend on ``global_step`` — the
        training wheel (models.py:59) and the count prior (models.py:184-193) — in fp32 exactly as the
        reference does, and ships them to two small persistent device buffers (1 and HW+1 floats) through
        a ring of pinned staging buffers, so no step ever synchronises with the host.  Called by
        ``forward`` itself, or by ``GraphedTrainStep`` before replaying a captured step."""
        _, Hc, Wc = (int(v) for v in self.feature_space_dim)
        HW = Hc * Wc
        st = self._step_state
        if st is None or st.wheel.device != device:
            ring = 64
            st = SimpleNamespace(wheel=torch.zeros(1, device=device), count=torch.zeros(HW + 1, device=device),
                                 pin=torch.zeros(ring, HW + 2).pin_memory() if device.type == "cuda" else torch.zeros(ring, HW + 2),
                                 slot=0, ring=ring, copied=[None] * ring)
            self._step_state = st
        self.global_step = global_step
        wheel_host = exponential_decay(global_step, 'cpu', **self._cfg.wheel)
        self.training_wheel = wheel_host
        slot = st.slot
        st.slot = (slot + 1) % st.ring
        if st.copied[slot] is not None:
            st.copied[slot].synchronize()     # the async H2D copy that last read this pinned row has run (host > ring steps ahead)
        row = st.pin[slot]
        row[0] = wheel_host
        row[1:] = self._count_prior(HW)
        st.wheel.copy_(row[:1], non_blocking=True)
        st.count.copy_(row[1:], non_blocking=True)
        if device.type == "cuda":
            if st.copied[slot] is None:
                st.copied[slot] = torch.cuda.Event()
            st.copied[slot].record(torch.cuda.current_stream(device))
        return st

    def _count_prior(self, HW):
        """Truncated geometric prior over the object count, evaluated in fp32 on the host exactly as
        the reference does (models.py:184-193), then shipped to the device ([HW+1] floats)."""
        support = torch.arange(HW + 1, dtype=torch.float32)
        log_odds = exponential_decay(self.global_step, 'cpu', **self._cfg.count_prior)
        prob = 1 / ((-log_odds).exp() + 1)
        dist = (1 - prob) * (prob ** support)
        return dist / dist.sum()

    # ------------------------------------------------------------------------------------
    def forward(self, x, global_step=0):
        debug_tools.benchmark_init()
        K.require_cuda(x, "SPAIR.forward input")
        c = self._cfg
        dev = x.device
        plan = self._get_plan(dev)
        _, Hc, Wc = (int(v) for v in self.feature_space_dim)
        HW = Hc * Wc
        B = x.shape[0]
        C, Ih, Iw = plan.C, plan.Ih, plan.Iw
        self.batch_size = B
        if self._static_step:
            st = self._step_state            # filled by prepare_step() before the graph is replayed
        else:
            st = self.prepare_step(global_step, dev)
            self.writer.add_scalar('training_wheel', self.training_wheel, global_step)
        wheel, count_dist0 = st.wheel, st.count

        x = x.float().contiguous()
        feat = self.backbone(x)
        noise = self._draw_noise(B, HW, dev)

        z_where, attr, depth, pres, dmean, dstd, box = ops.CellSweepFunction.apply(
            plan, x, feat, self.virtual_edge_element, *noise, wheel, *self._sweep_params())

        # KL terms + count-prior scan (models.py:169-262): launched on a side stream, joined after the renderer
        kl_sync = {}
        kl_side = plan.side_streams(dev, 1)[0] if dev.type == "cuda" and "SPAIR_SERIAL_KL" not in os.environ else None
        kl_sums, kl_map = ops.KLFunction.apply(dmean, dstd, pres, plan.prior_mean, plan.prior_std, count_dist0, c.n_attr,
                                               kl_side, kl_sync)

        # decoder MLP over all N = B*HW objects at once (reference models.py:474-481)
        dec = self.object_decoder
        linears = [m for m in dec if isinstance(m, nn.Linear)]
        decoded = ops.USE_TENSOR_CORE_GEMM and len(linears) == 3 and linears[1].in_features % 4 == 0 \
            and linears[2].in_features % 4 == 0
        if decoded:
            # tcgen05 GEMMs; the last one writes sigmoid-decoded texel records instead of logits (ops.DecoderFunction)
            logits = ops.DecoderFunction.apply(attr.reshape(B * HW, c.n_attr), *(p for m in linears for p in (m.weight, m.bias)),
                                               C + 1, c.scales)
        else:
            hidden = dec[:-1](attr.reshape(B * HW, c.n_attr))
            logits = ops.WideLinearFunction.apply(hidden, dec[-1].weight, dec[-1].bias)
        recon_x, recon_loss, _ = ops.RenderFunction.apply(logits, z_where.reshape(B * HW, 4), depth.reshape(-1),
                                                          pres.reshape(-1), x, B, HW, C, plan.G, Ih, Iw, c.scales, decoded)
        if "done" in kl_sync:
            torch.cuda.current_stream(dev).wait_event(kl_sync["done"])
        kl_means = kl_sums.mean(dim=0)                       # batch mean of per-image sums (models.py:553)
        loss = recon_loss + (c.beta * self.kl_scale) * kl_means.sum()   # models.py:558

        # side attributes the reference keeps (models.py:44-48,122-125) — views, no extra compute
        self._latents = SimpleNamespace(z_where=z_where, box=box, attr=attr, depth=depth, pres=pres, dmean=dmean,
                                        dstd=dstd, kl_map=kl_map, kl_sums=kl_sums, recon_loss=recon_loss, Hc=Hc, Wc=Wc)
        self._publish_dist_params(B, Hc, Wc)
        if not self._static_step:
            self._log_losses(loss, recon_loss, kl_means)

        z_where_out = z_where.permute(0, 2, 1).reshape(B, 4, Hc, Wc)
        z_pres_out = pres.reshape(B, 1, Hc, Wc)
        return loss, recon_x, z_where_out, z_pres_out

    # ------------------------------------------------------------------------------------
    def _publish_dist_params(self, B, Hc, Wc):
        L = self._latents
        A = self._cfg.n_attr

        def maps(t, lo, hi):
            return t[..., lo:hi].permute(0, 2, 1).reshape(B, hi - lo, Hc, Wc)

        cols = dict(cy_logit=(0, 1), cx_logit=(1, 2), height_logit=(2, 3), width_logit=(3, 4), attr=(4, 4 + A),
                    depth_logit=(4 + A, 5 + A))
        self.dist_param = {n: dict(mean=maps(L.dmean, *r), sigma=maps(L.dstd, *r)) for n, r in cols.items()}
        self.dist = {n: Normal(loc=p['mean'], scale=p['sigma'], validate_args=False) for n, p in self.dist_param.items()}

    def kl_maps(self):
        """The seven KL maps of the last forward in the reference's layout (models.py:169-262):
        name -> [B, c, Hc, Wc]."""
        L = self._latents
        A = self._cfg.n_attr
        B = L.kl_map.shape[0]
        cols = dict(cy_logit=(0, 1), cx_logit=(1, 2), height_logit=(2, 3), width_logit=(3, 4), attr=(4, 4 + A),
                    depth_logit=(4 + A, 5 + A), pres_dist=(5 + A, 6 + A))
        return {n: L.kl_map[..., lo:hi].permute(0, 2, 1).reshape(B, hi - lo, L.Hc, L.Wc) for n, (lo, hi) in cols.items()}

    def latent_maps(self):
        """z_attr [B,A,Hc,Wc] and z_depth [B,1,Hc,Wc] of the last forward (reference locals, models.py:53-54)."""
        L = self._latents
        B = L.attr.shape[0]
        return dict(z_attr=L.attr.permute(0, 2, 1).reshape(B, -1, L.Hc, L.Wc), z_depth=L.depth.reshape(B, 1, L.Hc, L.Wc))

    def _log_losses(self, loss, recon_loss, kl_means):
        """reference models.py:544-563: scalars go to the writer; the per-step prints (one host sync
        each) only when cfg.VERBOSE."""
        step = self.global_step
        w = self.writer
        w.add_scalar('losses/reconst', recon_loss, step)
        for i, name in enumerate(KL_NAMES):
            w.add_scalar('losses/KL{}'.format(name), kl_means[i], step)
        w.add_scalar('losses/total', loss, step)
        if cfg.VERBOSE:
            print('============ Losses =============')
            print('Reconstruction loss:', '{:.4f}'.format(recon_loss.item()))
            for i, name in enumerate(KL_NAMES):
                print('KL_%s_loss:' % name, '{:.4f}'.format(kl_means[i].item()))
            print('\n ===> total loss:', '{:.4f}'.format(loss.item()))


class Self_Attn(nn.Module):
    """1x1-conv self attention (reference models.py:667-699).  The reference evaluates it on a context
    map and discards the result (models.py:120), so its parameters never receive a gradient; the module
    is kept so that ``state_dict`` has the same ``attn.*`` entries."""

    def __init__(self, in_dim):
        super().__init__()
        self.chanel_in = in_dim
        self.query_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim // 8, kernel_size=1)
        self.key_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim // 8, kernel_size=1)
        self.value_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, x):
        b, ch, w, h = x.size()
        q = self.query_conv(x).view(b, -1, w * h).permute(0, 2, 1)
        k = self.key_conv(x).view(b, -1, w * h)
        attention = self.softmax(torch.bmm(q, k))
        v = self.value_conv(x).view(b, -1, w * h)
        out = torch.bmm(v, attention.permute(0, 2, 1)).view(b, ch, w, h)
        return out, attention
