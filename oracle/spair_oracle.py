"""CPU oracle for the SPAIR per-cell object pipeline — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (PyTorch CPU, fp32) of the hot path of yonkshi/SPAIR_pytorch:
``SPAIR.forward`` (reference ``spair/models.py:35-131``) and the helpers it calls in
``spair/modules.py``.  Every function cites the reference lines it follows.  The arithmetic
of the reference is executed by PyTorch itself (un-vendored, unpinned by the reference:
``README.md:12-15`` "Pytorch 1.0+"); the oracle is therefore defined as *the reference's op
sequence on this image's torch 2.11 CPU build* and it calls the same torch primitives at the
same places (``F.affine_grid``/``F.grid_sample`` with ``align_corners`` left at its default,
``Tensor.inverse``, ``torch.distributions.kl_divergence``, ``F.binary_cross_entropy``).

Differences from the reference are structural only:
  * parameters come in as a ``state_dict``-keyed dict instead of ``nn.Module`` attributes;
  * the random draws (reference ``models.py:402-403,439,450``) are explicit inputs
    (``Noise``); ``draw_noise`` replays the reference's draw order so that seeding the global
    generator and running the reference gives bit-identical samples;
  * shapes come from an ``OracleConfig`` instead of module-level constants.

PINNING STATUS.  The reference's own tests pin nothing at this boundary
(``spair/test/test_backbone.py`` = ``self.fail()`` x4, ``spair/test/test_renderer.py`` has no
assertions) — "parity unpinned" by the reference.  This restatement is instead pinned against
outputs of the reference itself: ``tests/test_oracle_vs_reference.py`` runs the real reference
(``oracle/ref_harness.py``) next to this file inside the build container, and
``tests/golden/*.npz`` (made by ``tests/golden/make_golden.py`` from the real reference) travel
to the GPU box.  The only known answer shipped by the reference, the backbone geometry printed
at ``test_notebook.ipynb:350``, is checked in ``tests/test_oracle.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this module.
"""
from __future__ import annotations

import copy
import itertools
import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F
from torch.distributions import Normal
from torch.distributions.kl import kl_divergence

DEFAULT_TOPOLOGY = [
    dict(filters=128, kernel_size=4, stride=3),
    dict(filters=128, kernel_size=4, stride=2),
    dict(filters=128, kernel_size=4, stride=2),
    dict(filters=128, kernel_size=1, stride=1),
    dict(filters=128, kernel_size=1, stride=1),
    dict(filters=128, kernel_size=1, stride=1),
]
CELL8_TOPOLOGY = [dict(t, stride=2) if t["kernel_size"] == 4 else dict(t) for t in DEFAULT_TOPOLOGY]


@dataclass
class OracleConfig:
    """The values the reference keeps as module constants in ``spair/config.py``."""

    image_shape: tuple = (1, 128, 128)          # config.py:4
    object_shape: tuple = (28, 28)              # config.py:33
    anchor: tuple = (48, 48)                    # config.py:34
    topology: list = field(default_factory=lambda: copy.deepcopy(DEFAULT_TOPOLOGY))  # config.py:7-14
    n_backbone_features: int = 100              # config.py:22
    n_passthrough: int = 100                    # config.py:24
    n_attr: int = 50                            # config.py:27
    n_lookback: int = 1                         # config.py:31
    max_yx: float = 1.5                         # config.py:38
    min_yx: float = -0.5                        # config.py:39
    max_hw: float = 1.0                         # config.py:40
    min_hw: float = 0.0                         # config.py:41
    priors: dict = field(default_factory=lambda: {                   # config.py:45-52
        "cy_logit": (0.0, 1.0), "cx_logit": (0.0, 1.0),
        "height_logit": (7.0, 0.5), "width_logit": (7.0, 0.5),
        "attr": (0.0, 1.0), "depth_logit": (0.0, 1.0)})
    beta: float = 1.0                           # config.py:55
    wheel: dict = field(default_factory=lambda: dict(start=1.0, end=0.0, decay_rate=0.0,
                                                     decay_step=1000.0, staircase=True))   # config.py:58-62
    count_prior: dict = field(default_factory=lambda: dict(start=1000000.0, end=0.0125, decay_rate=0.1,
                                                           decay_step=1000.0, log_space=True))  # config.py:65-69
    obj_logit_scale: float = 2.0                # config.py:74
    alpha_logit_scale: float = 0.1              # config.py:75
    alpha_logit_bias: float = 5.0               # config.py:76

    @property
    def geometry(self):
        return backbone_geometry(self.topology, self.image_shape[-2:])

    @property
    def grid(self):
        return tuple(int(v) for v in self.geometry["n_grid_cells"])

    @property
    def context_dim(self):                      # models.py:26
        return (self.n_lookback * 2 + 1) ** 2 // 2 * (4 + self.n_attr + 1 + 1)


def config_A():  # BASELINE.json configs[0]/[1]: spair/config.py as shipped
    return OracleConfig()


def config_C():  # 128x128, 16x16 cells, 14x14 glimpses
    return OracleConfig(object_shape=(14, 14), topology=copy.deepcopy(CELL8_TOPOLOGY))


def config_D():  # 256x256 RGB, 32x32 cells, 28x28 glimpses
    return OracleConfig(image_shape=(3, 256, 256), topology=copy.deepcopy(CELL8_TOPOLOGY))


def config_tiny():  # small fixture: 1x40x40 canvas, 5x5 cells of 8 px, 8x8 glimpses, 16 px anchor
    return OracleConfig(image_shape=(1, 40, 40), object_shape=(8, 8), anchor=(16, 16),
                        topology=copy.deepcopy(CELL8_TOPOLOGY))


# --------------------------------------------------------------------------------------
# geometry and scalar schedules
# --------------------------------------------------------------------------------------
def backbone_geometry(topology, image_hw):
    """Receptive-field bookkeeping of ``Backbone._build_receptive_field_padding``
    (modules.py:68-105): returns rf size, cell size (cumulative stride), grid and the
    ``ZeroPad2d`` amounts (left, right, top, bottom)."""
    jump = np.array([1, 1])
    rf = np.array([1, 1])
    for layer in topology:
        rf = rf + (np.array(layer["kernel_size"]) - 1) * jump      # modules.py:83
        jump = jump * np.array(layer["stride"])                    # modules.py:84
    cell = jump
    pre = np.floor(rf / 2 - cell / 2).astype("i")                  # modules.py:92
    image_hw = np.array(image_hw)
    n_cells = np.ceil(image_hw / cell).astype("i")                 # modules.py:95
    required = rf + (n_cells - 1) * cell                           # modules.py:96
    post = required - image_hw - pre                               # modules.py:97
    return dict(rf_size=rf, grid_cell_size=cell, n_grid_cells=n_cells, pre_padding=pre,
                post_padding=post, required_image_size=required,
                pad=(int(pre[1]), int(post[1]), int(pre[0]), int(post[0])))


def exponential_decay(global_step, start, end, decay_rate, decay_step, staircase=False, log_space=False):
    """modules.py:191-213, evaluated in fp32 tensors exactly as there."""
    gs = torch.tensor(global_step, dtype=torch.get_default_dtype())
    t = gs // decay_step if staircase else gs / decay_step
    value = (start - end) * (decay_rate ** t) + end
    if log_space:
        value = (value + 1e-6).log()
    return value


def latent_to_mean_std(latent):
    """modules.py:167-176."""
    mean, log_std = torch.chunk(latent, 2, dim=-1)
    return mean, torch.sigmoid(log_std.clamp(-10, 10)) * 2


def clamped_sigmoid(logit, use_analytical=False):
    """modules.py:178-189."""
    if use_analytical:
        return 1 / ((-logit).exp() + 1)
    return torch.sigmoid(torch.clamp(logit, -10, 10))


def safe_log(t):
    """modules.py:296-297."""
    return torch.log(t + 1e-9)


# --------------------------------------------------------------------------------------
# spatial transformer (both directions)
# --------------------------------------------------------------------------------------
def stn(image, z_where, out_hw, inverse=False):
    """modules.py:216-273.  ``z_where`` rows are (xt, yt, xs, ys) normalised to the image;
    forward = cut a glimpse (border padding), inverse = paste a glimpse on the canvas (zeros
    padding; theta is inverted through ``Tensor.inverse``, modules.py:256-262).
    ``align_corners`` is left at torch's default (False on torch >= 1.3)."""
    n = image.shape[0]
    xt, yt, xs, ys = (z_where[:, i] for i in range(4))
    xt = xt * 2 - 1                                                # modules.py:243-244
    yt = yt * 2 - 1
    theta = torch.zeros(n, 2, 3)
    theta[:, 0, 0] = xs
    theta[:, 1, 1] = ys
    theta[:, 0, 2] = xt
    theta[:, 1, 2] = yt
    if inverse:
        last = torch.tensor([0.0, 0.0, 1.0]).repeat(n, 1, 1)
        theta = torch.cat([theta, last], dim=-2).inverse()[:, :2, :]
    grid = F.affine_grid(theta, [n, image.shape[1], int(out_hw[0]), int(out_hw[1])], align_corners=False)
    return F.grid_sample(image, grid, mode="bilinear", padding_mode="zeros" if inverse else "border",
                         align_corners=False)


# --------------------------------------------------------------------------------------
# networks (dense contractions; same nn.functional calls the reference modules dispatch)
# --------------------------------------------------------------------------------------
def backbone_forward(params, x, cfg: OracleConfig):
    """modules.py:107-111 — ZeroPad2d + conv/ReLU stack + 1x1 output conv."""
    h = F.pad(x, cfg.geometry["pad"])
    for i, layer in enumerate(cfg.topology):
        h = F.relu(F.conv2d(h, params["backbone.net.conv_%d.weight" % i], params["backbone.net.conv_%d.bias" % i],
                            stride=layer["stride"]))
    return F.conv2d(h, params["backbone.net.conv_out.weight"], params["backbone.net.conv_out.bias"])


def mlp_single(params, prefix, x, n_hidden=2):
    """build_MLP(..., output=n) path, modules.py:143-156."""
    for i in range(n_hidden):
        x = F.relu(F.linear(x, params["%s.dense%d.weight" % (prefix, i)], params["%s.dense%d.bias" % (prefix, i)]))
    return F.linear(x, params[prefix + ".out.weight"], params[prefix + ".out.bias"])


def mlp_multi(params, prefix, x, n_hidden=2, n_out=2):
    """build_MLP(..., multiple_output=(..)) path, modules.py:158-163,276-284."""
    for i in range(n_hidden):
        x = F.relu(F.linear(x, params["%s.body.dense%d.weight" % (prefix, i)],
                            params["%s.body.dense%d.bias" % (prefix, i)]))
    return [F.linear(x, params["%s.output_layers.%d.weight" % (prefix, j)],
                     params["%s.output_layers.%d.bias" % (prefix, j)]) for j in range(n_out)]


# --------------------------------------------------------------------------------------
# noise
# --------------------------------------------------------------------------------------
@dataclass
class Noise:
    eps_where: torch.Tensor   # [B,4,Hc,Wc]  order (cy, cx, height, width) — models.py:333-336
    eps_attr: torch.Tensor    # [B,A,Hc,Wc]
    eps_depth: torch.Tensor   # [B,1,Hc,Wc]
    u_pres: torch.Tensor      # [B,1,Hc,Wc]  uniform(0,1) — models.py:402-403


def draw_noise(seed, batch, grid, n_attr=50) -> Noise:
    """Replays the reference's draw order: per cell, row-major, normal [B,1] x4
    (cy, cx, height, width), normal [B,A], normal [B,1], uniform [B,1]."""
    Hc, Wc = grid
    n = Noise(torch.empty(batch, 4, Hc, Wc), torch.empty(batch, n_attr, Hc, Wc),
              torch.empty(batch, 1, Hc, Wc), torch.empty(batch, 1, Hc, Wc))
    torch.manual_seed(seed)
    for h, w in itertools.product(range(Hc), range(Wc)):
        for k in range(4):
            n.eps_where[:, k, h, w] = torch.empty(batch, 1).normal_()[:, 0]
        n.eps_attr[:, :, h, w] = torch.empty(batch, n_attr).normal_()
        n.eps_depth[:, :, h, w] = torch.empty(batch, 1).normal_()
        n.u_pres[:, :, h, w] = torch.rand(batch, 1)
    return n


def random_noise(generator, batch, grid, n_attr=50) -> Noise:
    Hc, Wc = grid
    return Noise(torch.randn(batch, 4, Hc, Wc, generator=generator), torch.randn(batch, n_attr, Hc, Wc, generator=generator),
                 torch.randn(batch, 1, Hc, Wc, generator=generator), torch.rand(batch, 1, Hc, Wc, generator=generator))


# --------------------------------------------------------------------------------------
# per-cell pieces
# --------------------------------------------------------------------------------------
def context_offsets(n_lookback=1):
    """Neighbour offsets (dh, dw) in concat order, models.py:292-307:
    for lookback 1 -> (-1,-1), (-1,0), (-1,+1), (0,-1)."""
    r = n_lookback
    cols = np.arange(-r, r + 1)
    rows = np.arange(-r, 1)
    mesh = np.array(np.meshgrid(rows, cols)).T
    flat = np.reshape(mesh, (-1, 2))
    return [tuple(int(v) for v in c) for c in flat[:-(r + 1), :]]


def freeze(wheel, *ts):
    """models.py:413-429: value unchanged, gradient scaled by (1 - wheel)."""
    out = [wheel * t.detach() + (1 - wheel) * t for t in ts]
    return out[0] if len(out) == 1 else out


def build_box(latent, h, w, eps4, wheel, cfg: OracleConfig, cell_px):
    """models.py:322-381.  ``eps4`` = [B,4] normal draws for (cy, cx, height, width).
    Returns box=[cell_x, cell_y, width, height], normalized_box=[xt, yt, xs, ys] and the
    (frozen) means / stds [B,4] in (cy, cx, height, width) order."""
    mean, std = latent_to_mean_std(latent)
    mean, std = freeze(wheel, mean, std)
    z = mean + eps4 * std                                          # Normal.rsample, models.py:439,450
    cy_l, cx_l, h_l, w_l = torch.chunk(z, 4, dim=-1)
    cell_y = float(cfg.max_yx - cfg.min_yx) * clamped_sigmoid(cy_l) + cfg.min_yx     # models.py:339-347
    cell_x = float(cfg.max_yx - cfg.min_yx) * clamped_sigmoid(cx_l) + cfg.min_yx
    height = float(cfg.max_hw - cfg.min_hw) * clamped_sigmoid(h_l) + cfg.min_hw      # models.py:351-359
    width = float(cfg.max_hw - cfg.min_hw) * clamped_sigmoid(w_l) + cfg.min_hw
    box = torch.cat([cell_x, cell_y, width, height], dim=-1)
    _, img_h, img_w = cfg.image_shape
    anchor = cfg.anchor[0]                                         # models.py:366 (uses [0] for both axes)
    ys = height * anchor / img_h                                   # models.py:369-370
    xs = width * anchor / img_w
    yt = (cell_px[0] / img_h) * (cell_y + h)                       # models.py:373-374
    xt = (cell_px[1] / img_w) * (cell_x + w)
    return box, torch.cat([xt, yt, xs, ys], dim=-1), mean, std


def build_obj_pres(logit, u, wheel):
    """models.py:393-411: relaxed-Bernoulli sample with logistic noise, temperature 1."""
    logit = freeze(wheel, logit)
    log_odds = torch.clamp(logit, -10.0, 10.0)
    eps = 10e-10
    noise = torch.log(u + eps) - torch.log(1.0 - u + eps)
    return torch.sigmoid((log_odds + noise) / 1.0)


def count_prior_distribution(step, n_cells, cfg: OracleConfig):
    """models.py:184-193: truncated geometric prior over the object count."""
    support = torch.arange(n_cells + 1, dtype=torch.get_default_dtype())
    log_odds = exponential_decay(step, **cfg.count_prior)
    prob = 1 / ((-log_odds).exp() + 1)
    dist = (1 - prob) * (prob ** support)
    return support, dist / dist.sum(), prob


def compute_kl(dist_mean, dist_std, z_pres, z_pres_prob, step, cfg: OracleConfig, check_finite=True):
    """models.py:169-262.  ``dist_mean/std`` are dicts name -> [B,c,Hc,Wc]."""
    kl = {}
    for name in dist_mean:
        prior = Normal(*[float(v) for v in cfg.priors[name]])
        kl[name] = z_pres * kl_divergence(Normal(dist_mean[name], dist_std[name]), prior)   # models.py:175-177
    B, _, Hc, Wc = z_pres.shape
    HW = Hc * Wc
    support, dist, _ = count_prior_distribution(step, HW, cfg)
    dist = dist.repeat(B, 1)
    count_so_far = torch.zeros(B, 1)
    obj_kl = torch.ones(B, 1, Hc, Wc)
    i = 0
    for h, w in itertools.product(range(Hc), range(Wc)):
        p_z_given_c = torch.clamp(support - count_so_far, min=0.0, max=(HW - i)) / (HW - i)   # models.py:206
        p_z = torch.bmm(dist[:, None, :], p_z_given_c[:, :, None]).squeeze(-1)                # models.py:214
        prob = z_pres_prob[:, :, h, w]
        obj_kl[:, :, h, w] = (prob * (safe_log(prob) - safe_log(p_z))
                              + (1 - prob) * (safe_log(1 - prob) - safe_log(1 - p_z)))        # models.py:223-228
        sample = torch.round(z_pres[:, :, h, w])                                             # models.py:232
        mult = sample * p_z_given_c + (1 - sample) * (1 - p_z_given_c)
        dist1 = mult * dist
        dist = dist1 / dist1.sum(dim=1, keepdim=True).clamp(min=1e-6)                         # models.py:237-241
        if check_finite:
            _nan_check(obj_kl[:, :, h, w], p_z, prob, dist, p_z_given_c)                     # models.py:245-253
        count_so_far = count_so_far + sample
        i += 1
    kl["pres_dist"] = obj_kl
    return kl


def render(params, z_attr, z_where, z_depth, z_pres, cfg: OracleConfig):
    """models.py:452-542: decode every object, inverse-warp it onto the canvas (materialising
    [N, C+2, I, I]), alpha * colour, importance-normalised sum over objects, clamp to [0,1]."""
    B, _, Hc, Wc = z_where.shape
    C, img_h, img_w = cfg.image_shape
    G = cfg.object_shape[0]
    where = z_where.permute(0, 2, 3, 1).contiguous().view(-1, 4)
    depth = z_depth.view(-1, 1, 1)
    pres = z_pres.view(-1, 1, 1)
    logits = mlp_single(params, "object_decoder", z_attr.permute(0, 2, 3, 1).contiguous().view(-1, cfg.n_attr))
    logits = logits.view(-1, G, G, C + 1)
    scale = torch.tensor([cfg.obj_logit_scale] * C + [cfg.alpha_logit_scale])
    bias = torch.tensor([0.0] * C + [cfg.alpha_logit_bias])
    objects = clamped_sigmoid(logits * scale + bias, use_analytical=True)         # models.py:485-492
    colour = objects[..., :C]
    alpha = objects[..., C] * pres                                                # models.py:496
    importance = torch.clamp(alpha * depth, min=0.01)                             # models.py:499-500
    stacked = torch.cat([colour, alpha[..., None], importance[..., None]], dim=-1).permute(0, 3, 1, 2)
    warped = stn(stacked, where, [img_h, img_w], inverse=True).contiguous().view(B, Hc * Wc, C + 2, img_h, img_w)
    colour_w = warped[:, :, :C]
    alpha_w = warped[:, :, C:C + 1]
    imp_w = warped[:, :, C + 1:C + 2] + 1e-9                                      # models.py:527
    img = alpha_w * colour_w
    imp_w = imp_w / imp_w.sum(dim=1, keepdim=True)                                # models.py:532
    return torch.clamp((img * imp_w).sum(dim=1), min=0, max=1)                    # models.py:535-540


def build_loss(x, recon, kl, cfg: OracleConfig):
    """models.py:544-563: BCE(sum) + beta * sum_names mean_b sum_{c,h,w} KL."""
    recon_loss = F.binary_cross_entropy(recon, x, reduction="sum")
    kl_loss = 0
    kl_means = {}
    for name, m in kl.items():
        kl_means[name] = torch.mean(torch.sum(m, dim=[1, 2, 3]))
        kl_loss = kl_loss + kl_means[name]
    return recon_loss + cfg.beta * kl_loss, recon_loss, kl_means


def _nan_check(*tensors):
    """debug_tools.py:245-271 (``nan_hunter``): one isnan-reduction per tensor."""
    for t in tensors:
        if torch.isnan(t).sum() > 0:
            raise AssertionError("NAN Detected by Nan detector")


# --------------------------------------------------------------------------------------
# the whole path
# --------------------------------------------------------------------------------------
def forward(params, x, global_step, noise: Noise, cfg: OracleConfig, check_finite=True):
    """models.py:35-131.  Returns a dict with loss, recon_x, z_where, z_attr, z_depth, z_pres,
    the distribution parameter maps and the seven KL maps."""
    B = x.shape[0]
    Hc, Wc = cfg.grid
    geom = cfg.geometry
    cell_px = tuple(int(v) for v in geom["grid_cell_size"])                       # models.py:32
    feat = backbone_forward(params, x, cfg)
    assert tuple(feat.shape[-2:]) == (Hc, Wc), (feat.shape, Hc, Wc)
    A = cfg.n_attr
    z_where = torch.empty(B, 4, Hc, Wc)
    z_attr = torch.empty(B, A, Hc, Wc)
    z_depth = torch.empty(B, 1, Hc, Wc)
    z_pres = torch.empty(B, 1, Hc, Wc)
    names = ["cy_logit", "cx_logit", "height_logit", "width_logit", "attr", "depth_logit"]
    widths = dict(cy_logit=1, cx_logit=1, height_logit=1, width_logit=1, attr=A, depth_logit=1)
    dist_mean = {n: torch.empty(B, widths[n], Hc, Wc) for n in names}
    dist_std = {n: torch.empty(B, widths[n], Hc, Wc) for n in names}
    edge = params["virtual_edge_element"][None, :].repeat(B, 1)                   # models.py:58
    wheel = exponential_decay(global_step, **cfg.wheel)                           # models.py:59
    if check_finite:
        _nan_check(edge, feat)                                                    # models.py:65
    offsets = context_offsets(cfg.n_lookback)
    cell_out = {}
    for h, w in itertools.product(range(Hc), range(Wc)):                          # models.py:68
        cell_feat = feat[:, :, h, w]
        context = torch.cat([cell_out.get((h + dh, w + dw), edge) for dh, dw in offsets], dim=-1)  # models.py:292-320
        # z_where  (models.py:76-79)
        rep, passthru = mlp_multi(params, "box_network", torch.cat((cell_feat, context), dim=-1))
        box, nbox, mean4, std4 = build_box(rep, h, w, noise.eps_where[:, :, h, w], wheel, cfg, cell_px)
        z_where[:, :, h, w] = nbox
        for k, n in enumerate(names[:4]):
            dist_mean[n][:, :, h, w] = mean4[:, k:k + 1]
            dist_std[n][:, :, h, w] = std4[:, k:k + 1]
        # z_what   (models.py:82-85, 383-391)
        glimpse = stn(x, nbox, cfg.object_shape)
        a_mean, a_std = latent_to_mean_std(mlp_single(params, "object_encoder", glimpse.flatten(start_dim=1)))
        attr = a_mean + noise.eps_attr[:, :, h, w] * a_std
        dist_mean["attr"][:, :, h, w] = a_mean
        dist_std["attr"][:, :, h, w] = a_std
        z_attr[:, :, h, w] = attr
        # z_depth  (models.py:88-97)
        d_lat, passthru = mlp_multi(params, "z_network", torch.cat([cell_feat, context, passthru, box, attr], dim=1))
        d_mean, d_std = freeze(wheel, *latent_to_mean_std(d_lat))
        d_logit = d_mean + noise.eps_depth[:, :, h, w] * d_std
        dist_mean["depth_logit"][:, :, h, w] = d_mean
        dist_std["depth_logit"][:, :, h, w] = d_std
        depth = 4 * clamped_sigmoid(d_logit)
        z_depth[:, :, h, w] = depth
        # z_pres   (models.py:100-105)
        p_logit = mlp_single(params, "obj_network", torch.cat([cell_feat, context, passthru, box, attr, depth], dim=1))
        pres = build_obj_pres(p_logit, noise.u_pres[:, :, h, w], wheel)
        z_pres[:, :, h, w] = pres
        cell_out[(h, w)] = torch.cat((box, attr, depth, pres), dim=-1)            # models.py:106
        if check_finite:
            _nan_check(context, nbox, attr, depth, pres)                          # models.py:108-117
    kl = compute_kl(dist_mean, dist_std, z_pres, z_pres, global_step, cfg, check_finite)   # models.py:127 (prob == sample)
    recon = render(params, z_attr, z_where, z_depth, z_pres, cfg)                 # models.py:128
    loss, recon_loss, kl_means = build_loss(x, recon, kl, cfg)                    # models.py:129
    return dict(loss=loss, recon_loss=recon_loss, kl_means=kl_means, recon_x=recon, z_where=z_where, z_attr=z_attr,
                z_depth=z_depth, z_pres=z_pres, kl=kl, dist_mean=dist_mean, dist_std=dist_std, feat=feat)


def params_from_state_dict(state_dict, requires_grad=True, dtype=torch.float32):
    """Detached CPU copies (fp32 by default) of a SPAIR ``state_dict`` (reference key names,
    SURVEY.md §8(b)), as leaves that can receive gradients.  ``attn.*`` entries are carried but never
    used (models.py:120 discards the attention output)."""
    out = {}
    for k, v in state_dict.items():
        t = v.detach().to("cpu", dtype).clone()
        t.requires_grad_(requires_grad and not k.startswith("attn."))
        out[k] = t
    return out


def forward_backward(params, x, global_step, noise, cfg, check_finite=True):
    """train.py:65-66 — forward then ``loss.backward``.  Gradients land in ``params[k].grad``."""
    for p in params.values():
        p.grad = None
    out = forward(params, x, global_step, noise, cfg, check_finite)
    out["loss"].backward()
    return out


def forward_backward_fp64(state_dict, x, global_step, noise, cfg):
    """The same op sequence evaluated in float64 — the "true" value of the reference's formulas.

    Needed because the reference's own fp32 backward is ill-conditioned on some inputs: the
    importance normalisation (models.py:527-535) makes d(loss)/d(importance) a difference of nearly
    equal numbers at singly-covered pixels, and BCE gradients reach 1e12 on uncovered pixels
    (SURVEY.md §7).  On such inputs the reference's fp32 gradients differ from this fp64 evaluation by
    up to tens of percent, so no independent implementation can match them to 1e-4; the parity tests then
    require the CUDA result to be at least as close to fp64 as the fp32 reference path is.
    Returns (out dict, params dict with .grad in float64)."""
    saved = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        params = params_from_state_dict(state_dict, dtype=torch.float64)
        n64 = Noise(*(t.double() for t in (noise.eps_where, noise.eps_attr, noise.eps_depth, noise.u_pres)))
        out = forward_backward(params, x.double(), global_step, n64, cfg, check_finite=False)
    finally:
        torch.set_default_dtype(saved)
    return out, params


# --------------------------------------------------------------------------------------
# synthetic scenes — shared by tests and bench so both arms see one workload
# --------------------------------------------------------------------------------------
def scattered_sprites(batch, image_shape, seed=1234, max_sprites=9, sprite_px=(10, 20)):
    """Procedural stand-in for the un-shipped ``scattered_mnist_128x128_obj14x14.hdf5``
    (train.py:38; schema dataloader.py:23-33): ``k ~ U{1..max}`` soft blobs per image, each
    a random glyph-like union of strokes, max-composited on black, values in [0,1]."""
    C, H, W = image_shape
    rng = np.random.RandomState(seed)
    imgs = np.zeros((batch, C, H, W), np.float32)
    for b in range(batch):
        for _ in range(rng.randint(1, max_sprites + 1)):
            s = rng.randint(sprite_px[0], min(sprite_px[1], H, W) + 1)
            yy, xx = np.mgrid[0:s, 0:s].astype(np.float32) / max(s - 1, 1)
            sprite = np.zeros((s, s), np.float32)
            for _ in range(rng.randint(2, 5)):
                cy, cx, r = rng.uniform(0.2, 0.8), rng.uniform(0.2, 0.8), rng.uniform(0.08, 0.3)
                sprite = np.maximum(sprite, np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r)))
            sprite = np.clip(sprite * 1.2, 0, 1)
            y0, x0 = rng.randint(0, H - s + 1), rng.randint(0, W - s + 1)
            colour = np.ones(C, np.float32) if C == 1 else rng.uniform(0.3, 1.0, C).astype(np.float32)
            for c in range(C):
                imgs[b, c, y0:y0 + s, x0:x0 + s] = np.maximum(imgs[b, c, y0:y0 + s, x0:x0 + s], sprite * colour[c])
    return torch.from_numpy(imgs)
