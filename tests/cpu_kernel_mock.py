"""Test double of ``spair_pytorch_b200.kernels`` for HOST-LOGIC tests on machines without a GPU.

TEST INFRASTRUCTURE ONLY.  The product has no CPU path; these functions exist so that the
host orchestration in ``ops.py`` / ``models.py`` (wavefront schedule, buffer/column bookkeeping,
the hand-written backward sweep, gradient routing) can be exercised on the CPU-only build
container by monkeypatching the kernel binding — exactly like mocking a native driver.  Each
mock re-states one C-ABI entry point with the oracle's torch-CPU arithmetic
(``oracle/spair_oracle.py``); backward mocks differentiate the forward mock with autograd.
They are never imported by ``spair_pytorch_b200``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle import spair_oracle as so


def install(monkeypatch):
    """Replace every kernel entry of ``spair_pytorch_b200.kernels`` by its mock."""
    from spair_pytorch_b200 import kernels as K
    for name, fn in list(globals().items()):
        if callable(fn) and hasattr(K, name) and not name.startswith("_") and name not in ("install",):
            monkeypatch.setattr(K, name, fn)
    monkeypatch.setattr(K, "require_cuda", lambda *a, **k: None)
    # the tcgen05 GEMM path (decoder texel records, gemm3x weight gradients) has no test double: host-logic tests run
    # the cuBLAS-shaped branch of ops.py / models.py; the GPU tests cover both branches
    from spair_pytorch_b200 import ops
    monkeypatch.setattr(ops, "USE_TENSOR_CORE_GEMM", False)


def parallel_branches(device, stream_pool, thunks):
    for fn in thunks:       # no streams on the CPU: the branches run one after the other
        fn()


def sweep_max_rows():
    return 0        # the fused forward sweep has no test double: host-logic tests exercise the per-wavefront path


def _rows(cells, B):
    return [int(c) for c in cells.tolist()]


# ------------------------------------------------------------------------------------------
def context_gather_fwd(feat, box, attr, depth, pres, edge, cells, offsets, dsts):
    B, Fdim, Hc, Wc = feat.shape
    for k, cell in enumerate(_rows(cells, B)):
        h, w = divmod(cell, Wc)
        parts = [feat[:, :, h, w]]
        for dh, dw in offsets:
            nh, nw = h + dh, w + dw
            if 0 <= nh < Hc and 0 <= nw < Wc:
                o = nh * Wc + nw
                parts.append(torch.cat([box[:, o], attr[:, o], depth[:, o, None], pres[:, o, None]], -1))
            else:
                parts.append(edge[None].expand(B, -1))
        row = torch.cat(parts, -1)
        for d in dsts:
            if d is not None:
                d[k * B:(k + 1) * B, :row.shape[1]] = row


def context_grad_gather(dxs, col0, cells, wf_pos, offsets, B, Hc, Wc, A, out):
    E = A + 6
    for k, cell in enumerate(_rows(cells, B)):
        h, w = divmod(cell, Wc)
        acc = torch.zeros(B, E)
        for s, (dh, dw) in enumerate(offsets):
            ch, cw = h - dh, w - dw
            if 0 <= ch < Hc and 0 <= cw < Wc:
                pos = int(wf_pos[ch * Wc + cw])
                for dx in dxs:
                    if dx is not None:
                        acc = acc + dx[pos * B:(pos + 1) * B, col0 + s * E: col0 + (s + 1) * E]
        out[k * B:(k + 1) * B, :E] = acc


# ------------------------------------------------------------------------------------------
def _box_fwd(y8, eps4, h, w, g):
    mean, std = so.latent_to_mean_std(y8)
    z = mean + eps4 * std
    cy, cx, hh, ww = torch.chunk(z, 4, -1)
    cell_y = g.yx_scale * so.clamped_sigmoid(cy) + g.yx_min
    cell_x = g.yx_scale * so.clamped_sigmoid(cx) + g.yx_min
    height = g.hw_scale * so.clamped_sigmoid(hh) + g.hw_min
    width = g.hw_scale * so.clamped_sigmoid(ww) + g.hw_min
    box = torch.cat([cell_x, cell_y, width, height], -1)
    ys = height * g.anchor / g.img_h
    xs = width * g.anchor / g.img_w
    yt = g.cell_ratio_y * (cell_y + h)
    xt = g.cell_ratio_x * (cell_x + w)
    return box, torch.cat([xt, yt, xs, ys], -1), mean, std


def box_head_fwd(y, eps, cells, B, HW, Wc, geom, box, z_where, dmean, dstd, xdsts, n_pt, pt_dst):
    for k, cell in enumerate(_rows(cells, B)):
        h, w = divmod(cell, Wc)
        sl = slice(k * B, (k + 1) * B)
        b, zw, mean, std = _box_fwd(y[sl, :8], eps[:, cell], h, w, geom)
        box[:, cell], z_where[:, cell] = b, zw
        dmean[:, cell, :4], dstd[:, cell, :4] = mean, std
        for d in xdsts:
            if d is not None:
                d[sl, :4] = b
        if n_pt:
            pt_dst[sl, :n_pt] = y[sl, 8:8 + n_pt]


@torch.enable_grad()
def box_head_bwd(y, eps, cells, B, HW, Wc, geom, wheel, d_boxes, d_zw_local, d_zw_img, d_dmean, d_dstd, ld_dist,
                 n_pt, d_pt_src, d_y):
    keep = 1.0 - float(wheel[0])
    for k, cell in enumerate(_rows(cells, B)):
        h, w = divmod(cell, Wc)
        sl = slice(k * B, (k + 1) * B)
        y8 = y[sl, :8].detach().clone().requires_grad_(True)
        b, zw, mean, std = _box_fwd(y8, eps[:, cell], h, w, geom)
        g_b = sum(d[sl, :4] for d in d_boxes if d is not None)
        g_zw = torch.zeros(B, 4)
        if d_zw_local is not None:
            g_zw = g_zw + d_zw_local[sl, :4]
        if d_zw_img is not None:
            g_zw = g_zw + d_zw_img[:, cell]
        terms = (b * g_b).sum() + (zw * g_zw).sum()
        if d_dmean is not None:
            terms = terms + (mean * d_dmean[:, cell, :4]).sum() + (std * d_dstd[:, cell, :4]).sum()
        (g,) = torch.autograd.grad(terms, y8)
        d_y[sl, :8] = keep * g
        if n_pt:
            d_y[sl, 8:8 + n_pt] = d_pt_src[sl, :n_pt] if d_pt_src is not None else 0.0


def _normal_fwd(y, W, eps, squash, scale):
    mean, std = so.latent_to_mean_std(y[:, :2 * W])
    z = mean + eps * std
    if squash:
        z = scale * so.clamped_sigmoid(z)
    return z, mean, std


def normal_head_fwd(y, W, eps, cells, B, HW, squash, scale, out, dmean_view, dstd_view, ld_dist, xdsts, n_pt, pt_dst):
    out3 = out.view(B, HW, W)
    eps3 = eps.view(B, HW, W)
    for k, cell in enumerate(_rows(cells, B)):
        sl = slice(k * B, (k + 1) * B)
        z, mean, std = _normal_fwd(y[sl], W, eps3[:, cell], squash, scale)
        out3[:, cell] = z
        dmean_view[:, cell, :W], dstd_view[:, cell, :W] = mean, std
        for d in xdsts:
            if d is not None:
                d[sl, :W] = z
        if n_pt:
            pt_dst[sl, :n_pt] = y[sl, 2 * W:2 * W + n_pt]


@torch.enable_grad()
def normal_head_bwd(y, W, eps, cells, B, HW, squash, scale, wheel, d_outs, d_out_img, d_dmean_view, d_dstd_view,
                    ld_dist, n_pt, d_pt_src, d_y):
    keep = 1.0 if wheel is None else 1.0 - float(wheel[0])
    eps3 = eps.view(B, HW, W)
    for k, cell in enumerate(_rows(cells, B)):
        sl = slice(k * B, (k + 1) * B)
        yy = y[sl, :2 * W].detach().clone().requires_grad_(True)
        z, mean, std = _normal_fwd(yy, W, eps3[:, cell], squash, scale)
        g_z = sum(d[sl, :W] for d in d_outs if d is not None)
        if d_out_img is not None:
            g_z = g_z + d_out_img.view(B, HW, W)[:, cell]
        terms = (z * g_z).sum()
        if d_dmean_view is not None:
            terms = terms + (mean * d_dmean_view[:, cell, :W]).sum() + (std * d_dstd_view[:, cell, :W]).sum()
        (g,) = torch.autograd.grad(terms, yy)
        d_y[sl, :2 * W] = keep * g
        if n_pt:
            d_y[sl, 2 * W:2 * W + n_pt] = d_pt_src[sl, :n_pt] if d_pt_src is not None else 0.0


def _pres_fwd(logit, u):
    lo = torch.clamp(logit, -10.0, 10.0)
    noise = torch.log(u + 10e-10) - torch.log(1.0 - u + 10e-10)
    return torch.sigmoid((lo + noise) / 1.0)


def pres_head_fwd(y, u, cells, B, HW, pres):
    for k, cell in enumerate(_rows(cells, B)):
        pres[:, cell] = _pres_fwd(y[k * B:(k + 1) * B, 0], u[:, cell])


@torch.enable_grad()
def pres_head_bwd(y, u, cells, B, HW, wheel, d_local, d_img, d_y):
    keep = 1.0 - float(wheel[0])
    for k, cell in enumerate(_rows(cells, B)):
        sl = slice(k * B, (k + 1) * B)
        yy = y[sl, 0].detach().clone().requires_grad_(True)
        p = _pres_fwd(yy, u[:, cell])
        g = torch.zeros(B)
        if d_local is not None:
            g = g + d_local[sl, 0]
        if d_img is not None:
            g = g + d_img[:, cell]
        (gy,) = torch.autograd.grad((p * g).sum(), yy)
        d_y[sl, 0] = keep * gy


def relu_bwd(dh, h):
    dh.mul_((h > 0).float())


def broadcast_rows(row, rows, out):
    out.copy_(row.unsqueeze(0).expand(rows, -1))


# ------------------------------------------------------------------------------------------
def glimpse_fwd(image, z_where, cells, B, HW, Gh, Gw, out):
    if cells is None:
        out[:, :] = so.stn(image, z_where.view(-1, 4), [Gh, Gw]).flatten(1)
        return
    zw = z_where.view(B, HW, 4)
    for k, cell in enumerate(_rows(cells, B)):
        out[k * B:(k + 1) * B] = so.stn(image, zw[:, cell], [Gh, Gw]).flatten(1)


@torch.enable_grad()
def glimpse_bwd(image, z_where, cells, B, HW, Gh, Gw, d_out, d_zw_local, d_image=None):
    def one(zw_rows, g_rows):
        z = zw_rows.detach().clone().requires_grad_(True)
        im = image.detach().clone().requires_grad_(d_image is not None)
        o = so.stn(im, z, [Gh, Gw]).flatten(1)
        grads = torch.autograd.grad((o * g_rows).sum(), [z] + ([im] if d_image is not None else []))
        if d_image is not None:
            d_image.add_(grads[1])
        return grads[0]

    if cells is None:
        d_zw_local[:] = one(z_where.view(-1, 4), d_out)
        return
    zw = z_where.view(B, HW, 4)
    for k, cell in enumerate(_rows(cells, B)):
        sl = slice(k * B, (k + 1) * B)
        d_zw_local[sl] = one(zw[:, cell], d_out[sl])


def paste_fwd(image, z_where, Oh, Ow, out):
    out[:] = so.stn(image, z_where, [Oh, Ow], inverse=True)


@torch.enable_grad()
def paste_bwd(image, z_where, Oh, Ow, d_out, d_image, d_z_where):
    im = image.detach().clone().requires_grad_(True)
    z = z_where.detach().clone().requires_grad_(True)
    o = so.stn(im, z, [Oh, Ow], inverse=True)
    gi, gz = torch.autograd.grad((o * d_out).sum(), [im, z])
    d_image.add_(gi)
    d_z_where[:] = gz


# ------------------------------------------------------------------------------------------
def render_num_tiles(B, Ih, Iw):
    return B


def _render(logits, z_where, z_depth, z_pres, B, HW, C, G, Ih, Iw, scales):
    """models.py:481-540 from raw decoder logits [N,G,G,C+1]; returns (clamped recon, denominator)."""
    logits = logits.view(-1, G, G, C + 1)
    scale = torch.tensor([scales[0]] * C + [scales[1]])
    bias = torch.tensor([0.0] * C + [scales[2]])
    objects = so.clamped_sigmoid(logits * scale + bias, use_analytical=True)
    colour = objects[..., :C]
    alpha = objects[..., C] * z_pres.view(-1, 1, 1)
    imp = torch.clamp(alpha * z_depth.view(-1, 1, 1), min=0.01)
    stacked = torch.cat([colour, alpha[..., None], imp[..., None]], -1).permute(0, 3, 1, 2)
    warped = so.stn(stacked, z_where, [Ih, Iw], inverse=True).contiguous().view(B, HW, C + 2, Ih, Iw)
    imp_w = warped[:, :, C + 1:C + 2] + 1e-9
    S = imp_w.sum(dim=1, keepdim=True)
    img = warped[:, :, C:C + 1] * warped[:, :, :C]
    return torch.clamp((img * (imp_w / S)).sum(dim=1), min=0, max=1), S[:, 0, 0]


def render_fwd(logits, z_where, z_depth, z_pres, B, HW, C, G, Ih, Iw, scales, recon, denom, target, bce_partial, decoded=False):
    assert not decoded
    r, S = _render(logits, z_where, z_depth, z_pres, B, HW, C, G, Ih, Iw, scales)
    recon[:] = r
    if denom is not None:
        denom[:] = S
    if bce_partial is not None:
        bce_partial[:] = F.binary_cross_entropy(r, target, reduction="none").sum(dim=(1, 2, 3))


@torch.enable_grad()
def render_bwd(logits, z_where, z_depth, z_pres, B, HW, C, G, Ih, Iw, scales, recon, denom, d_recon, target, bce_scale,
               gs_ws, d_logits, d_z_where, d_z_depth, d_z_pres, decoded=False):
    assert not decoded
    leaves = [t.detach().clone().requires_grad_(True) for t in (logits, z_where, z_depth, z_pres)]
    r, _ = _render(*leaves, B, HW, C, G, Ih, Iw, scales)
    total = 0
    if d_recon is not None:
        total = total + (r * d_recon).sum()
    if target is not None:
        s = 1.0 if bce_scale is None else bce_scale[0]
        total = total + s * F.binary_cross_entropy(r, target, reduction="sum")
    grads = torch.autograd.grad(total, leaves)
    for dst, g in zip((d_logits, d_z_where, d_z_depth, d_z_pres), grads):
        dst[:] = g.view(dst.shape)


# ------------------------------------------------------------------------------------------
_KL_COLS = lambda A: dict(cy_logit=(0, 1), cx_logit=(1, 2), height_logit=(2, 3), width_logit=(3, 4), attr=(4, 4 + A),
                          depth_logit=(4 + A, 5 + A))


def _kl(dmean, dstd, pres, prior_mean, prior_std, count_dist0, A):
    """models.py:169-262 on image-major tensors; the count prior arrives precomputed."""
    from torch.distributions import Normal
    from torch.distributions.kl import kl_divergence
    B, HW, D = dmean.shape
    kl = pres[..., None] * kl_divergence(Normal(dmean, dstd, validate_args=False), Normal(prior_mean, prior_std))
    support = torch.arange(HW + 1, dtype=torch.float32)
    dist = count_dist0.repeat(B, 1)
    count = torch.zeros(B, 1)
    obj_kl, pzs = [], []
    for i in range(HW):
        q = torch.clamp(support - count, min=0.0, max=(HW - i)) / (HW - i)
        p_z = torch.bmm(dist[:, None, :], q[:, :, None]).squeeze(-1)
        prob = pres[:, i:i + 1]
        obj_kl.append(prob * (so.safe_log(prob) - so.safe_log(p_z)) + (1 - prob) * (so.safe_log(1 - prob) - so.safe_log(1 - p_z)))
        pzs.append(p_z)
        sample = torch.round(prob.detach())
        d1 = (sample * q + (1 - sample) * (1 - q)) * dist
        dist = d1 / d1.sum(dim=1, keepdim=True).clamp(min=1e-6)
        count = count + sample
    kl_map = torch.cat([kl, torch.cat(obj_kl, 1)[..., None]], -1)
    cols = list(_KL_COLS(A).values()) + [(D, D + 1)]
    sums = torch.stack([kl_map[..., lo:hi].sum(dim=(1, 2)) for lo, hi in cols], 1)
    return kl_map, torch.cat(pzs, 1), sums


def kl_fwd(dmean, dstd, pres, prior_mean, prior_std, count_dist0, B, HW, A, kl_map, p_z, kl_sums):
    m, p, s = _kl(dmean, dstd, pres, prior_mean, prior_std, count_dist0, A)
    kl_map[:], p_z[:], kl_sums[:] = m.detach(), p.detach(), s.detach()


@torch.enable_grad()
def kl_bwd(dmean, dstd, pres, prior_mean, prior_std, kl_map, p_z, d_sums, B, HW, A, d_dmean, d_dstd, d_pres):
    leaves = [t.detach().clone().requires_grad_(True) for t in (dmean, dstd, pres)]
    # the count prior only enters through p_z, which carries no gradient (round() in the scan)
    D = dmean.shape[-1]
    from torch.distributions import Normal
    from torch.distributions.kl import kl_divergence
    m, s, p = leaves
    kl = p[..., None] * kl_divergence(Normal(m, s, validate_args=False), Normal(prior_mean, prior_std))
    okl = p * (so.safe_log(p) - so.safe_log(p_z)) + (1 - p) * (so.safe_log(1 - p) - so.safe_log(1 - p_z))
    full = torch.cat([kl, okl[..., None]], -1)
    cols = list(_KL_COLS(A).values()) + [(D, D + 1)]
    sums = torch.stack([full[..., lo:hi].sum(dim=(1, 2)) for lo, hi in cols], 1)
    g = torch.autograd.grad((sums * d_sums).sum(), leaves)
    d_dmean[:], d_dstd[:], d_pres[:] = g
