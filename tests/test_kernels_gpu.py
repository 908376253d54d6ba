"""Parity tests proper (need a B200): every C-ABI kernel against the CPU oracle on the same seeded
inputs, and against the golden vectors produced by the unmodified reference.

The per-kernel checker is ``tests/cpu_kernel_mock.py`` — the oracle's torch-CPU arithmetic behind
the same call signature as the binding — so each test calls the SAME function twice: once through
``spair_pytorch_b200.kernels`` on CUDA tensors (the C-ABI), once through the checker on CPU copies.
Tolerance: rtol 1e-4 / atol 1e-5 in fp32 (BASELINE.json north_star), written in helpers.RTOL/ATOL.
"""
import ctypes

import numpy as np
import pytest
import torch

from tests import cpu_kernel_mock as M
from tests import helpers
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def K():
    from spair_pytorch_b200 import kernels
    return kernels


def gen(seed):
    return torch.Generator().manual_seed(seed)


def both(fn_name, args_cpu, out_names):
    """Call kernel ``fn_name`` on CUDA copies of ``args_cpu`` and the checker on the CPU originals;
    returns {name: (cuda_result, cpu_result)} for the output tensors."""
    def to_dev(v):
        if isinstance(v, torch.Tensor):
            return v.to(DEV)
        if isinstance(v, (tuple, list)) and any(isinstance(e, torch.Tensor) for e in v):
            return type(v)(to_dev(e) for e in v)
        return v

    args_gpu = {k: to_dev(v) for k, v in args_cpu.items()}
    getattr(K(), fn_name)(**args_gpu)
    torch.cuda.synchronize()
    getattr(M, fn_name)(**args_cpu)
    return {n: (args_gpu[n], args_cpu[n]) for n in out_names}


def geom_default(I=128, ppc=12, anchor=48.0):
    return K().BoxGeom(yx_scale=2.0, yx_min=-0.5, hw_scale=1.0, hw_min=0.0, anchor=anchor, img_h=float(I), img_w=float(I),
                       cell_ratio_y=ppc / I, cell_ratio_x=ppc / I)


# ------------------------------------------------------------------------------------------
# G: glimpse / paste (stn both directions)
# ------------------------------------------------------------------------------------------
def _stn_golden():
    g = helpers.load_golden("stn.npz")
    img = torch.from_numpy(g["image_u8"].astype(np.float32) / 255.0)
    return g, img, torch.from_numpy(g["z_where"])


@pytest.mark.parametrize("G", [45, 28])
def test_glimpse_matches_reference_golden(G):
    """stn() forward + both gradients against the reference's own stn (incl. the notebook round trip
    box [.5,.5,.8,.8], whose 102-px window takes the non-staged path, and a border-clamped box)."""
    from spair_pytorch_b200.modules import stn
    from tests.golden.make_golden import stn_cotangent
    g, img, zw = _stn_golden()
    im = img.to(DEV).requires_grad_(True)
    z = zw.to(DEV).requires_grad_(True)
    out = stn(im, z, [G, G], DEV)
    (out * stn_cotangent(out.shape, 1000 + G).to(DEV)).sum().backward()
    nm = "g%d" % G
    assert_close(out, g[nm + "/out"], "glimpse")
    assert_close(z.grad, g[nm + "/d_z_where"], "d z_where", atol=1e-5 + 1e-4 * float(np.abs(g[nm + "/d_z_where"]).max()))
    assert_close(im.grad, g[nm + "/d_image"], "d image")


def test_paste_matches_reference_golden():
    from spair_pytorch_b200.modules import stn
    from tests.golden.make_golden import stn_cotangent
    g, _, zw = _stn_golden()
    gl = torch.from_numpy(g["g45/out"]).to(DEV).requires_grad_(True)
    z = zw.to(DEV).requires_grad_(True)
    back = stn(gl, z, [128, 128], DEV, inverse=True)
    (back * stn_cotangent(back.shape, 2045).to(DEV)).sum().backward()
    assert_close(back, g["g45/inv_out"], "paste")
    assert_close(gl.grad, g["g45/inv_d_image"], "d glimpse")
    assert_close(z.grad, g["g45/inv_d_z_where"], "d z_where", atol=1e-5 + 1e-4 * float(np.abs(g["g45/inv_d_z_where"]).max()))


@pytest.mark.parametrize("per_object", [False, True], ids=["image-resident", "per-object"])
@pytest.mark.parametrize("C,I,G,B,Hc", [(1, 128, 28, 4, 11), (3, 64, 14, 3, 8), (1, 40, 8, 2, 5), (2, 50, 7, 2, 4), (1, 36, 40, 2, 3), (1, 37, 12, 2, 3)])
def test_glimpse_wavefront_rows_vs_oracle(C, I, G, B, Hc, per_object, monkeypatch):
    """Wavefront addressing (row r = k*B + b samples image b with z_where[b, cells[k]]), forward and
    the z_where gradient; includes boxes hanging over the border and I % 4 != 0 (non-staged path).  Both schedules of
    csrc/glimpse.cu: the image-resident kernel (whole image in shared memory by TMA bulk copy, a warp per glimpse) and the
    per-object kernel it replaces on this call shape (kept for images that do not fit and for the generic stn() API)."""
    if per_object:
        monkeypatch.setenv("SPAIR_GLIMPSE_PER_OBJECT", "1")
    HW = Hc * Hc
    rs = gen(C * 100 + I)
    image = torch.rand(B, C, I, I, generator=rs)
    zw = torch.rand(B, HW, 4, generator=rs)
    zw[..., 2:] = 0.02 + 0.36 * zw[..., 2:]
    zw[:, 0, :2] = torch.tensor([0.01, 0.99])
    cells = torch.tensor([0, 3, HW - 1, 7], dtype=torch.int32)
    n = cells.numel() * B
    r = both("glimpse_fwd", dict(image=image, z_where=zw, cells=cells, B=B, HW=HW, Gh=G, Gw=G,
                                 out=torch.zeros(n, C * G * G)), ["out"])
    assert_close(*r["out"], "glimpse rows")
    d_out = torch.randn(n, C * G * G, generator=rs)
    r = both("glimpse_bwd", dict(image=image, z_where=zw, cells=cells, B=B, HW=HW, Gh=G, Gw=G, d_out=d_out,
                                 d_zw_local=torch.zeros(n, 4)), ["d_zw_local"])
    ref = r["d_zw_local"][1]
    assert_close(*r["d_zw_local"], "d z_where rows", atol=1e-5 + 1e-4 * float(ref.abs().max()))


def test_glimpse_constant_image_property():
    """Size-independent property at config-D size: a glimpse of a constant image is that constant
    (bilinear weights sum to 1, border padding), for every box."""
    B, C, I, G, HW = 8, 3, 256, 28, 1024
    image = torch.full((B, C, I, I), 0.625, device=DEV)
    zw = torch.rand(B, HW, 4, device=DEV)
    cells = torch.arange(0, HW, 37, dtype=torch.int32, device=DEV)
    out = torch.empty(cells.numel() * B, C * G * G, device=DEV)
    K().glimpse_fwd(image, zw, cells, B, HW, G, G, out)
    assert float((out - 0.625).abs().max()) <= 1e-6


# ------------------------------------------------------------------------------------------
# R: fused renderer
# ------------------------------------------------------------------------------------------
def _render_cuda(lg, zw, zd, zp, target, B, HW, C, G, I):
    from spair_pytorch_b200 import ops
    leaves = [t.clone().to(DEV).requires_grad_(True) for t in (lg, zw, zd, zp)]
    recon, bce, _ = ops.RenderFunction.apply(leaves[0], leaves[1], leaves[2], leaves[3],
                                             None if target is None else target.to(DEV), B, HW, C, G, I, I, (2.0, 0.1, 5.0))
    return leaves, recon, bce


def test_render_reference_scene():
    """The scene of the reference's spair/test/test_renderer.py:8-36 (every object exactly on its cell,
    saturated logits, 2 channels) against the reference's own _render output."""
    g = helpers.load_golden("render.npz")
    B, Hc, G, C, I = 2, 11, 28, 2, 128
    logits = torch.full((B * Hc * Hc, G, G, C + 1), -1000.0)
    logits[:, 10:18, :, 0] = 1000.0
    logits[:, :, 10:18, 1] = 1000.0
    logits[..., -1] = 1000.0
    zw = torch.from_numpy(g["scene/z_where"]).permute(0, 2, 3, 1).reshape(-1, 4)
    ones = torch.ones(B * Hc * Hc)
    _, recon, _ = _render_cuda(logits, zw, ones, ones, None, B, Hc * Hc, C, G, I)
    assert_close(recon, g["scene/recon"], "scene canvas")


@pytest.mark.parametrize("nm", ["randA", "randRGB"])
def test_render_matches_reference_golden(nm):
    """Fused render fwd + bwd (with fused BCE) against the reference's _render + F.binary_cross_entropy."""
    from tests.golden.make_golden import RENDER_RANDOM, render_random_inputs
    g = helpers.load_golden("render.npz")
    C, I, Hc, G = RENDER_RANDOM[nm][1]
    lg, zw, zd, zp, target = render_random_inputs(nm)
    B, HW = 2, Hc * Hc
    leaves, recon, bce = _render_cuda(lg, zw.permute(0, 2, 3, 1).reshape(-1, 4), zd.reshape(-1), zp.reshape(-1), target,
                                      B, HW, C, G, I)
    bce.backward()
    assert_close(recon, g[nm + "/recon"], "canvas")
    assert_close(bce, g[nm + "/bce"], "bce")
    dl = leaves[0].grad.flatten().cpu()
    want = torch.from_numpy(g[nm + "/d_logits_val"])
    scale = float(g[nm + "/d_logits_stat"][1]) / np.sqrt(dl.numel())
    assert_close(dl[torch.from_numpy(g[nm + "/d_logits_idx"])], want, "d logits", atol=1e-5 + 1e-4 * scale)
    assert abs(float(dl.double().norm()) - g[nm + "/d_logits_stat"][1]) <= 1e-4 * g[nm + "/d_logits_stat"][1]
    for key, leaf, shape in (("d_z_where", leaves[1], (B, Hc, Hc, 4)), ("d_z_depth", leaves[2], (B, Hc, Hc, 1)),
                             ("d_z_pres", leaves[3], (B, Hc, Hc, 1))):
        want = torch.from_numpy(g[nm + "/" + key])
        got = leaf.grad.reshape(shape).permute(0, 3, 1, 2)
        assert_close(got, want, key, atol=1e-5 + 1e-4 * float(want.abs().max()))


@pytest.mark.parametrize("C,I,Hc,G,B", [(1, 40, 5, 8, 3), (4, 48, 6, 10, 2), (3, 72, 9, 14, 2)])
def test_render_vs_oracle_random(C, I, Hc, G, B):
    """More shapes (incl. canvases that are not a multiple of the 32x16 tile, C=4, boxes partly off-canvas)
    against the oracle's materialised render."""
    HW, rs = Hc * Hc, gen(C + I)
    N = B * HW
    zw = torch.rand(N, 4, generator=rs)
    zw[:, :2] = zw[:, :2] * 1.2 - 0.1
    zw[:, 2:] = 0.08 + 0.5 * zw[:, 2:]
    args = dict(logits=torch.randn(N, G, G, C + 1, generator=rs), z_where=zw, z_depth=4 * torch.rand(N, generator=rs),
                z_pres=torch.rand(N, generator=rs), B=B, HW=HW, C=C, G=G, Ih=I, Iw=I, scales=(2.0, 0.1, 5.0))
    target = torch.rand(B, C, I, I, generator=rs)
    f = both("render_fwd", dict(args, recon=torch.zeros(B, C, I, I), denom=torch.zeros(B, I, I), target=None,
                                bce_partial=None), ["recon", "denom"])
    assert_close(*f["recon"], "canvas")
    assert_close(*f["denom"], "denominator")
    recon_cpu, denom_cpu = f["recon"][1], f["denom"][1]
    bw = dict(args, recon=recon_cpu, denom=denom_cpu, d_recon=torch.randn(B, C, I, I, generator=rs), target=target,
              bce_scale=torch.tensor([0.5]), gs_ws=torch.zeros(B, C + 1, I, I), d_logits=torch.zeros(N, G, G, C + 1),
              d_z_where=torch.zeros(N, 4), d_z_depth=torch.zeros(N), d_z_pres=torch.zeros(N))
    r = both("render_bwd", bw, ["d_logits", "d_z_where", "d_z_depth", "d_z_pres"])
    for k, (got, want) in r.items():
        assert_close(got, want, k, atol=1e-5 + 1e-4 * float(want.abs().max()))


def test_render_properties_full_size():
    """Size-independent properties at BASELINE config-C size (B=64 slice of it; 16x16 cells, G=14):
    (1) z_pres = 0 gives an exactly black canvas; (2) the canvas does not depend on the order of the
    objects within an image (importance-normalised sum) up to fp32 summation order; (3) bitwise
    run-to-run determinism of forward and backward."""
    B, C, I, Hc, G = 64, 1, 128, 16, 14
    HW, N = Hc * Hc, 64 * 256
    rs = torch.Generator(device=DEV).manual_seed(0)
    lg = torch.randn(N, G, G, C + 1, device=DEV, generator=rs)
    zw = torch.rand(N, 4, device=DEV, generator=rs)
    zw[:, 2:] = (12 + 36 * zw[:, 2:]) / I
    zd, zp = 4 * torch.rand(N, device=DEV, generator=rs), torch.rand(N, device=DEV, generator=rs)
    target = torch.rand(B, C, I, I, device=DEV, generator=rs)
    (_, _, _, _), black, _ = _render_cuda(lg, zw, zd, torch.zeros(N), None, B, HW, C, G, I)
    assert float(black.abs().max()) == 0.0
    leaves, recon, bce = _render_cuda(lg, zw, zd, zp, target, B, HW, C, G, I)
    bce.backward()
    perm = torch.randperm(HW, device=DEV)
    idx = (torch.arange(B, device=DEV)[:, None] * HW + perm[None, :]).reshape(-1)
    _, recon_p, _ = _render_cuda(lg[idx], zw[idx], zd[idx], zp[idx], None, B, HW, C, G, I)
    assert_close(recon_p, recon, "object-order invariance")
    leaves2, recon2, bce2 = _render_cuda(lg, zw, zd, zp, target, B, HW, C, G, I)
    bce2.backward()
    assert torch.equal(recon, recon2) and torch.equal(bce, bce2)
    for a, b in zip(leaves, leaves2):
        assert torch.equal(a.grad, b.grad)


# ------------------------------------------------------------------------------------------
# K: KL terms
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("HW,A,B,step", [(25, 50, 3, 1), (121, 50, 4, 1001), (256, 50, 2, 5000), (1024, 50, 2, 2500)])
def test_kl_vs_oracle(HW, A, B, step):
    from oracle import spair_oracle as so
    rs = gen(HW)
    D = 4 + A + 1
    cfg = so.OracleConfig()
    _, cd0, _ = so.count_prior_distribution(step, HW, cfg)
    pm, ps = torch.zeros(D), torch.ones(D)
    pm[2:4], ps[2:4] = 7.0, 0.5
    base = dict(dmean=torch.randn(B, HW, D, generator=rs), dstd=0.05 + 1.9 * torch.rand(B, HW, D, generator=rs),
                pres=torch.rand(B, HW, generator=rs), prior_mean=pm, prior_std=ps)
    base["pres"][0, :5] = torch.tensor([0.5, 1.5e-8, 1.0 - 1e-7, 0.49999, 0.50001])   # round-half-even / saturation edges
    f = both("kl_fwd", dict(base, count_dist0=cd0, B=B, HW=HW, A=A, kl_map=torch.zeros(B, HW, D + 1),
                            p_z=torch.zeros(B, HW), kl_sums=torch.zeros(B, 7)), ["kl_map", "p_z", "kl_sums"])
    assert_close(*f["kl_map"], "KL maps")
    assert_close(*f["p_z"], "p_z", atol=2e-6)
    assert_close(*f["kl_sums"], "KL sums", atol=1e-3)   # sums of up to 51k terms of magnitude ~1
    d_sums = torch.rand(B, 7, generator=rs)
    r = both("kl_bwd", dict(base, kl_map=f["kl_map"][1], p_z=f["p_z"][1], d_sums=d_sums, B=B, HW=HW, A=A,
                            d_dmean=torch.zeros(B, HW, D), d_dstd=torch.zeros(B, HW, D), d_pres=torch.zeros(B, HW)),
             ["d_dmean", "d_dstd", "d_pres"])
    for k, (got, want) in r.items():
        assert_close(got, want, k, atol=1e-5 + 1e-4 * float(want.abs().max()) * (k == "d_pres"))


# ------------------------------------------------------------------------------------------
# L0-L4: context and heads
# ------------------------------------------------------------------------------------------
def _cells_for(Hc, Wc, t, L=1):
    from spair_pytorch_b200.schedule import build_schedule
    s = build_schedule(Hc, Wc, L)
    return s, torch.from_numpy(np.ascontiguousarray(s.cells_of(t)))


@pytest.mark.parametrize("L", [1, 2])
def test_context_gather_and_grad_vs_oracle(L):
    from spair_pytorch_b200.schedule import build_schedule
    B, F, Hc, Wc, A = 3, 10, 5, 6, 7
    HW, E = Hc * Wc, A + 6
    s = build_schedule(Hc, Wc, L)
    n_nb = len(s.offsets)
    rs = gen(L)
    base = dict(feat=torch.randn(B, F, Hc, Wc, generator=rs), box=torch.randn(B, HW, 4, generator=rs),
                attr=torch.randn(B, HW, A, generator=rs), depth=torch.randn(B, HW, generator=rs),
                pres=torch.randn(B, HW, generator=rs), edge=torch.randn(E, generator=rs))
    width = F + n_nb * E
    for t in (0, 3, s.n_wavefronts - 1):
        cells = torch.from_numpy(np.ascontiguousarray(s.cells_of(t)))
        n = cells.numel() * B
        dst = [torch.zeros(n, width + k) for k in (0, 5, 9)]
        args = dict(base, cells=cells, offsets=s.offsets, dsts=tuple(dst))
        r = both("context_gather_fwd", args, ["dsts"])
        for a, b in zip(*r["dsts"]):
            assert torch.equal(a.cpu()[:, :width], b[:, :width])
        dxs = tuple(torch.randn(HW * B, width + k, generator=rs) for k in (0, 5, 9))
        r = both("context_grad_gather", dict(dxs=dxs, col0=F, cells=cells, wf_pos=torch.from_numpy(s.wf_pos.copy()),
                                             offsets=s.offsets, B=B, Hc=Hc, Wc=Wc, A=A, out=torch.zeros(n, E)), ["out"])
        assert_close(*r["out"], "context grad gather")


def test_box_head_vs_oracle():
    B, Hc, Wc, P = 5, 11, 11, 6
    HW = Hc * Wc
    s, cells = _cells_for(Hc, Wc, 12)
    n = cells.numel() * B
    rs = gen(3)
    y = 3 * torch.randn(n, 8 + P, generator=rs)
    y[0, :8] = torch.tensor([11.0, -11.0, 10.0, -10.0, 12.0, -12.0, 0.0, 9.99])     # clamp edges
    eps = torch.randn(B, HW, 4, generator=rs)
    geom = geom_default()
    out = dict(box=torch.zeros(B, HW, 4), z_where=torch.zeros(B, HW, 4), dmean=torch.zeros(B, HW, 9), dstd=torch.zeros(B, HW, 9),
               xdsts=(torch.zeros(n, 7), torch.zeros(n, 4)), pt_dst=torch.zeros(n, P + 2))
    r = both("box_head_fwd", dict(y=y, eps=eps, cells=cells, B=B, HW=HW, Wc=Wc, geom=geom, n_pt=P, **out), list(out))
    for k in ("box", "z_where", "dmean", "dstd", "pt_dst"):
        assert_close(*r[k], k)
    for a, b in zip(*r["xdsts"]):
        assert_close(a, b, "box copy")
    for wheel in (0.0, 1.0, 0.25):
        grads = dict(d_boxes=(torch.randn(n, 6, generator=rs), None, torch.randn(n, 4, generator=rs)),
                     d_zw_local=torch.randn(n, 4, generator=rs), d_zw_img=torch.randn(B, HW, 4, generator=rs),
                     d_dmean=torch.randn(B, HW, 9, generator=rs), d_dstd=torch.randn(B, HW, 9, generator=rs),
                     d_pt_src=torch.randn(n, P + 1, generator=rs))
        r = both("box_head_bwd", dict(y=y, eps=eps, cells=cells, B=B, HW=HW, Wc=Wc, geom=geom, wheel=torch.tensor([wheel]),
                                      ld_dist=9, n_pt=P, d_y=torch.zeros(n, 8 + P), **grads), ["d_y"])
        assert_close(*r["d_y"], "box head grad (wheel %.2f)" % wheel, atol=2e-5)


@pytest.mark.parametrize("W,squash,P", [(50, 0, 0), (1, 1, 100)])
def test_normal_head_vs_oracle(W, squash, P):
    B, Hc, Wc = 4, 11, 11
    HW, D = Hc * Wc, 55
    s, cells = _cells_for(Hc, Wc, 9)
    n = cells.numel() * B
    rs = gen(W)
    y = 3 * torch.randn(n, 2 * W + P, generator=rs)
    eps = torch.randn(B, HW, W, generator=rs).squeeze(-1)
    col = 4 if W > 1 else D - 1
    dmean, dstd = torch.zeros(B, HW, D), torch.zeros(B, HW, D)
    out = torch.zeros(B, HW, W).squeeze(-1)
    a = dict(y=y, W=W, eps=eps, cells=cells, B=B, HW=HW, squash=squash, scale=4.0, out=out, ld_dist=D,
             xdsts=(torch.zeros(n, W + 3), torch.zeros(n, W)), n_pt=P, pt_dst=torch.zeros(n, P) if P else None)
    # views into the [B,HW,D] maps start at this head's column
    args_cpu = dict(a, dmean_view=dmean[..., col:], dstd_view=dstd[..., col:])
    dm_g, ds_g = dmean.to(DEV), dstd.to(DEV)
    args_gpu = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in a.items()}
    args_gpu["xdsts"] = tuple(t.to(DEV) for t in a["xdsts"])
    K().normal_head_fwd(**dict(args_gpu, dmean_ptr_view=dm_g[..., col:], dstd_ptr_view=ds_g[..., col:]))
    M.normal_head_fwd(**args_cpu)
    assert_close(args_gpu["out"], out, "head output")
    assert_close(dm_g, dmean, "mean map")
    assert_close(ds_g, dstd, "std map")
    for x_g, x_c in zip(args_gpu["xdsts"], a["xdsts"]):
        assert_close(x_g, x_c, "copy")
    if P:
        assert_close(args_gpu["pt_dst"], a["pt_dst"], "passthrough")
    d_dm, d_ds = torch.randn(B, HW, D, generator=rs), torch.randn(B, HW, D, generator=rs)
    b = dict(y=y, W=W, eps=eps, cells=cells, B=B, HW=HW, squash=squash, scale=4.0, wheel=torch.tensor([0.0]) if squash else None,
             d_outs=(torch.randn(n, W, generator=rs), torch.randn(n, W + 2, generator=rs)),
             d_out_img=torch.randn(B, HW, W, generator=rs).squeeze(-1), ld_dist=D, n_pt=P,
             d_pt_src=torch.randn(n, P, generator=rs) if P else None, d_y=torch.zeros(n, 2 * W + P))
    b_gpu = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
    b_gpu["d_outs"] = tuple(t.to(DEV) for t in b["d_outs"])
    K().normal_head_bwd(**dict(b_gpu, d_dmean_view=d_dm.to(DEV)[..., col:], d_dstd_view=d_ds.to(DEV)[..., col:]))
    M.normal_head_bwd(**dict(b, d_dmean_view=d_dm[..., col:], d_dstd_view=d_ds[..., col:]))
    assert_close(b_gpu["d_y"], b["d_y"], "normal head grad", atol=2e-5)


def test_pres_head_vs_oracle():
    B, Hc, Wc = 6, 11, 11
    HW = Hc * Wc
    s, cells = _cells_for(Hc, Wc, 10)
    n = cells.numel() * B
    rs = gen(9)
    y = 6 * torch.randn(n, 3, generator=rs)
    u = torch.rand(B, HW, generator=rs)
    u[0, int(cells[0])] = 0.0
    u[1, int(cells[0])] = 1.0 - 6e-8
    r = both("pres_head_fwd", dict(y=y, u=u, cells=cells, B=B, HW=HW, pres=torch.zeros(B, HW)), ["pres"])
    assert_close(*r["pres"], "z_pres")
    r = both("pres_head_bwd", dict(y=y, u=u, cells=cells, B=B, HW=HW, wheel=torch.tensor([0.0]),
                                   d_local=torch.randn(n, 2, generator=rs), d_img=torch.randn(B, HW, generator=rs),
                                   d_y=torch.zeros(n, 3)), ["d_y"])
    assert_close(r["d_y"][0][:, 0], r["d_y"][1][:, 0], "pres head grad")


def test_relu_bwd():
    rs = gen(1)
    h = torch.relu(torch.randn(300, 100, generator=rs))
    dh = torch.randn(300, 128, generator=rs)
    got = dh.to(DEV)
    K().relu_bwd(got[:, :100], h.to(DEV))
    want = dh.clone()
    want[:, :100] *= (h > 0).float()
    assert torch.equal(got.cpu(), want)


def test_invalid_arguments_are_rejected_not_launched():
    k = K()
    z = torch.zeros(4, 4, device=DEV)
    with pytest.raises(k.SpairKernelError):
        k.render_fwd(torch.zeros(4, 40, 40, 2, device=DEV), z, z[:, 0].contiguous(), z[:, 0].contiguous(), 1, 4, 1, 40, 64, 64,
                     (2.0, 0.1, 5.0), torch.zeros(1, 1, 64, 64, device=DEV), torch.zeros(1, 64, 64, device=DEV), None, None)
    with pytest.raises(k.SpairKernelError):
        k.glimpse_fwd(torch.zeros(1, 1, 8, 8), torch.zeros(1, 4), None, 1, 1, 4, 4, torch.zeros(1, 16))   # CPU tensors


@pytest.mark.parametrize("shapes", [[(100, 324), (100, 100), (108, 100)], [(1, 100), (256, 784), (7, 5)]])
def test_sweep_weight_packing_is_exact(shapes):
    """spair_sweep_pack_weights: fwd[(g*N + n)*4 + j] = W[n][4g+j], bwd[(g*K + k)*4 + j] = W[4g+j][k], zero padded."""
    rs = gen(5)
    ws = [torch.randn(n, k, generator=rs) for n, k in shapes]
    packed = K().PackedSweepWeights([w.to(DEV) for w in ws])
    for w, f, b in zip(ws, packed.fwd, packed.bwd):
        n, k = w.shape
        k4, n4 = (k + 3) // 4, (n + 3) // 4
        wf = torch.zeros(n, 4 * k4)
        wf[:, :k] = w
        assert torch.equal(f.cpu().view(k4, n, 4), wf.view(n, k4, 4).permute(1, 0, 2)), "forward layout"
        wb = torch.zeros(4 * n4, k)
        wb[:n] = w
        assert torch.equal(b.cpu().view(n4, k, 4), wb.view(n4, 4, k).permute(0, 2, 1)), "backward layout"


def _tf32_round_bits(x):
    """(bits + 0x1000) & 0xffffe000 on the raw fp32 pattern: nearest TF32, ties away from zero (csrc/sweep_tc.cuh tf32_round)."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)


@pytest.mark.parametrize("shapes", [
    [(100, 324), (100, 100), (108, 100), (256, 784), (128, 256), (100, 128), (100, 478), (100, 100), (102, 100), (100, 479), (100, 100), (1, 100)],
    [(40, 70), (33, 40), (12, 33), (256, 2352), (130, 256), (20, 130), (64, 90), (64, 64), (6, 64), (64, 91), (31, 64), (1, 31)]])
def test_tensor_core_sweep_weight_streams_are_exact(shapes):
    """spair_sweep_tc_pack: both weight streams of the tensor-core sweeps, stage by stage in the order tc::mma_loop consumes
    them (layer -> accumulator group of 8 feature tiles -> chunk of 8 k-blocks -> tile -> k-block); a stage = 128 features x
    32 reduction indices as the TF32 hi tile then the lo tile, rows of 128 bytes with the 16-byte chunks XOR-swizzled by the
    row (the K-major SWIZZLE_128B operand image), zero padded.  hi = TF32(w), lo = TF32(w - hi); the backward stream holds
    W^T tiles in the order obj, z, enc, box x layers 2, 1, 0.  The second shape set has a 2352-wide layer (config D's
    encoder: 19 feature tiles backward = three accumulator groups) and ragged tails everywhere."""
    rs = gen(9)
    ws = [torch.randn(n, k, generator=rs) * 0.3 for n, k in shapes]
    packed = K().PackedSweepWeightsTC([w.to(DEV) for w in ws])
    fl = torch.arange(128).view(128, 1)
    kl = torch.arange(32).view(1, 32)
    pos = (fl * 32 + ((((kl >> 2) ^ fl) & 7) << 2) + (kl & 3)).flatten()       # where element (fl, kl) of a tile lives
    for backward, stream in ((False, packed.fwd.cpu()), (True, packed.bwd.cpu())):
        stage = 0
        for e in range(12):
            l = 3 * (3 - e // 3) + (2 - e % 3) if backward else e
            W = ws[l].t().contiguous() if backward else ws[l]                   # [features, reduction]
            M, Kd = W.shape
            KB, tiles, chunks = (Kd + 31) // 32, (M + 127) // 128, (Kd + 255) // 256
            Wp = torch.zeros(tiles * 128, KB * 32)
            Wp[:M, :Kd] = W
            hi = _tf32_round_bits(Wp)
            lo = _tf32_round_bits(Wp - hi)
            for g0 in range(0, tiles, 8):
                for c in range(chunks):
                    for mt in range(g0, min(g0 + 8, tiles)):
                        for kb in range(8 * c, min(8 * c + 8, KB)):
                            got = stream[stage * 8192:(stage + 1) * 8192]
                            want_hi = hi[mt * 128:(mt + 1) * 128, kb * 32:(kb + 1) * 32].flatten()
                            want_lo = lo[mt * 128:(mt + 1) * 128, kb * 32:(kb + 1) * 32].flatten()
                            assert torch.equal(got[:4096][pos], want_hi), ("hi", backward, e, mt, kb)
                            assert torch.equal(got[4096:][pos], want_lo), ("lo", backward, e, mt, kb)
                            stage += 1
        assert stage * 8192 == stream.numel()


@pytest.mark.parametrize("B,C,I,stride,pad", [(3, 1, 128, 3, (9, 14, 9, 14)), (2, 3, 64, 2, (7, 7, 7, 7)), (2, 1, 37, 3, (2, 5, 1, 3)),
                                              (5, 3, 41, 1, (0, 0, 0, 0))])
def test_stem_conv_vs_torch_cpu(B, C, I, stride, pad):
    """Backbone stem (ZeroPad2d + Conv2d(C,128,4,stride) + ReLU, reference modules.py:44-66,86-87): forward map and
    weight / bias gradients against the same ops in torch fp32 on the CPU."""
    from spair_pytorch_b200 import ops
    rs = gen(11)
    x = torch.rand(B, C, I, I, generator=rs)
    w = (torch.randn(128, C, 4, 4, generator=rs) * 0.2).requires_grad_(True)
    b = (torch.randn(128, generator=rs) * 0.1).requires_grad_(True)
    want = torch.relu(torch.nn.functional.conv2d(torch.nn.functional.pad(x, pad), w, b, stride=stride))
    cot = torch.randn(want.shape, generator=rs)
    want.backward(cot)
    wd, bd = w.detach().to(DEV).requires_grad_(True), b.detach().to(DEV).requires_grad_(True)
    pl, pr, pt, pb = pad
    got = ops.StemConvFunction.apply(x.to(DEV), wd, bd, stride, pt, pl, want.shape[2], want.shape[3])
    got.backward(cot.to(DEV))
    assert_close(got, want, "stem forward")
    helpers.assert_close(wd.grad.cpu(), w.grad, "stem d_weight", atol=helpers.ATOL + helpers.RTOL * float(w.grad.norm()) / w.grad.numel() ** 0.5)
    helpers.assert_close(bd.grad.cpu(), b.grad, "stem d_bias", atol=helpers.ATOL + helpers.RTOL * float(b.grad.norm()) / b.grad.numel() ** 0.5)
    # channels-last variant (the layout the GEMM tail consumes): the same values, forward and backward, bit for bit
    wn, bn = w.detach().to(DEV).requires_grad_(True), b.detach().to(DEV).requires_grad_(True)
    got_cl = ops.StemConvFunction.apply(x.to(DEV), wn, bn, stride, pt, pl, want.shape[2], want.shape[3], True)
    assert got_cl.shape == (B, want.shape[2], want.shape[3], 128)
    got_cl.backward(cot.to(DEV).permute(0, 2, 3, 1).contiguous())
    assert torch.equal(got_cl.permute(0, 3, 1, 2), got), "channels-last stem forward"
    assert torch.equal(wn.grad, wd.grad) and torch.equal(bn.grad, bd.grad), "channels-last stem backward"


def test_backbone_uses_fused_stem_and_matches_library_path():
    """Backbone.forward routes layer 0 through the stem kernel on CUDA (and the later layers through the tcgen05 GEMM:
    9 more launches); same features and parameter gradients as the cuDNN path of the same module (image with
    requires_grad takes the library path)."""
    from spair_pytorch_b200 import kernels as kk
    from spair_pytorch_b200.modules import Backbone
    torch.backends.cudnn.allow_tf32 = False          # strict fp32 on the library path too (the model sets this itself)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(5)
    net = Backbone([1, 128, 128], 100).to(DEV)
    x = torch.rand(4, 1, 128, 128, device=DEV)
    n0 = kk.launch_count()
    y = net(x)
    assert kk.launch_count() > n0, "fused stem not used"
    y.square().sum().backward()
    g_fused = [p.grad.clone() for p in net.parameters()]
    net.zero_grad()
    xr = x.clone().requires_grad_(True)
    n0 = kk.launch_count()
    y_ref = net(xr)
    assert kk.launch_count() == n0, "library path expected when the image needs a gradient"
    y_ref.square().sum().backward()
    assert_close(y, y_ref, "backbone features")
    for (name, p), g in zip(net.named_parameters(), g_fused):
        helpers.assert_close(g, p.grad, "backbone grad " + name, atol=helpers.ATOL + helpers.RTOL * float(p.grad.norm()) / p.grad.numel() ** 0.5)


def test_render_saturated_canvas_matches_oracle():
    """A saturated object (sigmoid == 1.0f exactly for both channels, z_pres = 1): the composited value is a convex
    combination of alpha*colour <= 1, so the canvas reaches 1.0f (or 1 - ulp, depending on the rounding of the bilinear
    weights) and the clamp at 1 is active only through fp32 rounding.  render.cu treats recon == 1.0f as clamped in the
    backward where torch passes the gradient AT the bound; with a BCE gradient of (r - t) / max((1 - r) r, 1e-12) = +-1 or
    0 per pixel depending on that last bit, the reference's own gradient is rounding noise in this state (DESIGN.md,
    "Known, documented deviations").  Pinned here: the canvas equals the oracle's within tolerance (and is 1 within an
    ulp inside the object), every gradient is finite, and the logit gradient vanishes (sigmoid' == 0 at saturation) as in
    the oracle."""
    B, C, I, G, HW = 1, 1, 32, 8, 1
    logits = torch.empty(1, G, G, C + 1)
    logits[..., 0], logits[..., 1] = 40.0, 200.0              # sigma(2 * 40) == sigma(0.1 * 200 + 5) == 1.0f
    zw = torch.tensor([[0.5, 0.5, 0.6, 0.6]])
    zd, zp = torch.tensor([4.0]), torch.tensor([1.0])
    args = dict(logits=logits, z_where=zw, z_depth=zd, z_pres=zp, B=B, HW=HW, C=C, G=G, Ih=I, Iw=I, scales=(2.0, 0.1, 5.0))
    f = both("render_fwd", dict(args, recon=torch.zeros(B, C, I, I), denom=torch.zeros(B, I, I), target=None, bce_partial=None),
             ["recon", "denom"])
    recon_gpu, recon_cpu = f["recon"]
    assert float(recon_cpu.max()) >= 1.0 - 2e-7 and float(recon_gpu.max()) >= 1.0 - 2e-7 and float(recon_gpu.max()) <= 1.0
    assert_close(recon_gpu, recon_cpu, "saturated canvas")
    target = (recon_cpu > 0.5).float()
    bw = dict(args, recon=recon_cpu, denom=f["denom"][1], d_recon=None, target=target, bce_scale=None,
              gs_ws=torch.zeros(B, C + 1, I, I), d_logits=torch.zeros(1, G, G, C + 1), d_z_where=torch.zeros(1, 4),
              d_z_depth=torch.zeros(1), d_z_pres=torch.zeros(1))
    r = both("render_bwd", bw, ["d_logits", "d_z_where", "d_z_depth", "d_z_pres"])
    for k, (got, want) in r.items():
        assert torch.isfinite(got).all() and torch.isfinite(want).all(), k
    assert float(r["d_logits"][0].abs().max()) <= 1e-6 and float(r["d_logits"][1].abs().max()) <= 1e-6
