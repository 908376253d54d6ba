"""Generates tests/golden/*.npz from the UNMODIFIED reference (/root/reference).

Run inside the build container only (the reference is mounted there, not on the GPU box):

    python tests/golden/make_golden.py

Every fixture is produced by importing the real reference through ``oracle/ref_harness.py``
(stubs for tensorboardX / matplotlib / cycler, config mutated before import) and running its
own code: ``SPAIR.forward`` + ``loss.backward(retain_graph=True)`` (train.py:65-66),
``spair.modules.stn`` and ``SPAIR._render``.  Parameters are NOT stored: the reference model is
built under ``torch.manual_seed(3)`` (train.py:39) and the drop-in model reproduces the same
construction order, so the same seed gives the same parameters; per-parameter checksums are
stored so the tests can assert that before comparing anything else.

Large parameter gradients are stored as a fixed pseudo-random subsample (indices derived from
the parameter name) plus their sum and L2 norm, so the fixtures stay small.
"""
from __future__ import annotations

import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness as rh  # noqa: E402
from oracle import spair_oracle as so  # noqa: E402

GRAD_FULL_MAX = 2048
GRAD_SAMPLE = 1024

CONFIGS = {
    # name: (reference cfg overrides, oracle config factory, batch)
    "tiny": (dict(INPUT_IMAGE_SHAPE=[1, 40, 40], OBJECT_SHAPE=[8, 8], ANCHORBOX_SHAPE=[16, 16],
                  DEFAULT_BACKBONE_TOPOLOGY=rh.TOPOLOGY_CELL8, BATCH_SIZE=3), so.config_tiny, 3),
    "A": (dict(BATCH_SIZE=2), so.config_A, 2),
}


def grad_sample_indices(name: str, numel: int, n_sample: int = GRAD_SAMPLE) -> np.ndarray:
    if numel <= max(GRAD_FULL_MAX, n_sample):
        return np.arange(numel)
    rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return np.sort(rs.choice(numel, n_sample, replace=False))


def param_checksums(state_dict) -> dict:
    out = {}
    for k, v in state_dict.items():
        d = v.detach().double()
        out["psum/" + k] = np.array([d.sum().item(), d.abs().sum().item()])
    return out


def model_cases():
    for name, (overrides, cfg_fn, B) in CONFIGS.items():
        ns = rh.load_reference(overrides)
        net = rh.build_reference_model(ns, seed=3)
        ocfg = cfg_fn()
        x = so.scattered_sprites(B, ocfg.image_shape, seed=77, sprite_px=(6, 14) if name == "tiny" else (10, 20))
        for step in (1, 1001):
            noise_seed = 100 + step
            loss, recon, z_where, z_pres = rh.run_reference(net, x, step, noise_seed)
            # the reference keeps the rest as attributes / locals; recover them through the
            # pinned oracle run on the same noise (asserted bit-identical on the public outputs)
            params = so.params_from_state_dict(net.state_dict())
            noise = so.draw_noise(noise_seed, B, ocfg.grid, ocfg.n_attr)
            out = so.forward_backward(params, x, step, noise, ocfg)
            assert torch.equal(out["recon_x"], recon) and torch.equal(out["z_where"], z_where)
            assert torch.equal(out["z_pres"], z_pres) and torch.equal(out["loss"], loss)
            fx = dict(x=x.numpy(), step=np.array(step), noise_seed=np.array(noise_seed),
                      eps_where=noise.eps_where.numpy(), eps_attr=noise.eps_attr.numpy(),
                      eps_depth=noise.eps_depth.numpy(), u_pres=noise.u_pres.numpy(),
                      loss=loss.detach().numpy(), recon_loss=out["recon_loss"].detach().numpy(),
                      recon_x=recon.detach().numpy(), z_where=z_where.detach().numpy(),
                      z_pres=z_pres.detach().numpy(), z_attr=out["z_attr"].detach().numpy(),
                      z_depth=out["z_depth"].detach().numpy())
            for n, v in out["kl"].items():
                fx["kl/" + n] = v.detach().numpy()
            for n, v in out["kl_means"].items():
                fx["klmean/" + n] = v.detach().numpy()
            # reference's own distribution maps (models.py:122-125)
            for n, d in net.dist.items():
                fx["dist_mean/" + n] = d.loc.detach().numpy()
                fx["dist_std/" + n] = d.scale.detach().numpy()
                assert torch.equal(d.loc, out["dist_mean"][n])
            fx.update(param_checksums(net.state_dict()))
            # float64 evaluation of the same formulas: tells rounding noise of the reference's own fp32 backward
            # (and ReLU/clamp kink decisions) apart from genuine differences
            _, params64 = so.forward_backward_fp64(net.state_dict(), x, step, noise, ocfg)
            for k, p in net.named_parameters():
                if p.grad is None:
                    fx["gnone/" + k] = np.array(1)
                    continue
                g = p.grad.detach().flatten()
                idx = grad_sample_indices(k, g.numel())
                fx["gidx/" + k] = idx.astype(np.int64)
                fx["gval/" + k] = g[torch.from_numpy(idx)].numpy()
                fx["g64val/" + k] = params64[k].grad.detach().flatten()[torch.from_numpy(idx)].numpy()
                fx["gstat/" + k] = np.array([g.double().sum().item(), g.double().norm().item()])
            path = os.path.join(HERE, "model_%s_step%d.npz" % (name, step))
            np.savez_compressed(path, **fx)
            print("wrote", path, os.path.getsize(path) // 1024, "KiB  loss", loss.item())


def zw_all():
    return torch.tensor([[0.5, 0.5, 0.8, 0.8],          # notebook round trip
                         [0.3, 0.6, 0.25, 0.375],       # interior box
                         [0.02, 0.97, 0.3, 0.3],        # hangs over the border -> clamped taps
                         [0.7, 0.2, 0.05, 0.1]])        # small box


def stn_cotangent(shape, seed):
    """Cotangents are regenerated from a seed in the tests instead of being stored."""
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def stn_cases():
    """``spair.modules.stn`` both directions on procedural images (reference modules.py:216-273),
    including the scenario of test_notebook.ipynb:252-260 (z_where = [.5,.5,.8,.8], 45x45 cut,
    then pasted back at 128x128) on a procedural 3-channel image."""
    ns = rh.load_reference(dict(INPUT_IMAGE_SHAPE=[3, 128, 128]))
    dev = torch.device("cpu")
    g = torch.Generator().manual_seed(5)
    fx = {}
    img = so.scattered_sprites(4, (3, 128, 128), seed=9, max_sprites=12, sprite_px=(10, 40))
    img = (img + 0.1 * torch.rand(img.shape, generator=g)).clamp(0, 1)
    img = torch.round(img * 255) / 255            # stored as uint8; tests rebuild img = u8 / 255
    fx["image_u8"] = torch.round(img * 255).to(torch.uint8).numpy()
    fx["z_where"] = zw_all().numpy()
    zw = zw_all()
    for nm, G in (("g45", 45), ("g28", 28)):
        z = zw.clone().requires_grad_(True)
        im = img.clone().requires_grad_(True)
        out = ns.modules.stn(im, z, [G, G], dev)
        cot = stn_cotangent(out.shape, 1000 + G)
        (out * cot).sum().backward()
        fx.update({nm + "/out": out.detach().numpy(), nm + "/d_z_where": z.grad.numpy(), nm + "/d_image": im.grad.numpy()})
        if G != 45:
            continue
        # paste back (inverse=True appends one channel to out_dims but samples image.shape[1] channels)
        z2 = zw.clone().requires_grad_(True)
        gl = out.detach().clone().requires_grad_(True)
        back = ns.modules.stn(gl, z2, [128, 128], dev, inverse=True)
        cot2 = stn_cotangent(back.shape, 2000 + G)
        (back * cot2).sum().backward()
        fx.update({nm + "/inv_out": back.detach().numpy(),
                   nm + "/inv_d_z_where": z2.grad.numpy(), nm + "/inv_d_image": gl.grad.numpy()})
    path = os.path.join(HERE, "stn.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def _render_with_logits(ns, net, logits, z_where, z_depth, z_pres, x_like):
    """Runs the reference's own ``SPAIR._render`` (models.py:452-542) on externally supplied
    decoder outputs by replacing ``object_decoder`` with a function returning them."""
    class _Fixed(torch.nn.Module):
        def forward(self, _z):
            return logits.reshape(logits.shape[0], -1) * 1.0

    saved = net.object_decoder
    net.object_decoder = _Fixed()
    try:
        B, _, Hc, Wc = z_where.shape
        z_attr = torch.zeros(B, ns.cfg.N_ATTRIBUTES, Hc, Wc)
        net.global_step = 1
        return net._render(z_attr, z_where, z_depth, z_pres, x_like)
    finally:
        net.object_decoder = saved


RENDER_RANDOM = {
    # name: (reference cfg overrides, (C, I, Hc, G))
    "randA": (dict(BATCH_SIZE=2), (1, 128, 11, 28)),
    "randRGB": (dict(INPUT_IMAGE_SHAPE=[3, 64, 64], OBJECT_SHAPE=[14, 14], DEFAULT_BACKBONE_TOPOLOGY=rh.TOPOLOGY_CELL8,
                     BATCH_SIZE=2), (3, 64, 8, 14)),
}


def render_random_inputs(nm, B=2):
    """Seeded inputs of the random render scenes (regenerated in the tests, not stored):
    logits ~ N(0,1), centres ~ U(0,1), box side ~ U(12,48)/128 of the canvas, depth ~ U(0,4), pres ~ U(0,1)."""
    C, I, Hc, G = RENDER_RANDOM[nm][1]
    g = torch.Generator().manual_seed(21 + C)
    Wc = Hc
    zw = torch.empty(B, 4, Hc, Wc)
    zw[:, :2] = torch.rand(B, 2, Hc, Wc, generator=g)
    zw[:, 2:] = (12.0 + 36.0 * torch.rand(B, 2, Hc, Wc, generator=g)) / 128
    lg = torch.randn(B * Hc * Wc, G, G, C + 1, generator=g)
    zd = 4 * torch.rand(B, 1, Hc, Wc, generator=g)
    zp = torch.rand(B, 1, Hc, Wc, generator=g)
    target = so.scattered_sprites(B, (C, I, I), seed=3, sprite_px=(8, 20))
    return lg, zw, zd, zp, target


def render_cases():
    fx = {}
    # (a) the scene of spair/test/test_renderer.py:8-36 — B=2, 2 channels, every object exactly on its
    #     cell with size 1/11, colour logits -1000 except one stripe +1000, alpha logits +1000, depth = pres = 1
    ns = rh.load_reference(dict(INPUT_IMAGE_SHAPE=[2, 128, 128], BATCH_SIZE=2))
    net = rh.build_reference_model(ns)
    B, Hc, Wc, G, C = 2, 11, 11, 28, 2
    z_where = torch.empty(B, 4, Hc, Wc)
    for h in range(Hc):
        for w in range(Wc):
            z_where[:, 0, h, w] = (w + 0.5) / Wc
            z_where[:, 1, h, w] = (h + 0.5) / Hc
    z_where[:, 2:] = 1.0 / 11
    logits = torch.full((B * Hc * Wc, G, G, C + 1), -1000.0)
    logits[:, 10:18, :, 0] = 1000.0
    logits[:, :, 10:18, 1] = 1000.0
    logits[..., -1] = 1000.0
    z_depth = torch.ones(B, 1, Hc, Wc)
    z_pres = torch.ones(B, 1, Hc, Wc)
    recon = _render_with_logits(ns, net, logits, z_where, z_depth, z_pres, None)
    fx.update({"scene/logits_rule": np.array([10, 18]), "scene/z_where": z_where.numpy(), "scene/recon": recon.detach().numpy()})

    # (b),(c) random scenes with gradients, default config (C=1, 11x11, G=28) and RGB/8x8-grid/G=14
    for nm, (overrides, _shape) in RENDER_RANDOM.items():
        ns = rh.load_reference(overrides)
        net = rh.build_reference_model(ns)
        lg, zw, zd, zp, target = render_random_inputs(nm)
        leaves = [t.clone().requires_grad_(True) for t in (lg, zw, zd, zp)]
        recon = _render_with_logits(ns, net, *leaves, None)
        loss = torch.nn.functional.binary_cross_entropy(recon, target, reduction="sum")   # models.py:547
        loss.backward()
        dl = leaves[0].grad.flatten()
        idx = grad_sample_indices(nm + "/d_logits", dl.numel(), 16384)
        fx.update({nm + "/recon": recon.detach().numpy(), nm + "/bce": loss.detach().numpy(),
                   nm + "/d_logits_idx": idx.astype(np.int64), nm + "/d_logits_val": dl[torch.from_numpy(idx)].numpy(),
                   nm + "/d_logits_stat": np.array([dl.double().sum().item(), dl.double().norm().item()]),
                   nm + "/d_z_where": leaves[1].grad.numpy(), nm + "/d_z_depth": leaves[2].grad.numpy(),
                   nm + "/d_z_pres": leaves[3].grad.numpy()})
    path = os.path.join(HERE, "render.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    which = sys.argv[1:] or ["stn", "render", "model"]
    if "stn" in which:
        stn_cases()
    if "render" in which:
        render_cases()
    if "model" in which:
        model_cases()
