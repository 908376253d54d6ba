"""Diagnostic: fused sweep kernels vs the per-wavefront path on one golden case — max |diff| of every saved activation /
gradient buffer and every parameter gradient.  python tools/sweep_diag.py [config] [B] [step]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import spair_oracle as so  # noqa: E402
from tests import helpers  # noqa: E402


def run(net, x, noise, step, fused):
    net._plan = None
    net(x[:1], step)
    net._plan.fused_forward, net._plan.fused_backward = fused
    net.set_noise(noise.eps_where, noise.eps_attr, noise.eps_depth, noise.u_pres)
    for p in net.parameters():
        p.grad = None
    loss = net(x, step)[0]
    loss.backward()
    mlps = net._plan.last_mlps
    bufs = {}
    for name, m in zip(("box", "enc", "z", "obj"), mlps):
        bufs[name + ".X"] = m.X.clone(); bufs[name + ".H0"] = m.H[0].clone(); bufs[name + ".H1"] = m.H[1].clone()
        bufs[name + ".Y"] = m.Y.clone()
        bufs[name + ".dX"] = m.dX.clone(); bufs[name + ".dH0"] = m.dH[0].clone(); bufs[name + ".dH1"] = m.dH[1].clone()
        bufs[name + ".dY"] = m.dY.clone()
    grads = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    return float(loss), bufs, grads


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    step = int(sys.argv[3]) if len(sys.argv) > 3 else 1001
    dev = "cuda"
    torch.backends.cudnn.deterministic = True
    net = helpers.build_model(name, dev)
    cfg = helpers.oracle_config(name)
    if os.environ.get("DIAG_GOLDEN"):
        g = helpers.load_golden("model_%s_step%d.npz" % (name, step))
        x = torch.from_numpy(g["x"]).to(dev)
        noise = so.Noise(*(torch.from_numpy(g[k]) for k in ("eps_where", "eps_attr", "eps_depth", "u_pres")))
        B = x.shape[0]
    else:
        x = so.scattered_sprites(B, cfg.image_shape, seed=int(os.environ.get("DIAG_SEED", "11")), sprite_px=(8, 20)).to(dev)
        noise = so.random_noise(torch.Generator().manual_seed(5), B, cfg.grid, cfg.n_attr)
    ref = run(net, x, noise, step, (False, False))
    for label, fused in (("fused fwd + per-wavefront bwd", (True, False)), ("fused fwd + fused bwd", (True, True))):
        got = run(net, x, noise, step, fused)
        print("==", label, "loss", got[0], "ref", ref[0])
        for k in ref[1]:
            a, b = got[1][k], ref[1][k]
            d = (a - b).abs()
            i = int(d.argmax())
            r, c = divmod(i, a.shape[1])
            rowrel = (d.max(1).values / b.abs().max(1).values.clamp(min=1e-20))
            bad = [int(v) for v in torch.nonzero(rowrel > 1e-3).flatten()[:12]]
            print("  %-8s max|diff| %.3e (scale %.3e) at row %d col %d  got %.6g ref %.6g  rows>1e-3: %s" % (k, float(d.max()), float(b.abs().max()), r, c, float(a.flatten()[i]), float(b.flatten()[i]), bad))
        worst = sorted(((float((got[2][k] - ref[2][k]).norm() / ref[2][k].norm().clamp(min=1e-20)), k) for k in ref[2]), reverse=True)[:8]
        for rel, k in worst:
            print("  grad %-40s rel L2 diff %.3e" % (k, rel))


if __name__ == "__main__":
    main()
