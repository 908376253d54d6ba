"""Model-level parity on the B200: SPAIR.forward + loss.backward through the C-ABI kernels against
(a) golden vectors produced by the unmodified reference and (b) the CPU oracle run on the same
seeded inputs; plus size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest
import torch

from tests import helpers
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"
# measured budgets (profiles/r02_parity_escapes.json); see OTHER_SHAPES
# config A, B=32 (BASELINE configs[0]): 5 of the 200,704 elements of object_encoder.dense0.weight.grad (a K = 3,872-row
# reduction on the 3xTF32 tensor-core GEMM) sit just outside (rtol, atol) of the fp32 reference and inside it of the float64
# evaluation; no ref_err / kink escapes anywhere.  Budget: 1e-4 of a tensor's elements, float64-clause only in practice.
ESCAPE_BUDGET_A32 = 1e-4
ESCAPE_BUDGET_A2500 = 0.5
# config D, B=1: the weight gradients now come from the 3xTF32 tensor-core GEMM (csrc/gemm.cu), whose rounding differs from
# the fp32 reference's by ~1e-6 of sum|a||b|.  Measured: 2 of the 2,266,987 checked gradient elements (both in
# box_network.body.dense0.weight, 32,400 elements) sit just outside (rtol, atol) of the fp32 reference and inside it of the
# float64 evaluation; no ref_err / kink escapes.  Budget: 1e-4 of a tensor's elements.
ESCAPE_BUDGET_D1 = 1e-4


@pytest.mark.parametrize("fused", [(True, True), (True, False), (False, False)],
                         ids=["fused-sweep", "fused-forward-only", "per-wavefront"])
@pytest.mark.parametrize("name,step", [("tiny", 1), ("tiny", 1001), ("A", 1), ("A", 1001)])
def test_model_matches_reference_golden(name, step, fused):
    """The three launch modes of the cell sweep (one persistent kernel per direction / per-wavefront kernels + cuBLAS)
    against golden vectors produced by the unmodified reference."""
    net = helpers.build_model(name, DEV)
    g = helpers.load_golden("model_%s_step%d.npz" % (name, step))
    plan = net._get_plan(torch.device(DEV, torch.cuda.current_device()))
    plan.fused_forward, plan.fused_backward = fused
    report = helpers.check_model_against_golden(net, g, DEV)
    case = "golden %s step %d sweep=%s" % (name, step, "fused" if fused == (True, True) else "fwd-only" if fused[0] else "per-wavefront")
    helpers.record_escapes(case, report)
    # the plain fp32 criterion of the north_star: NO element of NO gradient tensor may need an escape clause
    helpers.assert_no_escapes(report, case)


def _run_oracle(net, x, step, noise, name):
    from oracle import spair_oracle as so
    cfg = helpers.oracle_config(name)
    params = so.params_from_state_dict(net.state_dict())
    out = so.forward_backward(params, x, step, noise, cfg, check_finite=False)
    return out, params


# (config, batch, step, fraction of a gradient tensor's elements that may need an escape clause).  0.0 = the plain fp32
# criterion.  ("A", 32, 1001) is BASELINE.json configs[0] as written: spair/config.py defaults, batch 32 (reference
# train.py:48-66 with cfg.BATCH_SIZE = 32).  Where the budget is non-zero the reference's own fp32 backward is rounding noise
# on those tensors and the counts are written to profiles/r02_parity_escapes.json.
OTHER_SHAPES = [("C", 2, 1001, 0.0), ("rgb64", 2, 1001, 0.0), ("D", 1, 1001, ESCAPE_BUDGET_D1), ("tiny_lb2", 3, 1001, 0.0),
                ("A", 32, 1001, ESCAPE_BUDGET_A32), ("A", 3, 2500, ESCAPE_BUDGET_A2500)]


@pytest.mark.parametrize("name,B,step,budget", OTHER_SHAPES, ids=["%s-%d-%d" % c[:3] for c in OTHER_SHAPES])
def test_model_vs_oracle_other_shapes(name, B, step, budget):
    """16x16 cells / 14x14 glimpses (config C), RGB with 8x8 cells, a later training step, a lookback-2 context, one image of
    BASELINE config 4 (256x256 RGB, 32x32 = 1024 cells, 28x28 glimpses; the oracle materialises 7.8 GB for it) and
    configs[0] itself (defaults, batch 32), checked against the oracle executed on the host in the same test."""
    from oracle import spair_oracle as so
    net = helpers.build_model(name, DEV)
    cfg = helpers.oracle_config(name)
    x = so.scattered_sprites(B, cfg.image_shape, seed=11, sprite_px=(8, 20))
    noise = so.random_noise(torch.Generator().manual_seed(5), B, cfg.grid, cfg.n_attr)
    want, params = _run_oracle(net, x, step, noise, name)
    _, params64 = so.forward_backward_fp64(net.state_dict(), x, step, noise, cfg)
    net.set_noise(noise.eps_where, noise.eps_attr, noise.eps_depth, noise.u_pres)
    loss, recon, z_where, z_pres = net(x.to(DEV), step)
    loss.backward(retain_graph=True)
    assert_close(loss, want["loss"], "loss")
    assert_close(recon, want["recon_x"], "recon_x")
    assert_close(z_where, want["z_where"], "z_where")
    assert_close(z_pres, want["z_pres"], "z_pres")
    assert_close(net.latent_maps()["z_attr"], want["z_attr"], "z_attr")
    for n, m in net.kl_maps().items():
        assert_close(m, want["kl"][n], "KL " + n)
    failures, report = [], {}
    for k, p in net.named_parameters():
        if k.startswith("attn."):
            assert p.grad is None
            continue
        ref = params[k].grad
        scale = float(ref.norm()) / np.sqrt(ref.numel())
        try:
            report[k] = helpers.grad_escape_report(p.grad, ref, params64[k].grad, "grad " + k, atol=1e-5 + 1e-4 * scale)
        except AssertionError as e:
            failures.append(str(e))
    assert not failures, "\n".join(failures)
    case = "oracle %s B=%d step %d" % (name, B, step)
    helpers.record_escapes(case, report)
    helpers.assert_no_escapes(report, case, max_fraction=budget, min_cos=(1.0 - 1e-6) if budget <= 1e-4 else 0.999)


def test_full_size_config_B_properties():
    """BASELINE config 2 (defaults, batch 256) at full size: finite loss/gradients, bitwise
    run-to-run determinism, and the data-parallel identity — two half batches with the KL term
    scaled by 1/2 give the gradients of the whole batch."""
    from oracle import spair_oracle as so
    net = helpers.build_model("A", DEV)
    torch.backends.cudnn.deterministic = True      # the cuDNN backbone is only reproducible when asked to be
    B = 256
    x = so.scattered_sprites(B, (1, 128, 128), seed=2).to(DEV)
    g = torch.Generator().manual_seed(1)
    noise = so.random_noise(g, B, (11, 11), 50)

    def run(sl, kl_scale):
        net.kl_scale = kl_scale
        net.set_noise(noise.eps_where[sl], noise.eps_attr[sl], noise.eps_depth[sl], noise.u_pres[sl])
        for p in net.parameters():
            p.grad = None
        loss = net(x[sl], 1500)[0]
        loss.backward()
        return loss.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}

    full = slice(0, B)
    loss1, g1 = run(full, 1.0)
    loss2, g2 = run(full, 1.0)
    assert torch.isfinite(loss1) and torch.equal(loss1, loss2)
    nondet = [k for k in g1 if not torch.equal(g1[k], g2[k])]
    assert not nondet, "non-deterministic gradients: %s" % nondet
    for k in g1:
        assert torch.isfinite(g1[k]).all(), k
    la, ga = run(slice(0, B // 2), 0.5)
    lb, gb = run(slice(B // 2, B), 0.5)
    net.kl_scale = 1.0
    torch.backends.cudnn.deterministic = False
    assert_close(la + lb, loss1, "sum of shard losses")
    failures = []
    for k in g1:
        scale = float(g1[k].norm()) / np.sqrt(g1[k].numel())
        try:
            assert_close(ga[k] + gb[k], g1[k], "sharded grad " + k, atol=1e-5 + 2e-4 * scale)
        except AssertionError as e:
            failures.append(str(e))
    assert not failures, "\n".join(failures)


def test_reference_training_loop_runs():
    """The reference's train.py step (zero_grad / forward / backward(retain_graph=True) / Adam step,
    train.py:64-67) on procedural scenes: loss stays finite and decreases over a few steps."""
    from oracle import spair_oracle as so
    net = helpers.build_model("A", DEV)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    x = so.scattered_sprites(32, (1, 128, 128), seed=4).to(DEV)
    losses = []
    noise = so.random_noise(torch.Generator().manual_seed(0), 32, (11, 11), 50)
    for it in range(8):
        opt.zero_grad()
        net.set_noise(noise.eps_where, noise.eps_attr, noise.eps_depth, noise.u_pres)   # fixed draws: deterministic objective
        loss, out_img, z_where, z_pres = net(x, 1000 + it)
        loss.backward(retain_graph=True)
        opt.step()
        losses.append(float(loss))
        assert out_img.shape == x.shape and z_where.shape == (32, 4, 11, 11) and z_pres.shape == (32, 1, 11, 11)
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_graphed_step_equals_eager_step():
    """The CUDA-graph replay of zero-grad + forward + backward gives bitwise the gradients of the eager
    step on the same batch and noise, and can be replayed on new batches / steps."""
    from oracle import spair_oracle as so
    from spair_pytorch_b200 import dp
    from spair_pytorch_b200.graphed import GraphedTrainStep
    net = helpers.build_model("tiny", DEV)
    torch.backends.cudnn.deterministic = True
    B = 8
    xs = [so.scattered_sprites(B, (1, 40, 40), seed=s, sprite_px=(6, 14)).to(DEV) for s in (1, 2)]
    noise = so.random_noise(torch.Generator().manual_seed(3), B, (5, 5), 50)
    dev_noise = [t.to(DEV) for t in (noise.eps_where, noise.eps_attr, noise.eps_depth, noise.u_pres)]
    bucket = dp.GradientBucket(dp.trainable_parameters(net))

    def eager(x, step):
        bucket.zero()
        net.set_noise(*dev_noise)
        loss = net(x, step)[0]
        loss.backward()
        return loss.detach().clone(), bucket.flat.clone()

    want = [eager(xs[0], 1200), eager(xs[1], 3000)]
    net.set_noise(*dev_noise, keep=True)            # device-resident noise becomes a static input of the graph
    gstep = GraphedTrainStep(net, xs[0], bucket=bucket, global_step=1200)
    for (x, step), (loss_w, grad_w) in zip(((xs[0], 1200), (xs[1], 3000)), want):
        loss = gstep(x, step)[0]
        torch.cuda.synchronize()
        assert torch.equal(loss, loss_w)
        assert torch.equal(bucket.flat, grad_w)
    # input double buffering: a pinned host batch staged on the copy stream, consumed by the next replay
    gstep.prefetch(xs[0].cpu().pin_memory())
    loss = gstep(None, 1200)[0]
    gstep.prefetch(xs[1].cpu().pin_memory())        # staged while the step above may still be running
    torch.cuda.synchronize()
    assert torch.equal(loss, want[0][0]) and torch.equal(bucket.flat, want[0][1])
    loss = gstep(None, 3000)[0]
    torch.cuda.synchronize()
    assert torch.equal(loss, want[1][0]) and torch.equal(bucket.flat, want[1][1])
    with pytest.raises(RuntimeError):
        gstep(None, 1200)                            # nothing staged
    torch.backends.cudnn.deterministic = False


@pytest.mark.parametrize("name,B", [("tiny", 5), ("A", 7), ("C", 3), ("rgb64", 2)])
def test_fused_forward_sweep_equals_per_wavefront_path(name, B):
    """The one-launch persistent forward sweep (csrc/sweep.cu) against the per-wavefront path (context/head/glimpse
    kernels + cuBLAS MLPs) on the same inputs and noise: same latents and same activation buffers up to the fp32
    summation order of the MLPs.  Gradients go through the same backward code; they are compared loosely here (a
    ReLU pre-activation within an ulp of 0 may take the other branch and shift a gradient by ~1e-3 of its scale) —
    their accuracy is pinned by the golden / oracle tests, which run both paths."""
    from oracle import spair_oracle as so
    net = helpers.build_model(name, DEV)
    cfg = helpers.oracle_config(name)
    x = so.scattered_sprites(B, cfg.image_shape, seed=21, sprite_px=(6, 14)).to(DEV)
    noise = so.random_noise(torch.Generator().manual_seed(4), B, cfg.grid, cfg.n_attr)
    results = []
    for fused in (True, False):
        net._plan = None
        net(x[:1], 1001)                    # builds the plan
        net._plan.fused_forward = net._plan.fused_backward = fused
        net.set_noise(noise.eps_where, noise.eps_attr, noise.eps_depth, noise.u_pres)
        for p in net.parameters():
            p.grad = None
        loss, recon, z_where, z_pres = net(x, 1001)
        loss.backward()
        L = net._latents
        results.append(dict(loss=loss.detach().clone(), recon=recon.detach().clone(), z_where=z_where.detach().clone(),
                            z_pres=z_pres.detach().clone(), attr=L.attr.detach().clone(), depth=L.depth.detach().clone(),
                            dmean=L.dmean.detach().clone(), dstd=L.dstd.detach().clone(),
                            grads={k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}))
    a, b = results
    for k in ("loss", "recon", "z_where", "z_pres", "attr", "depth", "dmean", "dstd"):
        assert_close(a[k], b[k], "fused vs per-wavefront " + k)
    for k in a["grads"]:      # norm-wise: some of these tensors are dominated by fp32 cancellation noise
        rel = float((a["grads"][k] - b["grads"][k]).norm() / b["grads"][k].norm().clamp(min=1e-12))
        assert rel <= 5e-2, "grad %s differs by %.3e (relative L2) between the two forward paths" % (k, rel)


@pytest.mark.parametrize("name,B", [("tiny", 5), ("tiny", 149), ("A", 4), ("A", 3), ("C", 3), ("rgb64", 2), ("tiny_lb2", 3), ("D", 1)])
def test_tensor_core_sweep_equals_simt_sweep(name, B, monkeypatch):
    """The fused sweeps with their dense layers on tcgen05 (csrc/sweep_tc.cuh: split-precision TF32 MMAs, weights streamed
    by bulk async copies; opt-in with SPAIR_SWEEP_TC=1) against the same kernels with fp32 SIMT dense layers (the default) on
    the same inputs and noise: every activation / gradient buffer of the four per-cell MLPs (reference modules.py:124-165 inside
    the loop of models.py:68-117) agrees to the split-precision rounding forward (~1e-6 of the buffer's scale).  The
    backward buffers see the same arithmetic but ALSO the forward's 1e-6 differences amplified by the renderer's loss
    (BCE gradients ~ 1 / recon): a few 1e-3 of their scale on these inputs, which is why the
    tensor-core FORWARD sweep is opt-in and not the parity path.  B = 149 (> 148 SMs) runs two images per CTA with a single
    image in the last CTA, the other cases one image per CTA; D has more feature tiles (3 x 28 x 28 inputs) than one
    accumulator group."""
    from oracle import spair_oracle as so
    net = helpers.build_model(name, DEV)
    cfg = helpers.oracle_config(name)
    x = so.scattered_sprites(B, cfg.image_shape, seed=33, sprite_px=(6, 14)).to(DEV)
    noise = so.random_noise(torch.Generator().manual_seed(6), B, cfg.grid, cfg.n_attr)
    results = []
    for tc in ("0", "1"):
        monkeypatch.setenv("SPAIR_SWEEP_TC", tc)
        net._plan = None
        net(x[:1], 1001)
        net.set_noise(noise.eps_where, noise.eps_attr, noise.eps_depth, noise.u_pres)
        for p in net.parameters():
            p.grad = None
        loss = net(x, 1001)[0]
        loss.backward()
        bufs = {}
        for mname, m in zip(("box", "enc", "z", "obj"), net._plan.last_mlps):
            for bname, t in (("X", m.X), ("H0", m.H[0]), ("H1", m.H[1]), ("Y", m.Y), ("dX", m.dX), ("dH0", m.dH[0]),
                             ("dH1", m.dH[1]), ("dY", m.dY)):
                bufs[mname + "." + bname] = t.detach().clone()
        results.append((loss.detach().clone(), bufs, {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}))
    (l0, b0, g0), (l1, b1, g1) = results
    assert abs(float(l0 - l1)) <= 1e-5 * abs(float(l0))
    for k in b0:
        scale = float(b0[k].abs().max())
        tol = (2e-5 if k.split(".")[1].startswith("d") is False else 1e-2) * max(scale, 1e-12)
        err = float((b0[k] - b1[k]).abs().max())
        assert err <= tol, "%s: tensor-core vs SIMT sweep differ by %.3e (scale %.3e)" % (k, err, scale)
    for k in g0:
        rel = float((g0[k] - g1[k]).norm() / g0[k].norm().clamp(min=1e-12))
        assert rel <= 1e-2, "grad %s differs by %.3e (relative L2) between the tensor-core and the SIMT sweep" % (k, rel)
