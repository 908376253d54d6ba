"""Host-logic tests that run WITHOUT a GPU: the orchestration in ops.py / models.py (wavefront
schedule, buffer bookkeeping, hand-written backward sweep) is exercised with the kernel binding
replaced by a test double (tests/cpu_kernel_mock.py) and compared with golden vectors produced by
the unmodified reference.  The CUDA kernels themselves are tested in test_kernels_gpu.py."""
import numpy as np
import pytest
import torch

from tests import cpu_kernel_mock, helpers


@pytest.mark.parametrize("step", [1, 1001])
def test_model_orchestration_matches_reference_golden_tiny(monkeypatch, step):
    cpu_kernel_mock.install(monkeypatch)
    net = helpers.build_model("tiny")
    g = helpers.load_golden("model_tiny_step%d.npz" % step)
    report = helpers.check_model_against_golden(net, g, "cpu")
    helpers.assert_no_escapes(report, "host logic tiny step %d" % step)


def test_model_orchestration_matches_reference_golden_A(monkeypatch):
    cpu_kernel_mock.install(monkeypatch)
    net = helpers.build_model("A")
    g = helpers.load_golden("model_A_step1001.npz")
    report = helpers.check_model_against_golden(net, g, "cpu")
    helpers.assert_no_escapes(report, "host logic A step 1001")


def test_model_orchestration_lookback2_matches_oracle(monkeypatch):
    """N_LOOKBACK = 2 (12 context neighbours, wavefronts t = w + 3h; reference models.py:26,292-320): schedule, context
    gather / gradient gather and the hand-written backward against the oracle on the same inputs and noise."""
    from oracle import spair_oracle as so
    cpu_kernel_mock.install(monkeypatch)
    net = helpers.build_model("tiny_lb2")
    cfg = helpers.oracle_config("tiny_lb2")
    x = so.scattered_sprites(3, cfg.image_shape, seed=11, sprite_px=(6, 14))
    noise = so.random_noise(torch.Generator().manual_seed(5), 3, cfg.grid, cfg.n_attr)
    params = so.params_from_state_dict(net.state_dict())
    want = so.forward_backward(params, x, 1001, noise, cfg, check_finite=False)      # fills params[k].grad
    net.set_noise(noise.eps_where, noise.eps_attr, noise.eps_depth, noise.u_pres)
    loss, recon, z_where, z_pres = net(x, 1001)
    loss.backward()
    helpers.assert_close(loss, want["loss"], "loss")
    helpers.assert_close(recon, want["recon_x"], "recon_x")
    helpers.assert_close(z_where, want["z_where"], "z_where")
    helpers.assert_close(z_pres, want["z_pres"], "z_pres")
    for k, p in net.named_parameters():
        if not k.startswith("attn."):
            ref = params[k].grad
            helpers.assert_close(p.grad, ref, "grad " + k, atol=1e-5 + 1e-4 * float(ref.norm()) / np.sqrt(ref.numel()))


def test_product_path_refuses_cpu_tensors():
    from spair_pytorch_b200 import kernels as K
    net = helpers.build_model("tiny")
    with pytest.raises(K.SpairKernelError):
        net(torch.zeros(2, 1, 40, 40), 1)


def test_flat_parameter_adam_equals_per_tensor_adam():
    """dp.GradientBucket.flatten_parameters: parameters aliased into one flat buffer, Adam over the single flat Parameter
    (grad = the gradient bucket) performs exactly the per-tensor updates."""
    import copy
    from spair_pytorch_b200 import dp
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Conv2d(2, 3, 3), torch.nn.Flatten(), torch.nn.Linear(3 * 6 * 6, 5))
    net = copy.deepcopy(ref)
    bucket = dp.GradientBucket(list(net.named_parameters()))
    flat = bucket.flatten_parameters()
    assert flat.grad is bucket.flat and bucket.flatten_parameters() is flat
    for (_, p), q in zip(net.named_parameters(), ref.parameters()):
        assert torch.equal(p, q) and p.data_ptr() >= flat.data_ptr() and p.data_ptr() < flat.data_ptr() + flat.numel() * 4
    opt_flat = torch.optim.Adam([flat], lr=1e-2, foreach=False)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-2, foreach=False)
    for it in range(4):
        x = torch.randn(7, 2, 8, 8)
        bucket.zero()
        net(x).square().sum().backward()
        bucket.check_attached()
        opt_flat.step()
        opt_ref.zero_grad()
        ref(x).square().sum().backward()
        opt_ref.step()
        for p, q in zip(net.parameters(), ref.parameters()):
            assert torch.equal(p, q), "update %d differs" % it


def test_sweep_tensor_core_mode_selection(monkeypatch):
    """Which sweep runs its dense layers on the tensor cores (ops.sweep_tc_choice; SPAIR_SWEEP_TC): by default only the
    backward sweep, and only when a CTA has >= 12 rows per wavefront; the forward never unless asked for (it is not
    parity-exact on tcgen05)."""
    from spair_pytorch_b200 import ops
    monkeypatch.delenv("SPAIR_SWEEP_TC", raising=False)
    assert ops.sweep_tc_choice(16) == (False, True) and ops.sweep_tc_choice(12) == (False, True)
    assert ops.sweep_tc_choice(8) == (False, False) and ops.sweep_tc_choice(6) == (False, False)
    assert ops.sweep_tc_choice(6, "bwd") == (False, True)
    assert ops.sweep_tc_choice(16, "0") == (False, False)
    assert ops.sweep_tc_choice(4, "1") == (True, True)
    monkeypatch.setenv("SPAIR_SWEEP_TC", "1")
    assert ops.sweep_tc_choice(16) == (True, True)
    monkeypatch.setenv("SPAIR_SWEEP_TC", "yes")
    try:
        ops.sweep_tc_choice(16)
    except ValueError:
        pass
    else:
        raise AssertionError("an unknown SPAIR_SWEEP_TC value must be rejected")
