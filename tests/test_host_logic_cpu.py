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
    helpers.check_model_against_golden(net, g, "cpu")


def test_model_orchestration_matches_reference_golden_A(monkeypatch):
    cpu_kernel_mock.install(monkeypatch)
    net = helpers.build_model("A")
    g = helpers.load_golden("model_A_step1001.npz")
    helpers.check_model_against_golden(net, g, "cpu")


def test_product_path_refuses_cpu_tensors():
    from spair_pytorch_b200 import kernels as K
    net = helpers.build_model("tiny")
    with pytest.raises(K.SpairKernelError):
        net(torch.zeros(2, 1, 40, 40), 1)
