"""The north_star's literal boundary claim: the reference's train.py runs UNCHANGED against the drop-in ``spair``
package (reference train.py:33-101).  The script comes from /root/reference (build container) or from the byte-for-byte
copy in baseline/_ref (GPU box); tests/train_py_runner.py supplies the modules this image lacks."""
import numpy as np
import pytest
import torch

from tests import train_py_runner as runner

needs_train_py = pytest.mark.skipif(runner.train_py_path() is None, reason="reference train.py not available")


def _check(writer, log, n_iter):
    assert writer.images == n_iter
    for it in range(n_iter):
        assert "Iteration %d" % it in log                   # train.py:63
    totals = writer.scalars["losses/total"]                 # written by SPAIR._build_loss through the writer train.py passes in
    assert [s for s, _ in totals] == list(range(n_iter))
    assert all(np.isfinite(v) for _, v in totals)
    assert [s for s, _ in writer.scalars["training_wheel"]] == list(range(n_iter))
    C, H, W2 = writer.last_image.shape                      # torch.cat([image_in, image_out], dim=2), train.py:70-73
    assert W2 == 2 * H
    return [v for _, v in totals]


@needs_train_py
def test_reference_train_py_runs_unchanged_on_the_host_logic(monkeypatch):
    """CPU-only container: same script, kernels replaced by the test double (tests/cpu_kernel_mock.py), batch 2."""
    from tests import cpu_kernel_mock
    cpu_kernel_mock.install(monkeypatch)
    writer, log = runner.run(n_iterations=2, gpu=False, batch_size=2, n_scenes=8)
    _check(writer, log, 2)


@pytest.mark.gpu
@needs_train_py
def test_reference_train_py_runs_unchanged_on_the_gpu():
    """`python train.py --gpu` for 4 iterations at cfg.BATCH_SIZE = 32 (reference defaults) on the sm_100a kernels."""
    writer, log = runner.run(n_iterations=4, gpu=True)
    _check(writer, log, 4)
