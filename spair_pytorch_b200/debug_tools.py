"""Debug helpers under the reference's ``spair.debug_tools`` names (debug_tools.py:13-271).

Only what the hot path and ``train.py`` touch is functional: ``nan_hunter``, ``benchmark_init`` /
``benchmark``.  The matplotlib figure helpers of the reference are visual debugging and out of
scope; they are importable and raise a clear error when called.
The drop-in model does NOT call ``nan_hunter`` per cell: each call is one device->host sync per
tensor (~1300 per forward in the reference).  Use ``check_finite`` once per step instead.
"""
import time

import torch

from . import config as cfg

GRID_SIZE = 11
BENCHMARK_INIT_TIME = time.time()


def benchmark_init():
    global BENCHMARK_INIT_TIME
    BENCHMARK_INIT_TIME = time.time()


def benchmark(name='', print_benchmark=True):
    """Wall-clock seconds since the previous call (debug_tools.py:30-40)."""
    global BENCHMARK_INIT_TIME
    now = time.time()
    diff, BENCHMARK_INIT_TIME = now - BENCHMARK_INIT_TIME, now
    if print_benchmark:
        print('{}: {:.4f} '.format(name, diff))
    return diff


def nan_hunter(hunter_name, **kwargs):
    """Raise AssertionError if any tensor argument holds a NaN, after dumping all arguments
    (debug_tools.py:245-271).  One host sync per tensor."""
    tensors = {k: v for k, v in kwargs.items() if isinstance(v, torch.Tensor)}
    if not any(bool(torch.isnan(v).any()) for v in tensors.values()):
        return
    print('======== NAN DETECTED in %s =======' % hunter_name)
    for k, v in kwargs.items():
        print(k, ':', v)
    print('======== END OF NAN DETECTED =======')
    raise AssertionError('NAN Detected by Nan detector')


def check_finite(**tensors):
    """Single fused non-finite check for a whole step: one reduction per tensor on the device, ONE
    host sync in total."""
    flags = torch.stack([(~torch.isfinite(t)).any() for t in tensors.values()])
    if bool(flags.any()):
        bad = [k for k, f in zip(tensors, flags.tolist()) if f]
        raise AssertionError('non-finite values in: %s' % ', '.join(bad))


def torch2npy(t: torch.Tensor, reshape=False):
    shape = t.shape[1:]
    if reshape:
        return t.cpu().view(cfg.BATCH_SIZE, GRID_SIZE, GRID_SIZE, *shape).detach().squeeze().numpy()
    return t.cpu().detach().numpy()


def _no_plot(*_a, **_k):
    raise NotImplementedError('matplotlib figure helpers of the reference are out of scope of the B200 hot-path build')


plot_torch_image_in_pyplot = plot_prerender_components = plot_cropped_input_images = _no_plot
plot_objet_attr_latent_representation = plot_stn_input_and_out = _no_plot
