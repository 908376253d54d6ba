"""spair_pytorch_b200 — B200-native implementation of SPAIR's per-cell object pipeline.

Layout:
  csrc/            hand-written sm_100a CUDA kernels + the C-ABI (include/spair_b200.h)
  kernels.py       ctypes binding of libspair_b200.so (no CPU implementation exists)
  ops.py           autograd wrappers + the wavefront sweep (host orchestration, cuBLAS for the MLPs)
  schedule.py      wavefront schedule of the autoregressive cell loop
  models.py, modules.py, config.py, debug_tools.py, metric.py, dataloader.py, logging.py
                   host-side mirror of the reference's ``spair`` package API
  dp.py            data-parallel training over images (one process per GPU, NCCL gradient allreduce)

``import spair`` (the top-level shim package of this repo) aliases these modules under the
reference's names so that the reference ``train.py`` runs unchanged.
"""
__version__ = "0.1.0"
