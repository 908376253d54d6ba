"""Runs the reference's ``train.py`` UNCHANGED against this repo's drop-in ``spair`` package (test infrastructure).

``train.py`` (reference train.py:1-105) imports ``tensorboardX``, ``coolname`` and — through ``spair.dataloader`` —
``h5py``, none of which exist in this image, opens an HDF5 file the reference does not ship (train.py:38) and loops
over 100,000 epochs.  This runner supplies what is missing WITHOUT touching the script:

  * stub modules ``tensorboardX`` (a recording SummaryWriter), ``coolname`` (``generate_slug``) and ``h5py`` (``File``
    returns procedurally generated scattered-sprite scenes under ``train/full/{image,bbox,digit_count}``, the schema of
    reference dataloader.py:13-33);
  * ``sys.path``: the repo root first, so ``from spair.models import SPAIR`` etc. (train.py:12-16) resolve to the drop-in;
  * a stop: the stub writer's ``add_image`` (called once per iteration, train.py:73) raises ``StopTraining`` after
    ``n_iterations`` calls;
  * ``torch.utils.data.DataLoader`` is forced to ``num_workers=0`` (a forked worker would inherit a CUDA context).

The script itself is executed with ``runpy.run_path(..., run_name="__main__")`` from the unmodified file.
"""
from __future__ import annotations

import contextlib
import io
import os
import runpy
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class StopTraining(Exception):
    pass


class RecordingWriter:
    """Stands in for tensorboardX.SummaryWriter (train.py:21; passed into SPAIR at train.py:41)."""
    instances = []

    def __init__(self, *a, **k):
        self.scalars, self.images, self.limit = {}, 0, None
        RecordingWriter.instances.append(self)

    def add_scalar(self, tag, value, step=None, *a, **k):
        self.scalars.setdefault(tag, []).append((step, float(value.detach()) if torch.is_tensor(value) else float(value)))

    def add_image(self, tag, img, step=None, *a, **k):
        self.images += 1
        self.last_image = img.detach().cpu()
        if self.limit is not None and self.images >= self.limit:
            raise StopTraining()

    def add_figure(self, *a, **k):
        pass

    add_histogram = add_figure


def train_py_path():
    for cand in ("/root/reference/train.py", os.path.join(ROOT, "baseline", "_ref", "train.py")):
        if os.path.isfile(cand):
            return cand
    return None


def run(n_iterations: int, gpu: bool, batch_size: int | None = None, n_scenes: int = 64):
    """Returns (writer, namespace of the executed script)."""
    from spair_pytorch_b200 import config as cfg
    from spair_pytorch_b200.dataloader import scattered_sprites
    path = train_py_path()
    assert path, "reference train.py not available"
    imgs, boxes, counts = scattered_sprites(n_scenes, tuple(cfg.INPUT_IMAGE_SHAPE), seed=99, return_boxes=True)
    scenes = {"image": imgs[:, 0].numpy(), "bbox": boxes.numpy(), "digit_count": counts.numpy()}

    class _File(dict):
        def __init__(self, *a, **k):
            super().__init__({"train/full": scenes})

    stubs = {"tensorboardX": types.ModuleType("tensorboardX"), "coolname": types.ModuleType("coolname"),
             "h5py": types.ModuleType("h5py")}
    stubs["tensorboardX"].SummaryWriter = RecordingWriter
    stubs["coolname"].generate_slug = lambda n=2: "drop-in-test"
    stubs["h5py"].File = _File
    saved_mods = {k: sys.modules.get(k) for k in stubs}
    saved_argv, saved_path, saved_bs = sys.argv, list(sys.path), cfg.BATCH_SIZE
    real_loader = torch.utils.data.DataLoader

    def loader(*a, **k):
        k["num_workers"] = 0
        k["pin_memory"] = k.get("pin_memory", False) and torch.cuda.is_available()
        return real_loader(*a, **k)

    RecordingWriter.instances.clear()
    try:
        sys.modules.update(stubs)
        sys.path.insert(0, ROOT)
        sys.argv = ["train.py"] + (["--gpu"] if gpu else [])
        if batch_size is not None:
            cfg.BATCH_SIZE = batch_size
        torch.utils.data.DataLoader = loader
        orig_init = RecordingWriter.__init__      # arm the stop on the writer train.py creates at import time (train.py:21)

        def armed_init(self, *a, **k):
            orig_init(self, *a, **k)
            self.limit = n_iterations
        RecordingWriter.__init__ = armed_init
        ns = None
        try:
            with contextlib.redirect_stdout(io.StringIO()) as out:
                try:
                    ns = runpy.run_path(path, run_name="__main__")
                except StopTraining:
                    pass
        finally:
            RecordingWriter.__init__ = orig_init
        return RecordingWriter.instances[0], out.getvalue()
    finally:
        torch.utils.data.DataLoader = real_loader
        cfg.BATCH_SIZE = saved_bs
        sys.argv, sys.path[:] = saved_argv, saved_path
        for k, v in saved_mods.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
