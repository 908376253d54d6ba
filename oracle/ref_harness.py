"""Loader for the UNMODIFIED reference (yonkshi/SPAIR_pytorch) — test infrastructure only.

This file is part of ``oracle/``: it is never imported by the product package
(``spair_pytorch_b200``).  It exists so that, inside the build container where
``/root/reference`` is mounted, the restatement in ``oracle/spair_oracle.py`` can be
pinned against the real reference and golden vectors can be generated
(``tests/golden/make_golden.py``).  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there may read that path; what does travel is ``baseline/_ref/`` (a byte-for-byte copy made by
``baseline/install_reference.py``, git-ignored), which ``bench.py``'s CPU legs and ``tests/test_train_py.py`` load
through this module.

What it does:
  * injects stub modules for ``tensorboardX``, ``matplotlib[.pyplot|.gridspec|.patches|
    .collections]`` and ``cycler`` (imported at reference ``models.py:11`` and
    ``debug_tools.py:1-5``; not installed in this image);
  * mutates ``spair.config`` BEFORE ``spair.modules`` is imported, because the reference
    captures config values as default arguments at import time (``modules.py:13,126``)
    and its ``Backbone._build_backbone`` mutates the topology dicts in place
    (``modules.py:53-55``), so a fresh deep copy is installed for every load;
  * replaces ``debug_tools.plot_prerender_components`` (hard-codes an 11x11 grid and
    ``cfg.BATCH_SIZE``, ``debug_tools.py:12,50``) by a no-op;
  * temporarily shadows this repo's own ``spair`` drop-in package in ``sys.modules`` and
    restores it afterwards.
"""
from __future__ import annotations

import contextlib
import copy
import importlib
import io
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# where the unmodified reference lives: the read-only mount of the build container, else the byte-for-byte copy that
# baseline/install_reference.py places in the git-ignored baseline/_ref/ (which gpurun ships to the GPU box)
_CANDIDATES = [os.environ.get("SPAIR_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
REFERENCE_ROOT = next((c for c in _CANDIDATES if c and os.path.isfile(os.path.join(c, "spair", "models.py"))), "/root/reference")

# stride-2/2/2 topology used by BASELINE.json configs 3 and 4 (cfg C/D)
TOPOLOGY_CELL8 = [
    dict(filters=128, kernel_size=4, stride=2),
    dict(filters=128, kernel_size=4, stride=2),
    dict(filters=128, kernel_size=4, stride=2),
    dict(filters=128, kernel_size=1, stride=1),
    dict(filters=128, kernel_size=1, stride=1),
    dict(filters=128, kernel_size=1, stride=1),
]


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "spair", "models.py"))


class NullWriter:
    """Stands in for tensorboardX.SummaryWriter (reference ``models.py:16``)."""

    def __getattr__(self, name):
        if name.startswith("add_"):
            return lambda *a, **k: None
        raise AttributeError(name)


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _install_stubs() -> None:
    if "tensorboardX" not in sys.modules:
        sys.modules["tensorboardX"] = _stub("tensorboardX", SummaryWriter=NullWriter)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        for sub in ("pyplot", "gridspec", "patches", "collections"):
            sm = _stub("matplotlib." + sub)
            setattr(mpl, sub, sm)
            sys.modules["matplotlib." + sub] = sm
        sys.modules["matplotlib.collections"].PatchCollection = object
        sys.modules["matplotlib"] = mpl
    if "cycler" not in sys.modules:
        sys.modules["cycler"] = _stub("cycler", cycler=lambda *a, **k: None)


def load_reference(overrides: dict | None = None) -> types.SimpleNamespace:
    """Import a fresh copy of the reference's ``spair`` package with ``overrides`` applied to
    ``spair.config``.  Returns a namespace with ``cfg``, ``modules``, ``models``,
    ``debug_tools`` (the reference's own module objects)."""
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    _install_stubs()
    saved = {k: v for k, v in sys.modules.items() if k == "spair" or k.startswith("spair.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        cfg = importlib.import_module("spair.config")
        cfg.DEFAULT_BACKBONE_TOPOLOGY = copy.deepcopy(cfg.DEFAULT_BACKBONE_TOPOLOGY)
        for k, v in (overrides or {}).items():
            setattr(cfg, k, copy.deepcopy(v))
        cfg.N_CONTEXT_DIM = 4 + cfg.N_ATTRIBUTES + 1 + 1
        debug_tools = importlib.import_module("spair.debug_tools")
        debug_tools.plot_prerender_components = lambda *a, **k: None
        modules = importlib.import_module("spair.modules")
        models = importlib.import_module("spair.models")
        metric = importlib.import_module("spair.metric")
        ns = types.SimpleNamespace(cfg=cfg, modules=modules, models=models, debug_tools=debug_tools, metric=metric)
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "spair" or k.startswith("spair.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ns


def build_reference_model(ns, seed: int = 3):
    """``SPAIR(image_shape, writer, device)`` exactly as ``train.py:39-41`` does (seed 3, CPU)."""
    import torch

    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        net = ns.models.SPAIR(ns.cfg.INPUT_IMAGE_SHAPE, NullWriter(), torch.device("cpu"))
    return net


def run_reference(net, x, global_step: int, noise_seed: int, backward: bool = True):
    """One ``forward`` (+ ``loss.backward(retain_graph=True)`` as ``train.py:65-66``) with the
    global torch RNG seeded right before the call, so the noise stream is the one
    ``oracle.spair_oracle.draw_noise`` replays."""
    import torch

    for p in net.parameters():
        p.grad = None
    torch.manual_seed(noise_seed)
    with contextlib.redirect_stdout(io.StringIO()):
        loss, recon, z_where, z_pres = net(x, global_step)
        if backward:
            loss.backward(retain_graph=True)
    return loss, recon, z_where, z_pres
