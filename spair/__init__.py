"""Drop-in ``spair`` package: the reference's module names, served by ``spair_pytorch_b200``.

``from spair.models import SPAIR``, ``from spair import config as cfg``,
``from spair.dataloader import SimpleScatteredMNISTDataset``, ``from spair import debug_tools, metric``
(reference train.py:12-16) resolve to the B200 implementation; ``spair.config`` IS
``spair_pytorch_b200.config`` (same module object), so ``cfg.X = ...`` reaches the model.
"""
import importlib
import sys

for _name in ("config", "logging", "debug_tools", "modules", "models", "metric", "dataloader"):
    _mod = importlib.import_module("spair_pytorch_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod
del _name, _mod
