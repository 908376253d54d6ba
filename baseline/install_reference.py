#!/usr/bin/env python
"""Installs the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` so that it travels to the GPU box.

    python baseline/install_reference.py            # build container only: needs /root/reference

The reference (yonkshi/SPAIR_pytorch) is a pure-Python package without ``setup.py`` / ``pyproject.toml``, so there is
nothing for ``pip install`` to build; its "installation" is its source tree on ``sys.path``.  This script copies
``spair/*.py`` and ``train.py`` byte-for-byte (no edits; SHA-256 of every file is written to ``baseline/_ref/MANIFEST.json``)
from ``/root/reference``.  ``baseline/_ref/`` is listed in ``.gitignore`` (reference sources never enter this repo's
history) and NOT in ``.gpurunignore`` (so ``gpurun`` ships it).  Consumers: ``bench.py --impl reference`` and
``cpu_baseline`` (the reference's own CPU path timed on the GPU box's host cores), ``tests/test_train_py.py``
(reference ``train.py`` executed unchanged against the drop-in ``spair`` package).  They import it through
``oracle/ref_harness.py``, which supplies the stub modules the reference's imports need (tensorboardX, matplotlib,
cycler) — the product package never touches it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

SRC = os.environ.get("SPAIR_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def install(verbose: bool = True) -> str | None:
    if not os.path.isfile(os.path.join(SRC, "spair", "models.py")):
        if verbose:
            print("reference not mounted at %s: nothing installed" % SRC)
        return None
    os.makedirs(os.path.join(DST, "spair"), exist_ok=True)
    manifest = {}
    files = [os.path.join("spair", f) for f in sorted(os.listdir(os.path.join(SRC, "spair"))) if f.endswith(".py")]
    files.append("train.py")
    for rel in files:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("installed %d reference files into %s" % (len(files), DST))
    return DST


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
