import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference mounted at /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch
    has_gpu = torch.cuda.is_available()
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(skip_gpu)


class NullWriter:
    def add_scalar(self, *a, **k):
        pass

    add_image = add_figure = add_histogram = add_scalar


@pytest.fixture
def null_writer():
    return NullWriter()


def pytest_sessionfinish(session, exitstatus):
    """Dump how every gradient tensor was judged (tests/helpers.ESCAPE_LOG) for profiles/r02_parity_escapes.json."""
    import json
    from tests import helpers
    if helpers.ESCAPE_LOG:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        summary = {case: {"tensors": len(rep), "elements": sum(r["n"] for r in rep.values()),
                          "fp64": sum(r["fp64"] for r in rep.values()), "ref_err": sum(r["ref_err"] for r in rep.values()),
                          "kink": sum(r["kink"] for r in rep.values()), "min_cos": min(r["cos"] for r in rep.values()),
                          "escaped": {k: r for k, r in rep.items() if r["fp64"] + r["ref_err"] + r["kink"]}}
                   for case, rep in helpers.ESCAPE_LOG.items()}
        with open(os.path.join(out, "parity_escapes.json"), "w") as f:
            json.dump(summary, f, indent=1, sort_keys=True)
