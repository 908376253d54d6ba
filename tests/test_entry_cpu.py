"""The driver's entry points, without a GPU: __graft_entry__.build() must compile the sm_100a library, import the package
and leave a library whose ABI version matches the binding (round-1 regression: a stale `== 1` assert made it raise)."""
import ctypes

import __graft_entry__ as entry
from spair_pytorch_b200 import _build, kernels


def test_build_entry_point_returns_and_library_matches_binding(capsys):
    entry.build()
    assert "built" in capsys.readouterr().out
    handle = ctypes.CDLL(_build.LIB_PATH)
    handle.spair_abi_version.restype = ctypes.c_int
    assert handle.spair_abi_version() == kernels.ABI_VERSION
    for name in kernels.EXPORTED_SYMBOLS:
        assert hasattr(handle, name), name
