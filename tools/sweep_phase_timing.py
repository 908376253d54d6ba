#!/usr/bin/env python
"""clock64 phase profile of the two persistent sweep kernels (CTA 0), the source of the phase table in
profiles/r01_sweep_kernels_full.md.  Needs a library built with the instrumentation compiled in:

    SPAIR_NVCC_EXTRA=-DSW_TIMING python spair_pytorch_b200/_build.py --force
    python tools/sweep_phase_timing.py
    python spair_pytorch_b200/_build.py --force          # back to the product build

The instrumented build is slower (thread 0 of CTA 0 does atomic-free global adds at every phase boundary); never
benchmark it.
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers  # noqa: E402
from spair_pytorch_b200 import kernels as K  # noqa: E402

NAMES = {0: 'fwd gather+tables', 1: 'fwd box mlp', 2: 'fwd box head', 3: 'fwd glimpse', 4: 'fwd enc mlp', 5: 'fwd attr head',
         6: 'fwd z mlp', 7: 'fwd depth head', 8: 'fwd obj mlp', 10: 'fwd layer0 (all mlps)', 11: 'fwd layer1', 12: 'fwd layer2',
         20: 'fwd accumulate', 21: 'fwd partial store+sync', 22: 'fwd finalize+sync',
         30: 'bwd ctx+pres head', 31: 'bwd obj mlp', 32: 'bwd depth head', 33: 'bwd z mlp', 34: 'bwd attr head', 35: 'bwd enc mlp',
         36: 'bwd box head', 37: 'bwd box mlp', 38: 'bwd glimpse grad', 40: 'bwd accumulate', 41: 'bwd partial+sync',
         42: 'bwd finalize+sync',
         13: 'fwd layer1->2 boundary (tensor-core path: 11 = layer 0, 13 = layer 1, 12 = layer 2)',
         43: 'bwd tc stage dY', 44: 'bwd tc layer 2', 45: 'bwd tc layer 1', 46: 'bwd tc layer 0'}
# tensor-core path (default): 20 / 40 = waiting for the accumulator, 21 / 41 = the rest of the epilogue


def main():
    lib = K.lib()
    if not hasattr(lib, "spair_debug_sweep_timing"):
        raise SystemExit("library was built without -DSW_TIMING (see the module docstring)")
    name = sys.argv[1] if len(sys.argv) > 1 else 'A'          # config (A / C / D) and batch: `... sweep_phase_timing.py C 64`
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    net = helpers.build_model(name).cuda()
    x = torch.rand(batch, *net.image_shape, device='cuda')
    if os.environ.get("TC_DEBUG_FLAGS"):     # instrumented builds: 1 = no weight copies, 2 = no MMAs (timing only, garbage results)
        lib.spair_debug_sweep_tc_flags(int(os.environ["TC_DEBUG_FLAGS"]))
    buf = (ctypes.c_longlong * 64)()
    for _ in range(3):
        net(x, 1000)[0].backward()
        lib.spair_debug_sweep_timing(buf)      # reads and clears: the last iteration's numbers remain
    v = list(buf)
    for k in sorted(NAMES):
        print('%-28s %10.1f kcyc' % (NAMES[k], v[k] / 1e3))
    print('fwd total %.1f kcyc, bwd total %.1f kcyc' % (sum(v[0:9]) / 1e3, sum(v[30:39]) / 1e3))


if __name__ == "__main__":
    main()
