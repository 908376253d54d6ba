"""GPU check + timing of the tcgen05 3xTF32 GEMM (csrc/gemm.cu) against torch fp64 on the shapes the path uses.
Run on a B200: `timeout 300 python tools/gemm_check.py [--quick]`.  Prints max relative error (vs fp64, scaled by the
row/column norms) next to cuBLAS fp32's, and the time of both."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from spair_pytorch_b200 import kernels as K


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def check(name, M, N, Kd, a_k, b_k, epilogue=0, bias=True, splits=None, time=True):
    torch.manual_seed(M * 7 + N * 3 + Kd)
    dev = "cuda"
    A = torch.randn((M, Kd) if a_k else (Kd, M), device=dev)
    B = torch.randn((N, Kd) if b_k else (Kd, N), device=dev) * 0.1
    b = torch.randn(N, device=dev) if bias else None
    Am = A if a_k else A.t()
    Bm = B.t() if b_k else B
    ref64 = Am.double() @ Bm.double()
    if b is not None:
        ref64 = ref64 + b.double()
    if epilogue == 1:
        ref64 = ref64.clamp_min(0)
    elif epilogue == 2:
        scale = torch.full((N,), 2.0, device=dev, dtype=torch.float64)
        off = torch.zeros(N, device=dev, dtype=torch.float64)
        scale[1::2], off[1::2] = 0.1, 5.0
        ref64 = 1 / (torch.exp(-(ref64 * scale + off)) + 1)
        ref64[:, 1::2] = 1 - ref64[:, 1::2]      # the alpha channel is stored as the complement
    out = torch.full((M, N), float("nan"), device=dev)
    fn = lambda: K.gemm3x(A, a_k, B, b_k, out, b, epilogue=epilogue, period=2, scales=(2.0, 0.1, 5.0), splits=splits)
    fn()
    torch.cuda.synchronize()
    ref32 = Am @ Bm
    if b is not None:
        ref32 = ref32 + b
    denom = (Am.double().abs() @ Bm.double().abs()) + 1e-30 if epilogue != 2 else torch.ones_like(ref64)
    err = ((out.double() - ref64).abs() / denom).max().item()
    err32 = ((ref32.double() - ref64).abs() / denom).max().item() if epilogue == 0 else float("nan")
    line = "%-28s M=%6d N=%5d K=%6d a_k=%d b_k=%d epi=%d splits=%s  err %.2e (cuBLAS fp32 %.2e)  nan=%d" % (
        name, M, N, Kd, a_k, b_k, epilogue, splits, err, err32, int(torch.isnan(out).sum()))
    if time:
        t = timeit(fn)
        t_ref = timeit(lambda: torch.mm(Am, Bm))
        line += "  %.3f ms (%.1f TFLOP/s) vs cuBLAS %.3f ms" % (t, 2.0 * M * N * Kd / t / 1e9, t_ref)
    print(line, flush=True)
    return err


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    torch.backends.cuda.matmul.allow_tf32 = False
    worst = 0.0
    if "--decoder-only" in sys.argv:
        check("decoder out fwd (plain)", 30976, 1568, 256, 1, 1)
        check("decoder out fwd (texel)", 30976, 1568, 256, 1, 1, epilogue=2)
        check("conv_1 fwd", 147456, 128, 2048, 1, 1, epilogue=1)
        check("conv_1 dgrad", 147456, 2048, 128, 1, 0, bias=False)
        check("conv_1 wgrad", 128, 2048, 147456, 0, 0, bias=False)
        sys.exit(0)
    # small shapes first: one tile, one k-block, each major combination
    for a_k in (1, 0):
        for b_k in (1, 0):
            worst = max(worst, check("tiny", 128, 64, 32, a_k, b_k, bias=False, time=False))
            worst = max(worst, check("tails", 200, 72, 100, a_k, b_k, time=False))
            worst = max(worst, check("multi-tile", 700, 520, 260, a_k, b_k, time=False))
    if not quick:
        worst = max(worst, check("decoder out fwd", 30976, 1568, 256, 1, 1, epilogue=2))
        worst = max(worst, check("decoder out fwd (plain)", 30976, 1568, 256, 1, 1))
        worst = max(worst, check("decoder dense1 fwd", 30976, 256, 128, 1, 1, epilogue=1))
        worst = max(worst, check("decoder out dgrad", 30976, 256, 1568, 1, 0, bias=False))
        worst = max(worst, check("decoder out wgrad", 1568, 256, 30976, 0, 0, bias=False))
        worst = max(worst, check("encoder dense0 wgrad", 256, 784, 30976, 0, 0, bias=False))
        worst = max(worst, check("box dense0 wgrad", 100, 324, 30976, 0, 0, bias=False))
        worst = max(worst, check("wgrad 1 split", 256, 784, 30976, 0, 0, bias=False, splits=1))
    print("worst err %.3e" % worst)
    sys.exit(0 if worst < 2e-6 else 1)
