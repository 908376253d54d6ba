"""GPU check of the TMA-im2col convolution (spair_conv_gemm3x) against the explicit patch matrix + gemm3x and float64.
`timeout 300 python tools/conv_check.py`"""
import os
import sys

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


def check(B, H, W, C, Cout, k, s, time=False):
    g = torch.Generator(device="cuda").manual_seed(B + H + C)
    x = torch.randn(B, H, W, C, device="cuda", generator=g)
    w = torch.randn(Cout, k * k * C, device="cuda", generator=g) * 0.05
    b = torch.randn(Cout, device="cuda", generator=g)
    Ho, Wo = (H - k) // s + 1, (W - k) // s + 1
    M = B * Ho * Wo
    col = torch.empty(M, k * k * C, device="cuda")
    K.im2col_nhwc(x, k, s, col)
    ref = (col.double() @ w.double().t() + b.double()).clamp_min(0)
    y = torch.full((M, Cout), float("nan"), device="cuda")
    K.conv_fwd(x, k, s, w, b, y, relu=True)
    torch.cuda.synchronize()
    e_f = float((y.double() - ref).abs().max() / ref.abs().max())
    dy = torch.randn(M, Cout, device="cuda", generator=g)
    dwr = (dy.double().t() @ col.double())
    dw = torch.full((Cout, k * k * C), float("nan"), device="cuda")
    K.conv_wgrad(x, k, s, dy, dw)
    torch.cuda.synchronize()
    e_w = float((dw.double() - dwr).abs().max() / dwr.abs().max())
    line = "B=%d H=%d W=%d C=%d Cout=%d k=%d s=%d  M=%d: fwd err %.2e nan %d | wgrad err %.2e nan %d" % (
        B, H, W, C, Cout, k, s, M, e_f, int(torch.isnan(y).sum()), e_w, int(torch.isnan(dw).sum()))
    if time:
        y2 = torch.empty_like(y)
        t_f = timeit(lambda: K.conv_fwd(x, k, s, w, b, y, relu=True))
        t_fe = timeit(lambda: (K.im2col_nhwc(x, k, s, col), K.gemm3x(col, True, w, True, y2, b, epilogue=K.GEMM_EPI_RELU)))
        t_w = timeit(lambda: K.conv_wgrad(x, k, s, dy, dw))
        t_we = timeit(lambda: K.gemm3x(dy, False, col, False, dw))
        line += " | fwd %.3f ms (explicit im2col + gemm %.3f) wgrad %.3f ms (explicit gemm %.3f)" % (t_f, t_fe, t_w, t_we)
    print(line, flush=True)
    return max(e_f, e_w)


if __name__ == "__main__":
    worst = 0.0
    worst = max(worst, check(1, 10, 10, 32, 64, 4, 2))
    worst = max(worst, check(2, 9, 11, 64, 128, 3, 1))
    worst = max(worst, check(3, 50, 50, 128, 128, 4, 2))
    worst = max(worst, check(2, 24, 24, 128, 128, 4, 2))
    if "--quick" not in sys.argv:
        worst = max(worst, check(256, 50, 50, 128, 128, 4, 2, time=True))
        worst = max(worst, check(256, 24, 24, 128, 128, 4, 2, time=True))
    print("worst %.3e" % worst)
    sys.exit(0 if worst < 1e-5 else 1)
