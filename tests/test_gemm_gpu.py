"""The tcgen05 3xTF32 GEMM (csrc/gemm.cu, spair_gemm3x) and the decoder path built on it, on the B200.

Checker: torch fp64 on the same inputs (the reference's `nn.Linear` / autograd arithmetic, modules.py:124-165,
models.py:474-493).  Bar: rtol 1e-4 / atol 1e-5 in fp32 (north_star); the split-precision product is measured against
the fp64 result relative to sum_k |a||b| — fp32 cuBLAS sits at ~3e-7 on that scale, 3xTF32 must stay below 5e-6.
"""
import copy

import pytest
import torch

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def K():
    from spair_pytorch_b200 import kernels
    return kernels


def _operands(M, N, Kd, a_k, b_k, seed, lda_pad=0, ldb_pad=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    if lda_pad:   # row pitch > width, kept a multiple of 4 floats (TMA)
        lda_pad += (-((Kd if a_k else M) + lda_pad)) % 4
    if ldb_pad:
        ldb_pad += (-((Kd if b_k else N) + ldb_pad)) % 4
    A = torch.randn((M, Kd + lda_pad) if a_k else (Kd, M + lda_pad), device=DEV, generator=g)
    B = torch.randn((N, Kd + ldb_pad) if b_k else (Kd, N + ldb_pad), device=DEV, generator=g) * 0.25
    A = A[:, :Kd] if a_k else A[:, :M]
    B = B[:, :Kd] if b_k else B[:, :N]
    return A, B, (A if a_k else A.t()), (B.t() if b_k else B)


@pytest.mark.parametrize("a_k,b_k", [(1, 1), (1, 0), (0, 1), (0, 0)])
@pytest.mark.parametrize("M,N,Kd", [(128, 64, 32), (1, 8, 4), (200, 72, 100), (700, 520, 260), (3872, 1568, 256), (257, 224, 36)])
def test_gemm3x_vs_fp64(a_k, b_k, M, N, Kd):
    """Every operand-major combination (y = x W^T, dx = dy W, dW = dy^T x run on the stored tensors), ragged M / N / K
    tails, strided rows (row pitch > width), one-row and one-k-block edge cases."""
    A, B, Am, Bm = _operands(M, N, Kd, a_k, b_k, seed=M + N + Kd, lda_pad=4, ldb_pad=8)
    bias = torch.randn(N, device=DEV)
    out = torch.full((M, N), float("nan"), device=DEV)
    K().gemm3x(A, a_k, B, b_k, out, bias)
    ref = Am.double() @ Bm.double() + bias.double()
    scale = Am.double().abs() @ Bm.double().abs() + bias.double().abs()
    err = ((out.double() - ref).abs() / scale).max().item()
    assert not torch.isnan(out).any()
    assert err < 5e-6, err
    assert_close(out, ref.float(), "gemm3x", atol=1e-5 + 1e-4 * float(ref.abs().max()) * 0.01)


@pytest.mark.parametrize("M,N,Kd", [(100, 324, 30976), (256, 784, 3872), (1568, 256, 7744)])
def test_gemm3x_split_k_weight_gradient_shapes(M, N, Kd):
    """dW = dy^T x with the reduction over all B*HW rows: split over CTAs into the workspace, summed in a fixed order
    (bitwise run-to-run reproducible), equal to the single-split result up to fp32 summation order."""
    A, B, Am, Bm = _operands(M, N, Kd, 0, 0, seed=Kd)
    k = K()
    splits = k.lib().spair_gemm_splits(M, N, Kd)
    assert splits > 1
    out1, out2 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    k.gemm3x(A, 0, B, 0, out1)
    k.gemm3x(A, 0, B, 0, out2)
    assert torch.equal(out1, out2)
    ref = Am.double() @ Bm.double()
    scale = Am.double().abs() @ Bm.double().abs()
    assert ((out1.double() - ref).abs() / scale).max().item() < 5e-6


def test_gemm3x_relu_and_texel_epilogues():
    """Epilogues of the decoder layers: bias + ReLU (build_MLP hidden layers, modules.py:141-144) and the texel decode
    of SPAIR._render (models.py:485-493) with the alpha channel stored as the complement 1 - sigma."""
    M, N, Kd, period = 500, 1568, 256, 2
    A, B, Am, Bm = _operands(M, N, Kd, 1, 1, seed=5)
    bias = torch.randn(N, device=DEV)
    pre = Am.double() @ Bm.double() + bias.double()
    out = torch.empty(M, N, device=DEV)
    K().gemm3x(A, 1, B, 1, out, bias, epilogue=K().GEMM_EPI_RELU)
    assert_close(out, pre.clamp_min(0).float(), "relu epilogue", atol=1e-5 + 1e-6 * float(pre.abs().max()))
    K().gemm3x(A, 1, B, 1, out, bias, epilogue=K().GEMM_EPI_TEXEL, period=period, scales=(2.0, 0.1, 5.0))
    colour = 1 / (torch.exp(-2.0 * pre[:, 0::2]) + 1)
    alpha_c = 1 - 1 / (torch.exp(-(0.1 * pre[:, 1::2] + 5.0)) + 1)
    assert_close(out[:, 0::2], colour.float(), "texel colour")
    # the complement carries sigma'(x) = s (1 - s): it must be RELATIVELY accurate where alpha saturates
    assert_close(out[:, 1::2], alpha_c.float(), "texel alpha complement", atol=0.0, rtol=1e-4)


def test_gemm3x_rejects_unaligned_rows():
    """TMA needs 16-byte row pitches: the entry point refuses (SPAIR_ERR_INVALID) instead of mis-loading; ops.py keeps
    such layers (K = 50 attribute columns, the 479-column obj_network input) on cuBLAS."""
    k = K()
    A = torch.randn(64, 50, device=DEV)
    B = torch.randn(32, 50, device=DEV)
    assert not k.gemm_supported(A, B)
    with pytest.raises(k.SpairKernelError):
        k.gemm3x(A, 1, B, 1, torch.empty(64, 32, device=DEV))


def test_decoder_function_matches_cublas_path():
    """ops.DecoderFunction + RenderFunction(decoded=True) against the round-1 path (nn.Sequential decoder on cuBLAS fp32 +
    RenderFunction on raw logits) on the same inputs: canvas, loss and every gradient."""
    from spair_pytorch_b200 import ops
    torch.manual_seed(0)
    B, HW, C, G, I, A = 4, 121, 1, 28, 128, 50
    N = B * HW
    dec = torch.nn.Sequential(torch.nn.Linear(A, 128), torch.nn.ReLU(), torch.nn.Linear(128, 256), torch.nn.ReLU(),
                              torch.nn.Linear(256, G * G * (C + 1))).to(DEV)
    attr = torch.randn(N, A, device=DEV, requires_grad=True)
    zw = torch.rand(N, 4, device=DEV)
    zw[:, 2:] = (12 + 36 * zw[:, 2:]) / I
    zw.requires_grad_(True)
    zd = (4 * torch.rand(N, device=DEV)).requires_grad_(True)
    zp = torch.rand(N, device=DEV).requires_grad_(True)
    x = torch.rand(B, C, I, I, device=DEV)
    scales = (2.0, 0.1, 5.0)
    leaves = [attr, zw, zd, zp] + list(dec.parameters())

    def run(decoded):
        for t in leaves:
            t.grad = None
        if decoded:
            lin = [m for m in dec if isinstance(m, torch.nn.Linear)]
            tex = ops.DecoderFunction.apply(attr, *(p for m in lin for p in (m.weight, m.bias)), C + 1, scales)
        else:
            tex = dec(attr)
        recon, bce, _ = ops.RenderFunction.apply(tex, zw, zd, zp, x, B, HW, C, G, I, I, scales, decoded)
        (bce + (recon * recon).sum()).backward()
        return recon.detach().clone(), bce.detach().clone(), [t.grad.detach().clone() for t in leaves]

    recon0, bce0, g0 = run(False)
    recon1, bce1, g1 = run(True)
    assert_close(recon1, recon0, "canvas")
    assert_close(bce1, bce0, "bce", atol=1e-5 + 1e-4 * float(bce0.abs()))
    for i, (a, b) in enumerate(zip(g1, g0)):
        rel = float((a - b).double().norm() / (b.double().norm() + 1e-30))
        assert rel < 1e-4, (i, rel)
        assert_close(a, b, "grad %d" % i, atol=1e-5 + 1e-4 * float(b.abs().max()))


@pytest.mark.parametrize("shape,topology", [((1, 96, 96), None), ((3, 64, 64), "cell8")])
def test_conv_tail_matches_cudnn(shape, topology, monkeypatch):
    """Backbone tail (conv_1 .. conv_out, reference modules.py:44-66) on the tcgen05 GEMM + patch gather (ops.ConvTailFunction,
    csrc/conv.cu) against the cuDNN fp32 path of the same module: features, input-side gradient (through the stem) and
    every weight / bias gradient."""
    from spair_pytorch_b200 import config as cfg, ops
    from spair_pytorch_b200.modules import Backbone
    torch.backends.cudnn.allow_tf32 = False          # strict fp32 on the library path too (the model sets this itself)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    topo = cfg.CELL8_BACKBONE_TOPOLOGY if topology == "cell8" else None
    net = Backbone(list(shape), 100, topology=topo).to(DEV)
    # (small batch: a pre-activation within ~1e-6 of zero takes the other ReLU branch under another rounding and moves
    # every upstream gradient by ~1/sqrt(#activations); with few activations none sits that close — checked below)
    x = torch.rand(2, *shape, device=DEV)

    def run(tc):
        monkeypatch.setattr(ops, "USE_TENSOR_CORE_GEMM", tc)
        for p in net.parameters():
            p.grad = None
        feat = net(x)
        g = torch.Generator(device=DEV).manual_seed(2)
        (feat * torch.randn(feat.shape, device=DEV, generator=g)).sum().backward()
        return feat.detach().clone(), [p.grad.detach().clone() for p in net.parameters()]

    f0, g0 = run(False)
    f1, g1 = run(True)
    # float64 evaluation of the same module on the CPU: the judge of both fp32 paths (cuDNN picks FFT / SIMT algorithms whose
    # own rounding is of the same size as the difference between the two paths)
    net64 = copy.deepcopy(net).double().cpu()
    for p in net64.parameters():
        p.grad = None
    feat64 = net64(x.double().cpu())
    g = torch.Generator(device=DEV).manual_seed(2)
    (feat64 * torch.randn(f0.shape, device=DEV, generator=g).double().cpu()).sum().backward()
    g64 = [p.grad for p in net64.parameters()]

    def err(a, ref):
        return float((a.double().cpu() - ref).norm() / (ref.norm() + 1e-30))

    assert err(f1, feat64) <= max(2e-6, 2.0 * err(f0, feat64)), (err(f1, feat64), err(f0, feat64))
    assert_close(f1, feat64.float(), "features", atol=1e-5 + 1e-4 * float(feat64.abs().max()) * 0.05)
    for (name, _), a, b, r in zip(net.named_parameters(), g1, g0, g64):
        # (elementwise closeness to float64 is not a criterion here: an activation within an ulp of zero takes the other
        # ReLU branch in fp32 and moves single gradient entries by percents — in the cuDNN path just as much)
        assert err(a, r) <= max(5e-5, 2.0 * err(b, r)), (name, err(a, r), err(b, r))
        print("%-28s rel. error vs float64: tcgen05 %.2e, cuDNN %.2e" % (name, err(a, r), err(b, r)))


def test_im2col_col2im_adjoint():
    """<im2col(x), c> == <x, col2im(c)> (the two kernels of csrc/conv.cu are adjoint), and im2col equals torch's unfold."""
    k_ = K()
    g = torch.Generator(device=DEV).manual_seed(3)
    for (B, H, W, C, k, s) in [(2, 50, 50, 128, 4, 2), (3, 9, 11, 8, 3, 1), (1, 7, 7, 4, 1, 1)]:
        x = torch.randn(B, H, W, C, device=DEV, generator=g)
        Ho, Wo = (H - k) // s + 1, (W - k) // s + 1
        col = torch.empty(B * Ho * Wo, k * k * C, device=DEV)
        k_.im2col_nhwc(x, k, s, col)
        ref = torch.nn.functional.unfold(x.permute(0, 3, 1, 2), k, stride=s)            # [B, C*k*k, L], (c, kh, kw) order
        ref = ref.view(B, C, k * k, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, k * k * C)
        assert torch.equal(col, ref)
        c = torch.randn(col.shape, device=DEV, generator=g)
        dx = torch.empty_like(x)
        k_.col2im_nhwc(c, k, s, dx)
        lhs, rhs = float((col.double() * c.double()).sum()), float((x.double() * dx.double()).sum())
        assert abs(lhs - rhs) <= 1e-6 * (abs(lhs) + 1.0)


def test_transpose_and_relu_bwd_colsum():
    """Layout / reduction helpers of csrc/conv.cu: batched transpose (NCHW <-> channels-last) is an exact permutation;
    the fused ReLU-mask + bias-gradient pass equals torch's mask and column sum (bitwise mask, sum to fp32 rounding)
    and is run-to-run reproducible."""
    k_ = K()
    g = torch.Generator(device=DEV).manual_seed(11)
    for (B, R, C) in [(3, 2500, 128), (2, 121, 100), (1, 33, 7)]:
        x = torch.randn(B, R, C, device=DEV, generator=g)
        out = torch.empty(B, C, R, device=DEV)
        k_.transpose_batched(x, out)
        assert torch.equal(out, x.transpose(1, 2).contiguous())
    for (rows, cols, masked) in [(30976, 1568, False), (3000, 128, True), (7, 4, True), (1000, 102, False), (513, 100, True)]:
        gr = torch.randn(rows, cols, device=DEV, generator=g)
        y = torch.randn(rows, cols, device=DEV, generator=g).clamp_min(0) if masked else None
        want_g = gr * (y > 0) if masked else gr.clone()
        want = want_g.double().sum(0)
        a, b = gr.clone(), gr.clone()
        o1, o2 = torch.empty(cols, device=DEV), torch.empty(cols, device=DEV)
        k_.relu_bwd_colsum(a, y, o1)
        k_.relu_bwd_colsum(b, y, o2)
        assert torch.equal(a, want_g) and torch.equal(o1, o2)
        assert_close(o1, want.float(), "column sums", atol=1e-5 + 1e-6 * float(want_g.abs().sum(0).max()))


@pytest.mark.parametrize("B,H,W,C,Cout,k,s", [(1, 10, 10, 32, 64, 4, 2), (2, 9, 11, 64, 128, 3, 1), (3, 50, 50, 128, 128, 4, 2),
                                              (5, 24, 24, 128, 100, 4, 2)])
def test_implicit_conv_matches_explicit_patches(B, H, W, C, Cout, k, s):
    """spair_conv_gemm3x (TMA im2col loads inside the GEMM's producer warp) against the materialised patch matrix in float64:
    forward with bias + ReLU (rows = output pixels, tiles spanning image rows and images, ragged last tile) and the weight
    gradient (reduction over the pixels, split-K, partial column tiles)."""
    k_ = K()
    g = torch.Generator(device=DEV).manual_seed(B * 31 + H)
    x = torch.randn(B, H, W, C, device=DEV, generator=g)
    w = torch.randn(Cout, k * k * C, device=DEV, generator=g) * 0.05
    b = torch.randn(Cout, device=DEV, generator=g)
    Ho, Wo = (H - k) // s + 1, (W - k) // s + 1
    M = B * Ho * Wo
    col = torch.empty(M, k * k * C, device=DEV)
    k_.im2col_nhwc(x, k, s, col)
    pre = col.double() @ w.double().t() + b.double()
    y = torch.full((M, Cout), float("nan"), device=DEV)
    k_.conv_fwd(x, k, s, w, b, y, relu=True)
    scale = col.double().abs() @ w.double().abs().t() + b.double().abs()
    assert ((y.double() - pre.clamp_min(0)).abs() / scale).max().item() < 5e-6
    assert torch.equal(y > 0, pre > 0)                 # the exact-sign fix-up: every ReLU branch is the float64 one
    dy = torch.randn(M, Cout, device=DEV, generator=g)
    dw = torch.full((Cout, k * k * C), float("nan"), device=DEV)
    k_.conv_wgrad(x, k, s, dy, dw)
    ref = dy.double().t() @ col.double()
    scale = dy.double().abs().t() @ col.double().abs()
    assert ((dw.double() - ref).abs() / scale).max().item() < 5e-6


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,s", [(1, 10, 10, 32, 64, 4, 2), (3, 50, 50, 128, 128, 4, 2), (2, 24, 24, 128, 128, 4, 2),
                                                (2, 13, 11, 16, 32, 4, 2), (2, 12, 12, 8, 32, 3, 3)])
def test_implicit_conv_dgrad_matches_torch(B, H, W, Cin, Cout, k, s):
    """spair_conv_dgrad3x (the input gradient as s*s implicit GEMMs over dy, rows scattered to their pixels by the epilogue)
    against torch's conv2d backward in float64, incl. odd sizes whose last rows / columns no window reaches (zero gradient)."""
    k_ = K()
    g = torch.Generator(device=DEV).manual_seed(H * 7 + Cin)
    w = torch.randn(Cout, Cin, k, k, device=DEV, generator=g) * 0.05
    Ho, Wo = (H - k) // s + 1, (W - k) // s + 1
    dy = torch.randn(B, Ho, Wo, Cout, device=DEV, generator=g)
    x64 = torch.zeros(B, Cin, H, W, device=DEV, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv2d(x64, w.double(), stride=s).backward(dy.permute(0, 3, 1, 2).double())
    ref = x64.grad.permute(0, 2, 3, 1)
    dx = torch.full((B, H, W, Cin), float("nan"), device=DEV)
    k_.conv_dgrad(dy.view(-1, Cout), B, k, s, k_.pack_dgrad_weights(w, s), dx)
    assert not torch.isnan(dx).any()
    scale = torch.nn.functional.conv_transpose2d(dy.permute(0, 3, 1, 2).double().abs(), w.double().abs(), stride=s)
    scale = torch.nn.functional.pad(scale, (0, W - scale.shape[3], 0, H - scale.shape[2])).permute(0, 2, 3, 1) + 1e-30
    assert ((dx.double() - ref).abs() / scale.clamp_min(1e-6)).max().item() < 5e-6


def test_presplit_weights_give_identical_results():
    """A weight operand split into its TF32 hi / lo planes ONCE (spair_split_tf32 / kernels.SplitWeight) instead of tile by
    tile inside the GEMM: bitwise the same results — plain / ReLU (incl. the exact-sign pass, which still reads the fp32
    weights) / texel epilogues with the weight K-major (y = x W^T), the same planes MN-major (dx = dy W), the implicit-GEMM
    convolution and its input gradient; ragged N, K tails and a padded class-weight tensor."""
    k_ = K()
    g = torch.Generator(device=DEV).manual_seed(77)
    for M, N, Kd in ((700, 520, 260), (3872, 1568, 256), (257, 100, 36)):
        x = torch.randn(M, Kd, device=DEV, generator=g)
        w = torch.randn(N, Kd, device=DEV, generator=g) * 0.2
        b = torch.randn(N, device=DEV, generator=g)
        ws = k_.SplitWeight(w)
        assert torch.equal(ws.hi + ws.lo, w) or float((ws.hi + ws.lo - w).abs().max()) <= 2.0 ** -23 * float(w.abs().max())
        for kwargs in (dict(), dict(bias=b, epilogue=k_.GEMM_EPI_RELU), dict(bias=b, epilogue=k_.GEMM_EPI_TEXEL, period=2, scales=(2.0, 0.1, 5.0))):
            y0, y1 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
            k_.gemm3x(x, True, w, True, y0, **kwargs)
            k_.gemm3x(x, True, ws, True, y1, **kwargs)
            assert torch.equal(y0, y1), (M, N, Kd, kwargs.get("epilogue"))
        dy = torch.randn(M, N, device=DEV, generator=g)
        d0, d1 = torch.empty(M, Kd, device=DEV), torch.empty(M, Kd, device=DEV)
        k_.gemm3x(dy, True, w, False, d0)
        k_.gemm3x(dy, True, ws, False, d1)
        assert torch.equal(d0, d1)
    for B, H, W, C, Cout, k, s in ((3, 50, 50, 128, 128, 4, 2), (5, 24, 24, 128, 100, 4, 2)):
        x = torch.randn(B, H, W, C, device=DEV, generator=g)
        w = torch.randn(Cout, k * k * C, device=DEV, generator=g) * 0.05
        b = torch.randn(Cout, device=DEV, generator=g)
        Ho, Wo = (H - k) // s + 1, (W - k) // s + 1
        y0, y1 = torch.empty(B * Ho * Wo, Cout, device=DEV), torch.empty(B * Ho * Wo, Cout, device=DEV)
        k_.conv_fwd(x, k, s, w, b, y0, relu=True)
        k_.conv_fwd(x, k, s, k_.SplitWeight(w), b, y1, relu=True)
        assert torch.equal(y0, y1)
        if Cout % 32 == 0:
            w4 = w.view(Cout, k, k, C).permute(0, 3, 1, 2)
            wc = k_.pack_dgrad_weights(w4, s)
            dyc = torch.randn(B * Ho * Wo, Cout, device=DEV, generator=g)
            dx0, dx1 = torch.empty(B, H, W, C, device=DEV), torch.empty(B, H, W, C, device=DEV)
            k_.conv_dgrad(dyc, B, k, s, wc, dx0)
            k_.conv_dgrad(dyc, B, k, s, k_.SplitWeight(wc), dx1)
            assert torch.equal(dx0, dx1)
