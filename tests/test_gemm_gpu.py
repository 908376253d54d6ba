"""The tcgen05 3xTF32 GEMM (csrc/gemm.cu, spair_gemm3x) and the decoder path built on it, on the B200.

Checker: torch fp64 on the same inputs (the reference's `nn.Linear` / autograd arithmetic, modules.py:124-165,
models.py:474-493).  Bar: rtol 1e-4 / atol 1e-5 in fp32 (north_star); the split-precision product is measured against
the fp64 result relative to sum_k |a||b| — fp32 cuBLAS sits at ~3e-7 on that scale, 3xTF32 must stay below 5e-6.
"""
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
