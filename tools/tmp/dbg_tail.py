import sys, os
sys.path.insert(0, os.getcwd())
import torch
from spair_pytorch_b200 import ops, kernels as K
from spair_pytorch_b200.modules import Backbone
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(1)
net = Backbone([1, 128, 128], 100).cuda()
x = torch.rand(3, 1, 128, 128, device="cuda")
# stage-by-stage: run the tail layer by layer with torch (cudnn) capturing activations and grads
y0 = net._fused_stem(x).detach().requires_grad_(True)
acts = [y0]
cur = y0
layers = list(net.net)[2:]
for l in layers:
    cur = l(cur)
    cur.retain_grad()
    acts.append(cur)
g = torch.Generator(device="cuda").manual_seed(2)
wgt = torch.randn(cur.shape, device="cuda", generator=g)
(cur * wgt).sum().backward()
ref_dy0 = y0.grad.clone()
ref_grads = [p.grad.clone() for p in list(net.parameters())[2:]]
for p in net.parameters(): p.grad = None
y0b = y0.detach().clone().requires_grad_(True)
ops.USE_TENSOR_CORE_GEMM = True
feat = net._gemm_tail(y0b)
print("feat rel", float((feat - cur).norm() / cur.norm()))
(feat * wgt).sum().backward()
print("dy0 rel", float((y0b.grad - ref_dy0).norm() / ref_dy0.norm()))
for (n, p), r in zip(list(net.named_parameters())[2:], ref_grads):
    print(n, float((p.grad - r).norm() / r.norm()))
# direct check of dgrad GEMM shapes
for (M, N, Kd) in [(363, 128, 128), (363, 2048, 128), (1728, 2048, 128), (363, 128, 100)]:
    A = torch.randn(M, Kd, device="cuda"); B = torch.randn(Kd, N, device="cuda") * 0.1
    out = torch.empty(M, N, device="cuda")
    K.gemm3x(A, True, B, False, out)
    ref = A.double() @ B.double()
    print("dgrad gemm", M, N, Kd, float((out.double() - ref).norm() / ref.norm()))
print("---- forward intermediates")
xh = y0.detach().permute(0, 2, 3, 1).contiguous()
convs = [l for l in layers if isinstance(l, torch.nn.Conv2d)]
ai = 0
for ci, conv in enumerate(convs):
    k, s = conv.kernel_size[0], conv.stride[0]
    relu = ci < len(convs) - 1
    Bn, H, W, Cin = xh.shape
    Ho, Wo = (H - k) // s + 1, (W - k) // s + 1
    M = Bn * Ho * Wo
    wr = conv.weight.detach().permute(0, 2, 3, 1).reshape(conv.out_channels, -1).contiguous()
    if k == 1:
        a = xh.view(M, Cin)
    else:
        a = torch.empty(M, k * k * Cin, device="cuda"); K.im2col_nhwc(xh, k, s, a)
    y = torch.empty(M, conv.out_channels, device="cuda")
    K.gemm3x(a, True, wr, True, y, conv.bias.detach(), epilogue=1 if relu else 0)
    ref = acts[2 * ci + 2 if relu else 2 * ci + 1].detach().permute(0, 2, 3, 1).reshape(M, -1)
    d = (y - ref).abs()
    mism = int(((y > 0) != (ref > 0)).sum())
    pre64 = a.double() @ wr.double().t() + conv.bias.detach().double()
    print("layer", ci, "M", M, "max abs diff %.3e" % float(d.max()), "rel %.3e" % float((y - ref).norm() / ref.norm()),
          "mask mismatches", mism, "of", y.numel(), "| ours vs fp64 max %.3e, cudnn vs fp64 max %.3e" % (
              float((y.double() - (pre64.clamp_min(0) if relu else pre64)).abs().max()),
              float((ref.double() - (pre64.clamp_min(0) if relu else pre64)).abs().max())),
          "near-zero(<1e-6) count", int((pre64.abs() < 1e-6).sum()))
    xh = y.view(Bn, Ho, Wo, -1)
