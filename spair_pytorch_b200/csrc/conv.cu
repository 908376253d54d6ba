// Patch gather / scatter around the tcgen05 GEMM (csrc/gemm.cu) for the strided convolutions of the backbone tail
// (reference modules.py:44-66: Conv2d(128,128,4,stride 2) x2 after the stem, then 1x1 convolutions), the caller side of
// the per-cell path.  Activations are kept channels-last ([B,H,W,C]) between the layers, so
//   * a 1x1 convolution IS a GEMM on the stored tensor ([B*H*W, Cin] x [Cout, Cin]^T), no copy;
//   * a k x k / stride s convolution is a GEMM on the patch matrix col[m][(kh*k + kw)*C + c] = x[b][s*oy+kh][s*ox+kw][c],
//     m = (b*Ho + oy)*Wo + ox, built here with 128-bit loads and stores (every tap is C contiguous floats);
//   * its input gradient is the transposed gather of the GEMM's d_col (fixed summation order, no atomics).
// Both kernels are pure HBM streams (one read + one write of the patch matrix).
#include "common.cuh"

namespace spair {

__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const float4* __restrict__ x, float4* __restrict__ col, int B, int H, int W,
                                                          int C4, int k, int s, int Ho, int Wo) {
    const long long M = (long long)B * Ho * Wo;
    const int row4 = k * k * C4;                      // float4s per patch row
    for (long long m = blockIdx.x; m < M; m += gridDim.x) {
        const int ox = (int)(m % Wo);
        const long long t = m / Wo;
        const int oy = (int)(t % Ho), b = (int)(t / Ho);
        const float4* src = x + (((long long)b * H + (long long)s * oy) * W + (long long)s * ox) * C4;
        float4* dst = col + m * row4;
        for (int i = threadIdx.x; i < row4; i += blockDim.x) {
            const int tap = i / C4, c4 = i - tap * C4;
            const int kh = tap / k, kw = tap - kh * k;
            dst[i] = __ldg(src + ((long long)kh * W + kw) * C4 + c4);
        }
    }
}

// dx[b][y][x][:] = sum over the taps (kh, kw) with (y - kh) % s == 0, (x - kw) % s == 0 and the output pixel inside the grid
__global__ void __launch_bounds__(256) col2im_nhwc_kernel(const float4* __restrict__ dcol, float4* __restrict__ dx, int B, int H,
                                                          int W, int C4, int k, int s, int Ho, int Wo) {
    const long long P = (long long)B * H * W;
    const int row4 = k * k * C4;
    const int pix_per_cta = blockDim.x / C4;           // host guarantees blockDim.x % C4 == 0
    const int sub = threadIdx.x / C4, c4 = threadIdx.x - sub * C4;
    for (long long p0 = (long long)blockIdx.x * pix_per_cta; p0 < P; p0 += (long long)gridDim.x * pix_per_cta) {
        const long long p = p0 + sub;
        if (p >= P) continue;
        const int xx = (int)(p % W);
        const long long t = p / W;
        const int yy = (int)(t % H), b = (int)(t / H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kh = yy % s; kh < k; kh += s) {
            const int oy = (yy - kh) / s;
            if (yy < kh || oy >= Ho) continue;
            for (int kw = xx % s; kw < k; kw += s) {
                const int ox = (xx - kw) / s;
                if (xx < kw || ox >= Wo) continue;
                const float4 v = __ldg(dcol + (((long long)b * Ho + oy) * Wo + ox) * row4 + (kh * k + kw) * C4 + c4);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        dx[p * C4 + c4] = acc;
    }
}

}  // namespace spair

using namespace spair;

extern "C" int spair_im2col_nhwc(const float* x, int B, int H, int W, int C, int k, int stride, float* col, void* stream) {
    SPAIR_REQUIRE(x && col && B > 0 && H >= k && W >= k && C > 0 && (C & 3) == 0 && k > 0 && stride > 0);
    SPAIR_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)col & 15) == 0);
    const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
    const long long M = (long long)B * Ho * Wo;
    const int grid = (int)(M < (long long)kSMs * 16 ? M : (long long)kSMs * 16);
    im2col_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(col), B, H,
                                                               W, C / 4, k, stride, Ho, Wo);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_col2im_nhwc(const float* dcol, int B, int H, int W, int C, int k, int stride, float* dx, void* stream) {
    SPAIR_REQUIRE(dcol && dx && B > 0 && H >= k && W >= k && C > 0 && (C & 3) == 0 && C <= 1024 && k > 0 && stride > 0);
    SPAIR_REQUIRE(((uintptr_t)dx & 15) == 0 && ((uintptr_t)dcol & 15) == 0);
    const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
    const int C4 = C / 4;
    const int threads = (256 / C4) * C4 > 0 ? (256 / C4) * C4 : C4;
    const long long P = (long long)B * H * W;
    const long long ctas = (P + threads / C4 - 1) / (threads / C4);
    const int grid = (int)(ctas < (long long)kSMs * 32 ? ctas : (long long)kSMs * 32);
    col2im_nhwc_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(dcol), reinterpret_cast<float4*>(dx), B,
                                                                   H, W, C4, k, stride, Ho, Wo);
    SPAIR_LAUNCH_CHECK();
}
