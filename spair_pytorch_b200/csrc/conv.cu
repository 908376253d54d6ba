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

// out[b][c][r] = in[b][r][c]: 32x32 tiles through shared memory, both sides coalesced (NCHW <-> NHWC of a feature map with
// R = H*W pixels and C channels per image).
__global__ void __launch_bounds__(256) transpose_batched_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
    __shared__ float tile[32][33];
    const size_t base = (size_t)blockIdx.z * R * C;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int i = ty; i < 32; i += 8)
        if (r0 + i < R && c0 + tx < C) tile[i][tx] = in[base + (size_t)(r0 + i) * C + c0 + tx];
    __syncthreads();
#pragma unroll
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < C && r0 + tx < R) out[base + (size_t)(c0 + i) * R + r0 + tx] = tile[tx][i];
}

// Column sums of a [rows, cols] matrix (the bias gradient of a dense / 1x1 layer), optionally fused with the ReLU mask of
// the layer's output: g[r][c] *= (y[r][c] > 0) in place, then summed.  Stage 1: CTA (column block of 128, row chunk) keeps
// per-thread partial sums of 4 adjacent columns (128-bit loads) and adds its 8 warps in shared memory; stage 2 adds the
// chunks in a fixed order.  Deterministic, one pass over g (and y).
// rows per CTA: enough chunks to put >= 4 CTAs on every SM whatever the column count (a [30976, 128] gradient has ONE
// column block), at least 32 rows each
static int colsum_rows_per_cta(int rows, int cols) {
    const int col_blocks = (cols + 127) / 128;
    int chunks = (4 * kSMs + col_blocks - 1) / col_blocks;
    if (chunks > (rows + 31) / 32) chunks = (rows + 31) / 32;
    if (chunks < 1) chunks = 1;
    return (rows + chunks - 1) / chunks;
}

__global__ void __launch_bounds__(256) colsum_stage1_kernel(float* __restrict__ g, int ld_g, const float* __restrict__ y, int ld_y,
                                                            int rows, int cols, int kCsRowsPerCta, float* __restrict__ partial) {
    __shared__ float4 red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 128 + lane * 4;
    const int r_begin = blockIdx.y * kCsRowsPerCta, r_end = min(rows, r_begin + kCsRowsPerCta);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < cols) {      // cols % 4 == 0 (host)
#pragma unroll 4
        for (int r = r_begin + warp; r < r_end; r += 8) {
            float4 v = *reinterpret_cast<const float4*>(g + (size_t)r * ld_g + c);
            if (y) {
                const float4 m = __ldg(reinterpret_cast<const float4*>(y + (size_t)r * ld_y + c));
                v.x = m.x > 0.0f ? v.x : 0.0f; v.y = m.y > 0.0f ? v.y : 0.0f;
                v.z = m.z > 0.0f ? v.z : 0.0f; v.w = m.w > 0.0f ? v.w : 0.0f;
                *reinterpret_cast<float4*>(g + (size_t)r * ld_g + c) = v;
            }
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    red[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && c < cols) {
        float4 s = red[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) { s.x += red[w][lane].x; s.y += red[w][lane].y; s.z += red[w][lane].z; s.w += red[w][lane].w; }
        *reinterpret_cast<float4*>(partial + (size_t)blockIdx.y * cols + c) = s;
    }
}

// one warp per column: lane l adds chunks l, l + 32, ... in order, then a fixed butterfly (deterministic)
__global__ void __launch_bounds__(256) colsum_stage2_kernel(const float* __restrict__ partial, int chunks, int cols, float* __restrict__ out) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= cols) return;
    float s = 0.0f;
    for (int k = lane; k < chunks; k += 32) s += partial[(size_t)k * cols + c];
    s = warp_sum(s);
    if (lane == 0) out[c] = s;
}

}  // namespace spair

using namespace spair;

extern "C" int spair_transpose_batched(const float* in, int B, int R, int C, float* out, void* stream) {
    SPAIR_REQUIRE(in && out && B > 0 && B <= 65535 && R > 0 && C > 0);
    dim3 grid((C + 31) / 32, (R + 31) / 32, B);
    SPAIR_REQUIRE(grid.y <= 65535);
    transpose_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, R, C);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_colsum_chunks(int rows, int cols) {
    const int per = colsum_rows_per_cta(rows, cols);
    return (rows + per - 1) / per;
}

extern "C" int spair_relu_bwd_colsum(float* g, int ld_g, const float* y, int ld_y, int rows, int cols, float* ws, float* out,
                                     void* stream) {
    SPAIR_REQUIRE(g && ws && out && rows > 0 && cols > 0 && (cols & 3) == 0 && (ld_g & 3) == 0 && ((uintptr_t)g & 15) == 0);
    SPAIR_REQUIRE(!y || ((ld_y & 3) == 0 && ((uintptr_t)y & 15) == 0));
    const int per = colsum_rows_per_cta(rows, cols);
    const int chunks = (rows + per - 1) / per;
    SPAIR_REQUIRE(chunks <= 65535);
    dim3 grid((cols + 127) / 128, chunks);
    colsum_stage1_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, ld_g, y, ld_y, rows, cols, per, ws);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    colsum_stage2_kernel<<<(cols + 7) / 8, 256, 0, (cudaStream_t)stream>>>(ws, chunks, cols, out);
    SPAIR_LAUNCH_CHECK();
}

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
