// Backbone stem: ZeroPad2d + first Conv2d (C -> 128 channels, 4x4, stride s) + bias + ReLU of the reference's Backbone
// (modules.py:12-111, topology config.py DEFAULT_BACKBONE_TOPOLOGY[0]) as ONE forward and ONE backward kernel.
//
// The layer has only C*16 taps per output, so it is HBM work, not a GEMM: at the default config it produces a
// [256,128,50,50] map (328 MB).  The library path runs it as conv (writes 328 MB) + bias pass (reads + writes 656 MB,
// un-vectorised) + ReLU pass (656 MB) forward, and ReLU-mask pass + bias reduction + weight-gradient kernel backward
// (each streaming the 328 MB gradient map again).  Here the forward writes the map once (bias and ReLU in registers)
// and the backward reads the gradient map and the saved output once, producing dW and db together; the image has no
// gradient in the model, so no dgrad exists.
//
// forward : thread = PP output pixels x all output channels; the C*16 input taps of a pixel live in registers and are
//           reused for every channel, weights / bias are 128-bit shared-memory broadcasts, stores are coalesced rows
//           (NCHW) or, for the GEMM tail that consumes the map channels-last, eight channels = one 32-byte sector per
//           pixel and store pair (NHWC: no transpose pass between the stem and conv_1, either way).
// backward: dW[o][t] = sum_{b,p} g[b,o,p] * tap[b,p,t], g = dY * (Y > 0), is a [128 x P] x [P x (T+1)] contraction with
//           P = B*Ho*Wo (640 k) — persistent CTAs walk 64-pixel segments, stage g (transposed) and the taps (plus a
//           ones column that yields db) in shared memory, and keep a 4 x 4 register tile per thread; per-CTA partials
//           go to a workspace and a second kernel adds them in a fixed order (deterministic, no atomics).
#include "common.cuh"

namespace spair {

constexpr int kStemK = 4;                 // kernel side
constexpr int kStemCout = 128;
constexpr int kStemFwdThreads = 128;
constexpr int kStemSeg = 64;              // pixels per backward segment
constexpr int kStemGPitch = kStemCout + 4;   // [pixel][channel] gradient tile pitch: 16-byte aligned rows (NHWC staging stores
                                             // 128 bits), conflict-free column writes for consecutive pixels (NCHW staging)

struct StemArgs {
    const float* x;      // [B,C,Ih,Iw]
    const float* w;      // [Cout,C,4,4]
    const float* bias;   // [Cout]
    float* y;            // [B,Cout,Ho,Wo] post-ReLU (NHWC variant: [B,Ho,Wo,Cout])
    int B, Ih, Iw, Ho, Wo, stride, pad_t, pad_l;
};

template <int C>
__device__ __forceinline__ void load_taps(const float* __restrict__ xb, int Ih, int Iw, int oy, int ox, int stride, int pad_t,
                                          int pad_l, bool valid, float (&tap)[C * 16]) {
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int ky = 0; ky < kStemK; ++ky) {
            const int iy = oy * stride - pad_t + ky;
#pragma unroll
            for (int kx = 0; kx < kStemK; ++kx) {
                const int ix = ox * stride - pad_l + kx;
                const bool in = valid && iy >= 0 && iy < Ih && ix >= 0 && ix < Iw;
                tap[(c * kStemK + ky) * kStemK + kx] = in ? __ldg(xb + ((size_t)c * Ih + iy) * Iw + ix) : 0.0f;
            }
        }
}

template <int C, int PP, bool NHWC>
__global__ void __launch_bounds__(kStemFwdThreads) stem_fwd_kernel(StemArgs p) {
    constexpr int T = C * 16;
    __shared__ __align__(16) float w_s[kStemCout * T];
    __shared__ float b_s[kStemCout];
    for (int i = threadIdx.x; i < kStemCout * T; i += kStemFwdThreads) w_s[i] = __ldg(p.w + i);
    for (int i = threadIdx.x; i < kStemCout; i += kStemFwdThreads) b_s[i] = __ldg(p.bias + i);

    const int b = blockIdx.y, npix = p.Ho * p.Wo;
    const int pix0 = blockIdx.x * (kStemFwdThreads * PP) + threadIdx.x;
    const float* xb = p.x + (size_t)b * C * p.Ih * p.Iw;
    float tap[PP][T];
    bool valid[PP];
#pragma unroll
    for (int j = 0; j < PP; ++j) {
        const int pix = pix0 + j * kStemFwdThreads;
        valid[j] = pix < npix;
        const int oy = pix / p.Wo, ox = pix - oy * p.Wo;
        load_taps<C>(xb, p.Ih, p.Iw, oy, ox, p.stride, p.pad_t, p.pad_l, valid[j], tap[j]);
    }
    __syncthreads();
    float* yb = p.y + (size_t)b * kStemCout * npix;
    if constexpr (NHWC) {
        // eight output channels per pass: a pixel's eight values are one 32-byte sector, written by two 128-bit stores of
        // the same thread (same FMA order per output as the NCHW variant: bitwise the same values)
#pragma unroll 1
        for (int o = 0; o < kStemCout; o += 8) {
            float acc[PP][8];
#pragma unroll
            for (int j = 0; j < PP; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = b_s[o + i];
#pragma unroll
            for (int t4 = 0; t4 < T / 4; ++t4) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 w = *reinterpret_cast<const float4*>(w_s + (o + i) * T + 4 * t4);
#pragma unroll
                    for (int j = 0; j < PP; ++j) {
                        acc[j][i] = fmaf(w.x, tap[j][4 * t4 + 0], acc[j][i]);
                        acc[j][i] = fmaf(w.y, tap[j][4 * t4 + 1], acc[j][i]);
                        acc[j][i] = fmaf(w.z, tap[j][4 * t4 + 2], acc[j][i]);
                        acc[j][i] = fmaf(w.w, tap[j][4 * t4 + 3], acc[j][i]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < PP; ++j)
                if (valid[j]) {
                    float4* dst = reinterpret_cast<float4*>(yb + (size_t)(pix0 + j * kStemFwdThreads) * kStemCout + o);
                    dst[0] = make_float4(fmaxf(acc[j][0], 0.0f), fmaxf(acc[j][1], 0.0f), fmaxf(acc[j][2], 0.0f), fmaxf(acc[j][3], 0.0f));
                    dst[1] = make_float4(fmaxf(acc[j][4], 0.0f), fmaxf(acc[j][5], 0.0f), fmaxf(acc[j][6], 0.0f), fmaxf(acc[j][7], 0.0f));
                }
        }
    } else {
#pragma unroll 2
        for (int o = 0; o < kStemCout; ++o) {
            float acc[PP];
#pragma unroll
            for (int j = 0; j < PP; ++j) acc[j] = b_s[o];
#pragma unroll
            for (int t4 = 0; t4 < T / 4; ++t4) {
                const float4 w = *reinterpret_cast<const float4*>(w_s + o * T + 4 * t4);
#pragma unroll
                for (int j = 0; j < PP; ++j) {
                    acc[j] = fmaf(w.x, tap[j][4 * t4 + 0], acc[j]);
                    acc[j] = fmaf(w.y, tap[j][4 * t4 + 1], acc[j]);
                    acc[j] = fmaf(w.z, tap[j][4 * t4 + 2], acc[j]);
                    acc[j] = fmaf(w.w, tap[j][4 * t4 + 3], acc[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < PP; ++j)
                if (valid[j]) yb[(size_t)o * npix + pix0 + j * kStemFwdThreads] = fmaxf(acc[j], 0.0f);
        }
    }
}

struct StemBwdArgs {
    const float* x;      // [B,C,Ih,Iw]
    const float* y;      // [B,Cout,Ho,Wo] saved post-ReLU output
    const float* dy;     // [B,Cout,Ho,Wo]
    float* ws;           // [Cout * TP][gridDim.x] per-CTA partial sums (CTA index fastest: the reduction reads rows)
    int B, Ih, Iw, Ho, Wo, stride, pad_t, pad_l, n_seg_per_image;
    int vec_ok;          // rows of y / dy are 16-byte aligned: 128-bit staging loads
    int nhwc;            // y / dy are [B,Ho,Wo,Cout]: a segment is one contiguous 64 x 512-byte block
};

// TP = padded tap count incl. the ones column (multiple of 4); threads = 32 * TP / 4
template <int C>
__global__ void __launch_bounds__(32 * (C * 16 + 4) / 4) stem_bwd_kernel(StemBwdArgs p) {
    constexpr int T = C * 16, TP = T + 4, TG = TP / 4, NT = 32 * TG;
    __shared__ __align__(16) float g_s[kStemSeg * kStemGPitch];
    __shared__ __align__(16) float in_s[kStemSeg * TP];
    const int og = threadIdx.x & 31, tg = threadIdx.x >> 5;
    const int npix = p.Ho * p.Wo;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = 0.0f;

    const int n_seg = p.B * p.n_seg_per_image;
    for (int seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
        const int b = seg / p.n_seg_per_image, pix0 = (seg - b * p.n_seg_per_image) * kStemSeg;
        __syncthreads();                                        // previous segment consumed
        // taps (+ ones column) of the segment's pixels
        const float* xb = p.x + (size_t)b * C * p.Ih * p.Iw;
        for (int idx = threadIdx.x; idx < kStemSeg * TP; idx += NT) {
            const int pl = idx / TP, t = idx - pl * TP;
            const int pix = pix0 + pl;
            float v = 0.0f;
            if (pix < npix) {
                if (t < T) {
                    const int oy = pix / p.Wo, ox = pix - oy * p.Wo;
                    const int c = t >> 4, ky = (t >> 2) & 3, kx = t & 3;
                    const int iy = oy * p.stride - p.pad_t + ky, ix = ox * p.stride - p.pad_l + kx;
                    if (iy >= 0 && iy < p.Ih && ix >= 0 && ix < p.Iw) v = __ldg(xb + ((size_t)c * p.Ih + iy) * p.Iw + ix);
                } else if (t == T) {
                    v = 1.0f;
                }
            }
            in_s[idx] = v;
        }
        // masked gradient, transposed to [pixel][channel]; 128-bit streaming loads along the pixels when rows allow it
        const float* yb = p.y + (size_t)b * kStemCout * npix;
        const float* dyb = p.dy + (size_t)b * kStemCout * npix;
        if (p.nhwc) {
            // channels-last: the segment's 64 pixels x 128 channels are contiguous in memory and already [pixel][channel]
            constexpr int O4 = kStemCout / 4;
#pragma unroll 4
            for (int idx = threadIdx.x; idx < kStemSeg * O4; idx += NT) {
                const int pl = idx / O4, o4 = idx - pl * O4;
                float4 g = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (pix0 + pl < npix) {
                    const size_t a = ((size_t)pix0 + pl) * kStemCout + 4 * o4;
                    const float4 yv = __ldcs(reinterpret_cast<const float4*>(yb + a));
                    const float4 dv = __ldcs(reinterpret_cast<const float4*>(dyb + a));
                    g.x = yv.x > 0.0f ? dv.x : 0.0f;
                    g.y = yv.y > 0.0f ? dv.y : 0.0f;
                    g.z = yv.z > 0.0f ? dv.z : 0.0f;
                    g.w = yv.w > 0.0f ? dv.w : 0.0f;
                }
                *reinterpret_cast<float4*>(g_s + pl * kStemGPitch + 4 * o4) = g;
            }
        } else if (p.vec_ok) {
            constexpr int Q = kStemSeg / 4;
#pragma unroll 4
            for (int idx = threadIdx.x; idx < kStemCout * Q; idx += NT) {
                const int o = idx / Q, q = idx - o * Q;
                const int pix = pix0 + 4 * q;
                float4 g = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (pix < npix) {                               // npix % 4 == 0: a quad is entirely inside or outside
                    const size_t a = (size_t)o * npix + pix;
                    const float4 yv = __ldcs(reinterpret_cast<const float4*>(yb + a));
                    const float4 dv = __ldcs(reinterpret_cast<const float4*>(dyb + a));
                    g.x = yv.x > 0.0f ? dv.x : 0.0f;
                    g.y = yv.y > 0.0f ? dv.y : 0.0f;
                    g.z = yv.z > 0.0f ? dv.z : 0.0f;
                    g.w = yv.w > 0.0f ? dv.w : 0.0f;
                }
                float* dst = g_s + (4 * q) * kStemGPitch + o;
                dst[0] = g.x;
                dst[kStemGPitch] = g.y;
                dst[2 * kStemGPitch] = g.z;
                dst[3 * kStemGPitch] = g.w;
            }
        } else {
            for (int idx = threadIdx.x; idx < kStemCout * kStemSeg; idx += NT) {
                const int o = idx / kStemSeg, pl = idx - o * kStemSeg;
                const int pix = pix0 + pl;
                float v = 0.0f;
                if (pix < npix) {
                    const size_t a = (size_t)o * npix + pix;
                    v = (ld_stream(yb + a) > 0.0f) ? ld_stream(dyb + a) : 0.0f;
                }
                g_s[pl * kStemGPitch + o] = v;
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int pl = 0; pl < kStemSeg; ++pl) {
            const float4 v = *reinterpret_cast<const float4*>(in_s + pl * TP + 4 * tg);
            const float* gp = g_s + pl * kStemGPitch + og;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float g = gp[32 * i];
                acc[i][0] = fmaf(g, v.x, acc[i][0]);
                acc[i][1] = fmaf(g, v.y, acc[i][1]);
                acc[i][2] = fmaf(g, v.z, acc[i][2]);
                acc[i][3] = fmaf(g, v.w, acc[i][3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) p.ws[(size_t)((og + 32 * i) * TP + 4 * tg + k) * gridDim.x + blockIdx.x] = acc[i][k];
}

// dW[o][t] = sum over CTAs of ws[o*TP + t][cta] (t < T), db[o] = sum of ws[o*TP + T][cta]: one warp per output, lanes
// stride over the CTAs, butterfly at the end — a fixed order, so the result is reproducible
__global__ void __launch_bounds__(256) stem_bwd_reduce_kernel(const float* __restrict__ ws, int n_cta, int T, int TP,
                                                             float* __restrict__ dW, float* __restrict__ db) {
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (idx >= kStemCout * TP) return;
    const int o = idx / TP, t = idx - o * TP;
    if (t > T) return;
    float s = 0.0f;
    for (int c = lane; c < n_cta; c += 32) s += ws[(size_t)idx * n_cta + c];
    s = warp_sum(s);
    if (lane == 0) {
        if (t < T) dW[o * T + t] = s;
        else db[o] = s;
    }
}

// out[r][:] = row[:] for all r: the bias rows a beta = 1 GEMM then accumulates into (ops.WideLinearFunction)
__global__ void __launch_bounds__(256) broadcast_rows_kernel(const float* __restrict__ row, int rows, int cols, float* __restrict__ out) {
    const size_t total = (size_t)rows * cols;
    if ((cols & 3) == 0) {
        const int c4 = cols >> 2;
        const float4* r4 = reinterpret_cast<const float4*>(row);
        float4* o4 = reinterpret_cast<float4*>(out);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total / 4; i += (size_t)gridDim.x * blockDim.x)
            o4[i] = __ldg(r4 + (int)(i % c4));
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
            out[i] = __ldg(row + (int)(i % cols));
    }
}

}  // namespace spair

using namespace spair;

static bool stem_shape_ok(int B, int C, int Ih, int Iw, int Cout, int k, int stride, int pad_t, int pad_l, int Ho, int Wo) {
    return B > 0 && (C == 1 || C == 3) && Ih > 0 && Iw > 0 && Cout == kStemCout && k == kStemK && stride >= 1 && pad_t >= 0 &&
           pad_l >= 0 && Ho > 0 && Wo > 0 && (long long)Ho * Wo < (1 << 30);
}

extern "C" int spair_stem_bwd_ctas(void) { return kSMs * 4; }

extern "C" int spair_stem_conv_fwd(const float* x, const float* w, const float* bias, int B, int C, int Ih, int Iw, int Cout,
                                   int k, int stride, int pad_t, int pad_l, int Ho, int Wo, int nhwc, float* y, void* stream) {
    SPAIR_REQUIRE(x && w && bias && y && (!nhwc || ((uintptr_t)y % 16) == 0));
    SPAIR_REQUIRE(stem_shape_ok(B, C, Ih, Iw, Cout, k, stride, pad_t, pad_l, Ho, Wo));
    StemArgs a{x, w, bias, y, B, Ih, Iw, Ho, Wo, stride, pad_t, pad_l};
    const int npix = Ho * Wo;
    if (C == 1) {
        dim3 grid((npix + kStemFwdThreads * 4 - 1) / (kStemFwdThreads * 4), B);
        if (nhwc) stem_fwd_kernel<1, 4, true><<<grid, kStemFwdThreads, 0, (cudaStream_t)stream>>>(a);
        else stem_fwd_kernel<1, 4, false><<<grid, kStemFwdThreads, 0, (cudaStream_t)stream>>>(a);
    } else {
        dim3 grid((npix + kStemFwdThreads - 1) / kStemFwdThreads, B);
        if (nhwc) stem_fwd_kernel<3, 1, true><<<grid, kStemFwdThreads, 0, (cudaStream_t)stream>>>(a);
        else stem_fwd_kernel<3, 1, false><<<grid, kStemFwdThreads, 0, (cudaStream_t)stream>>>(a);
    }
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_stem_conv_bwd(const float* x, const float* y, const float* dy, int B, int C, int Ih, int Iw, int Cout,
                                   int k, int stride, int pad_t, int pad_l, int Ho, int Wo, int nhwc, float* ws, float* d_w,
                                   float* d_bias, void* stream) {
    SPAIR_REQUIRE(x && y && dy && ws && d_w && d_bias && ((uintptr_t)ws % 16) == 0);
    SPAIR_REQUIRE(!nhwc || (((uintptr_t)y % 16) == 0 && ((uintptr_t)dy % 16) == 0));
    SPAIR_REQUIRE(stem_shape_ok(B, C, Ih, Iw, Cout, k, stride, pad_t, pad_l, Ho, Wo));
    const int n_seg_per_image = (Ho * Wo + kStemSeg - 1) / kStemSeg;
    const long long n_seg = (long long)B * n_seg_per_image;
    const int n_cta = (int)(n_seg < spair_stem_bwd_ctas() ? n_seg : spair_stem_bwd_ctas());
    const int vec_ok = ((Ho * Wo) % 4 == 0) && ((uintptr_t)y % 16) == 0 && ((uintptr_t)dy % 16) == 0;
    StemBwdArgs a{x, y, dy, ws, B, Ih, Iw, Ho, Wo, stride, pad_t, pad_l, n_seg_per_image, vec_ok, nhwc};
    const int T = C * 16, TP = T + 4;
    if (C == 1) stem_bwd_kernel<1><<<n_cta, 32 * (16 + 4) / 4, 0, (cudaStream_t)stream>>>(a);
    else stem_bwd_kernel<3><<<n_cta, 32 * (48 + 4) / 4, 0, (cudaStream_t)stream>>>(a);
    stem_bwd_reduce_kernel<<<(kStemCout * TP + 7) / 8, 256, 0, (cudaStream_t)stream>>>(ws, n_cta, T, TP, d_w, d_bias);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_broadcast_rows(const float* row, int rows, int cols, float* out, void* stream) {
    SPAIR_REQUIRE(row && out && rows > 0 && cols > 0);
    SPAIR_REQUIRE((cols % 4) != 0 || (((uintptr_t)row % 16) == 0 && ((uintptr_t)out % 16) == 0));
    broadcast_rows_kernel<<<kSMs * 8, 256, 0, (cudaStream_t)stream>>>(row, rows, cols, out);
    SPAIR_LAUNCH_CHECK();
}
