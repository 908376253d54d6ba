// G: spatial-transformer glimpse extractor and the generic paste (inverse stn).
//
// glimpse = stn(image, z_where, [Gh,Gw], inverse=False) (reference modules.py:216-273): fused
// affine_grid + bilinear grid_sample, border padding, align_corners=False.
//
// Layout / schedule: one CTA per object.  The object's source window (<= 48 px + 1 on a side for
// model-generated boxes) is staged once per channel into shared memory with 128-bit loads
// (rows are 16-byte aligned when Iw % 4 == 0); the per-column and per-row sample coordinates
// are computed once per CTA into shared tables, so the texel loop is 4 LDS + 4 FMA.  Windows
// that do not fit the tile (only reachable through the generic stn() API) sample global memory
// directly.  Backward wrt z_where is a per-object reduction (no atomics); the optional image
// gradient (generic stn() API only — the model's image has no grad) is scattered with atomics.
#include <stdlib.h>

#include "warp_math.cuh"

namespace spair {

constexpr int kGlimpseThreads = 256;
constexpr int kTileW = 72;   // floats, multiple of 4
constexpr int kTileH = 68;

struct Window {
    int x_lo, y_lo, tw, th;   // tile origin (x aligned down to 4) and extent
    bool staged;
};

// Fills the per-column / per-row coordinate tables (clipped source coordinate and the clip mask
// of grid_sampler's clip_coordinates_set_grad) and derives the window to stage.
__device__ __forceinline__ Window glimpse_setup(const FwdAffine& A, int Ih, int Iw, int Gh, int Gw, float* col_ix,
                                                float* row_iy, float* col_m, float* row_m, bool aligned) {
    for (int j = threadIdx.x; j < Gw; j += blockDim.x) {
        float ix = unnormalize(affine_coord(base_coord(j, Gw), A.ax, A.cx), 0.5f * (float)Iw);
        float m = 1.0f;
        if (ix <= 0.0f) { ix = 0.0f; m = 0.0f; }
        else if (ix >= (float)(Iw - 1)) { ix = (float)(Iw - 1); m = 0.0f; }
        col_ix[j] = ix;
        if (col_m) col_m[j] = m;
    }
    for (int i = threadIdx.x; i < Gh; i += blockDim.x) {
        float iy = unnormalize(affine_coord(base_coord(i, Gh), A.ay, A.cy), 0.5f * (float)Ih);
        float m = 1.0f;
        if (iy <= 0.0f) { iy = 0.0f; m = 0.0f; }
        else if (iy >= (float)(Ih - 1)) { iy = (float)(Ih - 1); m = 0.0f; }
        row_iy[i] = iy;
        if (row_m) row_m[i] = m;
    }
    __syncthreads();
    // coordinates are monotone in j (either direction): the window is spanned by the end points
    const float xa = col_ix[0], xb = col_ix[Gw - 1], ya = row_iy[0], yb = row_iy[Gh - 1];
    const int x0 = (int)floorf(fminf(xa, xb)), x1 = min((int)floorf(fmaxf(xa, xb)) + 1, Iw - 1);
    const int y0 = (int)floorf(fminf(ya, yb)), y1 = min((int)floorf(fmaxf(ya, yb)) + 1, Ih - 1);
    Window w;
    w.x_lo = x0 & ~3;
    w.y_lo = y0;
    const int x_end = min((x1 + 4) & ~3, Iw);
    w.tw = x_end - w.x_lo;
    w.th = y1 - y0 + 1;
    w.staged = aligned && w.tw <= kTileW && w.th <= kTileH;
    return w;
}

__device__ __forceinline__ void stage_window(const float* __restrict__ plane, int Iw, const Window& w,
                                             float* __restrict__ tile) {
    // 16 lanes per window row (a row holds at most kTileW / 4 = 18 128-bit vectors): no integer division
    const int vec_per_row = w.tw >> 2;
    const int sub = threadIdx.x & 15, rows_per_pass = blockDim.x >> 4;
#pragma unroll 1
    for (int ry = threadIdx.x >> 4; ry < w.th; ry += rows_per_pass) {
        const float4* src = reinterpret_cast<const float4*>(plane + (long long)(w.y_lo + ry) * Iw + w.x_lo);
        float4* dst = reinterpret_cast<float4*>(tile + ry * w.tw);
#pragma unroll 1
        for (int v = sub; v < vec_per_row; v += 16) dst[v] = __ldg(src + v);
    }
}

struct Taps {
    float v00, v01, v10, v11;
};

__device__ __forceinline__ Taps fetch_taps(const float* __restrict__ plane, int Iw, int Ih, const Window& w,
                                           const float* __restrict__ tile, int x0, int y0) {
    Taps t;
    const bool x1ok = x0 + 1 <= Iw - 1, y1ok = y0 + 1 <= Ih - 1;
    if (w.staged) {
        const float* p = tile + (y0 - w.y_lo) * w.tw + (x0 - w.x_lo);
        t.v00 = p[0];
        t.v01 = x1ok ? p[1] : 0.0f;
        t.v10 = y1ok ? p[w.tw] : 0.0f;
        t.v11 = (x1ok && y1ok) ? p[w.tw + 1] : 0.0f;
    } else {
        const float* p = plane + (long long)y0 * Iw + x0;
        t.v00 = __ldg(p);
        t.v01 = x1ok ? __ldg(p + 1) : 0.0f;
        t.v10 = y1ok ? __ldg(p + Iw) : 0.0f;
        t.v11 = (x1ok && y1ok) ? __ldg(p + Iw + 1) : 0.0f;
    }
    return t;
}

__global__ void __launch_bounds__(kGlimpseThreads)
glimpse_fwd_kernel(const float* __restrict__ image, const float* __restrict__ z_where, const int* __restrict__ cells,
                   int B, int HW, int C, int Ih, int Iw, int Gh, int Gw, float* __restrict__ out, int ld_out,
                   int aligned, float inv_Gw) {
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;
    float* col_ix = tile + kTileW * kTileH;
    float* row_iy = col_ix + Gw;
    const int b = blockIdx.x;                                   // image
    const int r = cells ? blockIdx.y * B + b : b;               // output row (k*B + b)
    const long long o = cells ? (long long)b * HW + cells[blockIdx.y] : b;
    const float4 zw = *reinterpret_cast<const float4*>(z_where + o * 4);
    const FwdAffine A(zw.x, zw.y, zw.z, zw.w);
    const Window w = glimpse_setup(A, Ih, Iw, Gh, Gw, col_ix, row_iy, nullptr, nullptr, aligned != 0);
    const int GG = Gh * Gw;
    const bool small = GG < (1 << 15) && Gw <= 1024;
    for (int c = 0; c < C; ++c) {
        const float* plane = image + ((long long)b * C + c) * Ih * Iw;
        if (w.staged) {
            stage_window(plane, Iw, w, tile);
            __syncthreads();
        }
        float* orow = out + (long long)r * ld_out + (long long)c * GG;
        for (int t = threadIdx.x; t < GG; t += blockDim.x) {
            const int i = small ? fast_div(t, inv_Gw) : t / Gw, j = t - i * Gw;
            const float ix = col_ix[j], iy = row_iy[i];
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            const float wx1 = ix - fx0, wx0 = fx0 + 1.0f - ix, wy1 = iy - fy0, wy0 = fy0 + 1.0f - iy;
            const Taps v = fetch_taps(plane, Iw, Ih, w, tile, (int)fx0, (int)fy0);
            float acc = __fmul_rn(v.v00, __fmul_rn(wx0, wy0));
            acc = fmaf(v.v01, __fmul_rn(wx1, wy0), acc);
            acc = fmaf(v.v10, __fmul_rn(wx0, wy1), acc);
            acc = fmaf(v.v11, __fmul_rn(wx1, wy1), acc);
            orow[t] = acc;
        }
        if (w.staged) __syncthreads();
    }
}

__global__ void __launch_bounds__(kGlimpseThreads)
glimpse_bwd_kernel(const float* __restrict__ image, const float* __restrict__ z_where, const int* __restrict__ cells,
                   int B, int HW, int C, int Ih, int Iw, int Gh, int Gw, const float* __restrict__ d_out, int ld_out,
                   float* __restrict__ d_zw, float* __restrict__ d_image, int aligned, float inv_Gw) {
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;
    float* col_ix = tile + kTileW * kTileH;
    float* row_iy = col_ix + Gw;
    float* col_m = row_iy + Gh;
    float* row_m = col_m + Gw;
    float* col_b = row_m + Gh;      // normalised base coordinates (d gx / d xs, d gy / d ys)
    float* row_b = col_b + Gw;
    __shared__ float red[4 * (kGlimpseThreads / 32)];
    for (int j = threadIdx.x; j < Gw; j += blockDim.x) col_b[j] = base_coord(j, Gw);
    for (int i = threadIdx.x; i < Gh; i += blockDim.x) row_b[i] = base_coord(i, Gh);
    const int b = blockIdx.x;                                   // image
    const int r = cells ? blockIdx.y * B + b : b;               // output row (k*B + b)
    const long long o = cells ? (long long)b * HW + cells[blockIdx.y] : b;
    const float4 zw = *reinterpret_cast<const float4*>(z_where + o * 4);
    const FwdAffine A(zw.x, zw.y, zw.z, zw.w);
    const Window w = glimpse_setup(A, Ih, Iw, Gh, Gw, col_ix, row_iy, col_m, row_m, aligned != 0);
    const int GG = Gh * Gw;
    const bool small = GG < (1 << 15) && Gw <= 1024;
    // acc[0] = sum dL/dgx, acc[1] = sum dL/dgy, acc[2] = sum dL/dgx * base_x, acc[3] = sum dL/dgy * base_y
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int c = 0; c < C; ++c) {
        const float* plane = image + ((long long)b * C + c) * Ih * Iw;
        float* gplane = d_image ? d_image + ((long long)b * C + c) * Ih * Iw : nullptr;
        if (w.staged) {
            stage_window(plane, Iw, w, tile);
            __syncthreads();
        }
        const float* grow = d_out + (long long)r * ld_out + (long long)c * GG;
        for (int t = threadIdx.x; t < GG; t += blockDim.x) {
            const int i = small ? fast_div(t, inv_Gw) : t / Gw, j = t - i * Gw;
            const float ix = col_ix[j], iy = row_iy[i];
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            const float wx1 = ix - fx0, wx0 = fx0 + 1.0f - ix, wy1 = iy - fy0, wy0 = fy0 + 1.0f - iy;
            const int x0 = (int)fx0, y0 = (int)fy0;
            const Taps v = fetch_taps(plane, Iw, Ih, w, tile, x0, y0);
            const float g = grow[t];
            // grid_sampler_2d_backward: d out / d ix and d out / d iy
            const float gix = g * ((v.v01 - v.v00) * wy0 + (v.v11 - v.v10) * wy1);
            const float giy = g * ((v.v10 - v.v00) * wx0 + (v.v11 - v.v01) * wx1);
            const float dgx = gix * col_m[j], dgy = giy * row_m[i];
            acc[0] += dgx;
            acc[1] += dgy;
            acc[2] = fmaf(dgx, col_b[j], acc[2]);
            acc[3] = fmaf(dgy, row_b[i], acc[3]);
            if (gplane) {
                const bool x1ok = x0 + 1 <= Iw - 1, y1ok = y0 + 1 <= Ih - 1;
                float* p = gplane + (long long)y0 * Iw + x0;
                atomicAdd(p, g * wx0 * wy0);
                if (x1ok) atomicAdd(p + 1, g * wx1 * wy0);
                if (y1ok) atomicAdd(p + Iw, g * wx0 * wy1);
                if (x1ok && y1ok) atomicAdd(p + Iw + 1, g * wx1 * wy1);
            }
        }
        if (w.staged) __syncthreads();
    }
    block_sum<4>(acc, red);
    if (threadIdx.x == 0) {
        const float hx = 0.5f * (float)Iw, hy = 0.5f * (float)Ih;   // d ix / d gx (unnormalize)
        // gx = xs * base + (2 xt - 1)
        d_zw[(long long)r * 4 + 0] = 2.0f * hx * acc[0];
        d_zw[(long long)r * 4 + 1] = 2.0f * hy * acc[1];
        d_zw[(long long)r * 4 + 2] = hx * acc[2];
        d_zw[(long long)r * 4 + 3] = hy * acc[3];
    }
}

// ------------------------------------------------------------------------------------------
// Image-resident variant (the model's path: many cells of the SAME image).  The per-object kernels above re-stage a
// ~49x49 window per object — 3x more staged than sampled, plus the coordinate set-up of a whole CTA per object.  Here a
// persistent CTA copies the whole image (C*Ih*Iw fp32 <= 96 KB: configs A and C) into shared memory with TMA bulk copies
// behind an mbarrier and every warp extracts whole glimpses from it: the image is read from HBM/L2 once per CTA that
// touches it, the only other traffic is the output (forward) / d_out (backward), written / read in 128-byte rows.
// Work is split evenly over the grid in (image, cell) order, so a CTA loads at most two images.
// ------------------------------------------------------------------------------------------
constexpr int kResThreads = 384, kResWarps = kResThreads / 32;   // 3 CTAs x 12 warps per SM at a 64 KB image
constexpr int kResMaxImageBytes = 96 * 1024;

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// whole image -> shared memory: bulk async copies (TMA, SASS UBLKCP) of <= 32 KB, completion counted on the mbarrier
__device__ __forceinline__ void load_image_async(float* dst, const float* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    for (uint32_t off = 0; off < bytes; off += 32768u) {
        const uint32_t n = min(32768u, bytes - off);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         g_smem_u32(dst) + off),
                     "l"(reinterpret_cast<const char*>(src) + off), "r"(n), "r"(bar)
                     : "memory");
    }
}

struct ResArgs {
    const float* image;
    const float* z_where;
    const int* cells;
    int n_cells, B, HW, C, Ih, Iw, Gh, Gw;
    float* out;            // forward: glimpses; backward: unused
    const float* d_out;    // backward
    float* d_zw;           // backward: [n_cells*B, 4]
    int ld_out;
    int per_cta;           // objects per CTA
    float inv_Gw;
};

// One axis of an object's sampling grid at glimpse index `base` (normalised coordinate): {first tap offset, second tap offset
// (clamped neighbour), weight of the first tap, weight of the second} with offsets pre-multiplied by `pitch`, and the
// clip mask of grid_sampler's clip_coordinates_set_grad.  The sample coordinate is clamped to [0, size-1]
// (padding_mode='border'); at the upper border the neighbour tap has weight exactly 0, so reading the clamped texel
// instead of skipping the tap gives the same bits.
__device__ __forceinline__ float4 res_axis(float base, float a, float c, int size, int pitch, float& mask) {
    float ix = unnormalize(affine_coord(base, a, c), 0.5f * (float)size);
    mask = 1.0f;
    if (ix <= 0.0f) { ix = 0.0f; mask = 0.0f; }
    else if (ix >= (float)(size - 1)) { ix = (float)(size - 1); mask = 0.0f; }
    const float f0 = floorf(ix);
    const int x0 = (int)f0;
    return make_float4(__int_as_float(x0 * pitch), __int_as_float(min(x0 + 1, size - 1) * pitch), f0 + 1.0f - ix, ix - f0);
}

// Thread mapping: a lane owns ONE glimpse column j for the whole kernel (its taps' x offsets and weights stay in registers
// per object) and the warp walks the glimpse rows, 32 / Gw of them per iteration (28 of 32 lanes busy at Gw = 28 or 14);
// the row entry is a broadcast shared-memory read.  Per texel that leaves the four gathers from the resident image, four
// weight products, the interpolation and one coalesced store (forward) or load (backward).  Gw > 32: columns in chunks of 32.
template <bool BWD>
__global__ void __launch_bounds__(kResThreads, 3) glimpse_resident_kernel(ResArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int plane = p.Ih * p.Iw, img_floats = p.C * plane;
    float* img = smem;
    float* base_x = img + img_floats;
    float* base_y = base_x + p.Gw;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* rows4 = reinterpret_cast<float4*>(smem + ((img_floats + p.Gw + p.Gh + 3) & ~3));
    float4* row = rows4 + (size_t)warp * p.Gh;
    float* row_m = BWD ? reinterpret_cast<float*>(rows4 + (size_t)kResWarps * p.Gh) + (size_t)warp * p.Gh : nullptr;   // clamp masks
    __shared__ __align__(8) uint64_t bar_storage;
    const uint32_t bar = g_smem_u32(&bar_storage);

    for (int j = threadIdx.x; j < p.Gw; j += kResThreads) base_x[j] = base_coord(j, p.Gw);
    for (int i = threadIdx.x; i < p.Gh; i += kResThreads) base_y[i] = base_coord(i, p.Gh);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // lane -> (row within an iteration, column); constant for the whole kernel
    const int rpi = p.Gw <= 32 ? fast_div(32, p.inv_Gw) : 1;          // glimpse rows per warp iteration
    const int sub = p.Gw <= 32 ? fast_div(lane, p.inv_Gw) : 0;
    const int j_lane = p.Gw <= 32 ? lane - sub * p.Gw : lane;
    const bool lane_on = sub < rpi;

    const long long total = (long long)p.B * p.n_cells;
    const long long o0 = (long long)blockIdx.x * p.per_cta, o1 = min(total, o0 + p.per_cta);
    const int GG = p.Gh * p.Gw;
    uint32_t parity = 0;
    for (int b = (int)(o0 / p.n_cells); (long long)b * p.n_cells < o1; ++b) {
        const int k0 = (int)max(o0 - (long long)b * p.n_cells, 0LL), k1 = (int)min(o1 - (long long)b * p.n_cells, (long long)p.n_cells);
        if (threadIdx.x == 0) load_image_async(img, p.image + (size_t)b * img_floats, (uint32_t)img_floats * 4u, bar);
        g_mbar_wait(bar, parity);
        parity ^= 1;
        for (int k = k0 + warp; k < k1; k += kResWarps) {
            const int cell = p.cells[k];
            const long long r = (long long)k * p.B + b;                 // wavefront-major output row
            const float4 zw = __ldg(reinterpret_cast<const float4*>(p.z_where) + (size_t)b * p.HW + cell);
            const FwdAffine A(zw.x, zw.y, zw.z, zw.w);
            for (int i = lane; i < p.Gh; i += 32) {
                float m;
                row[i] = res_axis(base_y[i], A.ay, A.cy, p.Ih, p.Iw, m);
                if (BWD) row_m[i] = m;
            }
            __syncwarp();
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int j = j_lane; j < p.Gw; j += 32) {
                float col_m;
                const float bx = base_x[j];
                const float4 cj = res_axis(bx, A.ax, A.cx, p.Iw, 1, col_m);
                const int x0 = __float_as_int(cj.x), x1 = __float_as_int(cj.y);
                // shared-memory byte addresses of the two tap columns (32-bit arithmetic, ld.shared) and a running output /
                // gradient pointer: the loop body is 1 broadcast + 4 gathers + the interpolation + one coalesced access
                const uint32_t img_s = g_smem_u32(img);
                const uint32_t row_s = g_smem_u32(row);
                for (int c = 0; c < p.C; ++c) {
                    const uint32_t a0 = img_s + 4u * (uint32_t)(c * plane + x0), a1 = img_s + 4u * (uint32_t)(c * plane + x1);
                    if (!lane_on) continue;
                    const size_t o_first = r * p.ld_out + (size_t)c * GG + j + (size_t)sub * p.Gw;
                    float* op = BWD ? nullptr : p.out + o_first;
                    const float* gp = BWD ? p.d_out + o_first : nullptr;
                    const int step = rpi * p.Gw;
                    constexpr int U = 7;      // rows per batch: the gradient loads of a whole batch are issued before its gathers
                    for (int i0 = sub; i0 < p.Gh; i0 += U * rpi) {
                        float g[U];
                        if (BWD) {
#pragma unroll
                            for (int u = 0; u < U; ++u) g[u] = (i0 + u * rpi < p.Gh) ? __ldg(gp + (size_t)u * step) : 0.0f;
                            gp += (size_t)U * step;
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int i = i0 + u * rpi;
                            if (i >= p.Gh) break;
                            float4 ri;
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(ri.x), "=f"(ri.y), "=f"(ri.z), "=f"(ri.w) : "r"(row_s + 16u * (uint32_t)i));
                            const uint32_t y0 = 4u * (uint32_t)__float_as_int(ri.x), y1 = 4u * (uint32_t)__float_as_int(ri.y);
                            float v00, v01, v10, v11;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v00) : "r"(a0 + y0));
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v01) : "r"(a1 + y0));
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v10) : "r"(a0 + y1));
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v11) : "r"(a1 + y1));
                            if (!BWD) {
                                float a = __fmul_rn(v00, __fmul_rn(cj.z, ri.z));
                                a = fmaf(v01, __fmul_rn(cj.w, ri.z), a);
                                a = fmaf(v10, __fmul_rn(cj.z, ri.w), a);
                                a = fmaf(v11, __fmul_rn(cj.w, ri.w), a);
                                *op = a;
                                op += step;
                            } else {
                                // grid_sampler_2d_backward: d out / d ix, d out / d iy; zero where the coordinate was clamped
                                const float dgx = g[u] * ((v01 - v00) * ri.z + (v11 - v10) * ri.w) * col_m;
                                const float dgy = g[u] * ((v10 - v00) * cj.z + (v11 - v01) * cj.w) * row_m[i];
                                acc[0] += dgx;
                                acc[1] += dgy;
                                acc[2] = fmaf(dgx, bx, acc[2]);
                                acc[3] = fmaf(dgy, base_y[i], acc[3]);
                            }
                        }
                    }
                }
            }
            if (BWD) {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = warp_sum(acc[q]);
                if (lane == 0) {
                    const float hx = 0.5f * (float)p.Iw, hy = 0.5f * (float)p.Ih;   // gx = xs * base + (2 xt - 1)
                    *reinterpret_cast<float4*>(p.d_zw + r * 4) = make_float4(2.0f * hx * acc[0], 2.0f * hy * acc[1], hx * acc[2], hy * acc[3]);
                }
            }
            __syncwarp();
        }
        __syncthreads();      // every warp is done with this image before the next bulk copy overwrites it
    }
}

static size_t resident_smem(int C, int Ih, int Iw, int Gh, int Gw, bool bwd) {
    const size_t head = ((size_t)C * Ih * Iw + Gw + Gh + 3) & ~(size_t)3;
    return sizeof(float) * head + (size_t)kResWarps * Gh * (bwd ? 20 : 16);
}

// The resident path needs the model's call shape (cells of shared images), a TMA-copyable image and no image gradient.
static bool resident_ok(const float* image, const int* cells, int C, int Ih, int Iw, int Gh, int Gw, const float* d_image) {
    const size_t bytes = (size_t)C * Ih * Iw * 4;
    return cells && !d_image && bytes <= (size_t)kResMaxImageBytes && bytes % 16 == 0 && ((uintptr_t)image % 16) == 0 &&
           (long long)Gh * Gw < (1 << 15) && Gw <= 1024;
}

template <bool BWD>
static int launch_resident(ResArgs a, cudaStream_t st) {
    const size_t smem = resident_smem(a.C, a.Ih, a.Iw, a.Gh, a.Gw, BWD);
    static size_t cache[kMaxDevices] = {};
    cudaError_t e = ensure_dynamic_smem(glimpse_resident_kernel<BWD>, smem, cache);
    if (e != cudaSuccess) return (int)e;
    const long long total = (long long)a.B * a.n_cells;
    const int ctas_per_sm = (int)((220 * 1024) / (smem + 1024)) < 1 ? 1 : (int)((220 * 1024) / (smem + 1024));
    long long per = (total + (long long)kSMs * ctas_per_sm - 1) / ((long long)kSMs * ctas_per_sm);
    if (per < kResWarps) per = kResWarps;
    a.per_cta = (int)per;
    glimpse_resident_kernel<BWD><<<(unsigned)((total + per - 1) / per), kResThreads, smem, st>>>(a);
    SPAIR_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// generic paste: stn(image, z_where, [Oh,Ow], inverse=True), zeros padding (modules.py:255-269)
// ------------------------------------------------------------------------------------------
__global__ void paste_fwd_kernel(const float* __restrict__ image, const float* __restrict__ z_where, int n, int C,
                                 int Gh, int Gw, int Oh, int Ow, float* __restrict__ out) {
    const int obj = blockIdx.y;
    const float4 zw = *reinterpret_cast<const float4*>(z_where + (long long)obj * 4);
    const InvAffine A(zw.x, zw.y, zw.z, zw.w);
    const int npix = Oh * Ow;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        const int Y = p / Ow, X = p - Y * Ow;
        const float ix = unnormalize(affine_coord(base_coord(X, Ow), A.ax, A.cx), 0.5f * (float)Gw);
        const float iy = unnormalize(affine_coord(base_coord(Y, Oh), A.ay, A.cy), 0.5f * (float)Gh);
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const bool inside = fx0 >= -1.0f && fx0 <= (float)(Gw - 1) && fy0 >= -1.0f && fy0 <= (float)(Gh - 1);
        const int x0 = inside ? (int)fx0 : 0, y0 = inside ? (int)fy0 : 0;
        const float wx1 = ix - fx0, wx0 = fx0 + 1.0f - ix, wy1 = iy - fy0, wy0 = fy0 + 1.0f - iy;
        const bool xa = inside && x0 >= 0, xb = inside && x0 + 1 <= Gw - 1, ya = inside && y0 >= 0, yb = inside && y0 + 1 <= Gh - 1;
        for (int c = 0; c < C; ++c) {
            const float* pl = image + ((long long)obj * C + c) * Gh * Gw;
            float acc = 0.0f;
            if (xa && ya) acc = __fmul_rn(pl[y0 * Gw + x0], __fmul_rn(wx0, wy0));
            if (xb && ya) acc = fmaf(pl[y0 * Gw + x0 + 1], __fmul_rn(wx1, wy0), acc);
            if (xa && yb) acc = fmaf(pl[(y0 + 1) * Gw + x0], __fmul_rn(wx0, wy1), acc);
            if (xb && yb) acc = fmaf(pl[(y0 + 1) * Gw + x0 + 1], __fmul_rn(wx1, wy1), acc);
            out[((long long)obj * C + c) * npix + p] = acc;
        }
    }
}

__global__ void paste_bwd_kernel(const float* __restrict__ image, const float* __restrict__ z_where, int n, int C,
                                 int Gh, int Gw, int Oh, int Ow, const float* __restrict__ d_out,
                                 float* __restrict__ d_image, float* __restrict__ d_zw) {
    __shared__ float red[4 * 8];
    const int obj = blockIdx.y;
    const float4 zw = *reinterpret_cast<const float4*>(z_where + (long long)obj * 4);
    const InvAffine A(zw.x, zw.y, zw.z, zw.w);
    const int npix = Oh * Ow;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // sum dgx, sum dgy, sum dgx*bX, sum dgy*bY
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        const int Y = p / Ow, X = p - Y * Ow;
        const float bX = base_coord(X, Ow), bY = base_coord(Y, Oh);
        const float ix = unnormalize(affine_coord(bX, A.ax, A.cx), 0.5f * (float)Gw);
        const float iy = unnormalize(affine_coord(bY, A.ay, A.cy), 0.5f * (float)Gh);
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const bool inside = fx0 >= -1.0f && fx0 <= (float)(Gw - 1) && fy0 >= -1.0f && fy0 <= (float)(Gh - 1);
        if (!inside) continue;
        const int x0 = (int)fx0, y0 = (int)fy0;
        const float wx1 = ix - fx0, wx0 = fx0 + 1.0f - ix, wy1 = iy - fy0, wy0 = fy0 + 1.0f - iy;
        const bool xa = x0 >= 0, xb = x0 + 1 <= Gw - 1, ya = y0 >= 0, yb = y0 + 1 <= Gh - 1;
        float gix = 0.0f, giy = 0.0f;
        for (int c = 0; c < C; ++c) {
            const long long po = ((long long)obj * C + c) * Gh * Gw;
            const float* pl = image + po;
            const float g = d_out[((long long)obj * C + c) * npix + p];
            const float v00 = (xa && ya) ? pl[y0 * Gw + x0] : 0.0f, v01 = (xb && ya) ? pl[y0 * Gw + x0 + 1] : 0.0f;
            const float v10 = (xa && yb) ? pl[(y0 + 1) * Gw + x0] : 0.0f, v11 = (xb && yb) ? pl[(y0 + 1) * Gw + x0 + 1] : 0.0f;
            gix += g * ((v01 - v00) * wy0 + (v11 - v10) * wy1);
            giy += g * ((v10 - v00) * wx0 + (v11 - v01) * wx1);
            if (d_image) {
                float* gp = d_image + po;
                if (xa && ya) atomicAdd(gp + y0 * Gw + x0, g * wx0 * wy0);
                if (xb && ya) atomicAdd(gp + y0 * Gw + x0 + 1, g * wx1 * wy0);
                if (xa && yb) atomicAdd(gp + (y0 + 1) * Gw + x0, g * wx0 * wy1);
                if (xb && yb) atomicAdd(gp + (y0 + 1) * Gw + x0 + 1, g * wx1 * wy1);
            }
        }
        const float dgx = gix * 0.5f * (float)Gw, dgy = giy * 0.5f * (float)Gh;
        acc[0] += dgx;
        acc[1] += dgy;
        acc[2] += dgx * bX;
        acc[3] += dgy * bY;
    }
    block_sum<4>(acc, red);
    if (threadIdx.x == 0 && d_zw) {
        // gx = bX * (1/xs) - (2xt-1)/xs
        const float xtp = 2.0f * zw.x - 1.0f, ytp = 2.0f * zw.y - 1.0f;
        atomicAdd(d_zw + (long long)obj * 4 + 0, -2.0f * acc[0] / zw.z);
        atomicAdd(d_zw + (long long)obj * 4 + 1, -2.0f * acc[1] / zw.w);
        atomicAdd(d_zw + (long long)obj * 4 + 2, (-acc[2] + acc[0] * xtp) / (zw.z * zw.z));
        atomicAdd(d_zw + (long long)obj * 4 + 3, (-acc[3] + acc[1] * ytp) / (zw.w * zw.w));
    }
}

}  // namespace spair

using namespace spair;

static size_t glimpse_smem(int Gh, int Gw, bool bwd) {
    return sizeof(float) * (size_t)(kTileW * kTileH + (bwd ? 3 : 1) * (Gh + Gw));
}

extern "C" int spair_base_grid(int n, float* out) {
    SPAIR_REQUIRE(n > 0 && out);
    for (int j = 0; j < n; ++j) out[j] = base_coord(j, n);
    return 0;
}

extern "C" int spair_glimpse_fwd(const float* image, const float* z_where, const int* cells, int n_cells, int B,
                                 int HW, int C, int Ih, int Iw, int Gh, int Gw, float* out, int ld_out,
                                 void* stream) {
    SPAIR_REQUIRE(image && z_where && out && B > 0 && C > 0 && Ih > 1 && Iw > 1 && Gh > 0 && Gw > 0);
    SPAIR_REQUIRE(ld_out >= C * Gh * Gw && (!cells || n_cells > 0) && ((uintptr_t)z_where % 16) == 0);
    SPAIR_REQUIRE(!cells || n_cells <= 65535);
    if (resident_ok(image, cells, C, Ih, Iw, Gh, Gw, nullptr) && !getenv("SPAIR_GLIMPSE_PER_OBJECT")) {
        ResArgs a{image, z_where, cells, n_cells, B, HW, C, Ih, Iw, Gh, Gw, out, nullptr, nullptr, ld_out, 0, 1.0f / (float)Gw};
        return launch_resident<false>(a, (cudaStream_t)stream);
    }
    const dim3 grid(B, cells ? n_cells : 1);
    const size_t smem = glimpse_smem(Gh, Gw, false);
    SPAIR_REQUIRE(smem <= 200 * 1024);
    const int aligned = (Iw % 4 == 0) && ((uintptr_t)image % 16 == 0);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(glimpse_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    glimpse_fwd_kernel<<<grid, kGlimpseThreads, smem, (cudaStream_t)stream>>>(image, z_where, cells, B, HW, C, Ih, Iw,
                                                                              Gh, Gw, out, ld_out, aligned,
                                                                              1.0f / (float)Gw);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_glimpse_bwd(const float* image, const float* z_where, const int* cells, int n_cells, int B,
                                 int HW, int C, int Ih, int Iw, int Gh, int Gw, const float* d_out, int ld_out,
                                 float* d_z_where_local, float* d_image, void* stream) {
    SPAIR_REQUIRE(image && z_where && d_out && d_z_where_local && B > 0 && C > 0 && Ih > 1 && Iw > 1 && Gh > 0 && Gw > 0);
    SPAIR_REQUIRE(ld_out >= C * Gh * Gw && (!cells || n_cells > 0) && ((uintptr_t)z_where % 16) == 0);
    SPAIR_REQUIRE(!cells || n_cells <= 65535);
    if (resident_ok(image, cells, C, Ih, Iw, Gh, Gw, d_image) && !getenv("SPAIR_GLIMPSE_PER_OBJECT")) {
        ResArgs a{image, z_where, cells, n_cells, B, HW, C, Ih, Iw, Gh, Gw, nullptr, d_out, d_z_where_local, ld_out, 0, 1.0f / (float)Gw};
        return launch_resident<true>(a, (cudaStream_t)stream);
    }
    const dim3 grid(B, cells ? n_cells : 1);
    const size_t smem = glimpse_smem(Gh, Gw, true);
    SPAIR_REQUIRE(smem <= 200 * 1024);
    const int aligned = (Iw % 4 == 0) && ((uintptr_t)image % 16 == 0);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(glimpse_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    glimpse_bwd_kernel<<<grid, kGlimpseThreads, smem, (cudaStream_t)stream>>>(
        image, z_where, cells, B, HW, C, Ih, Iw, Gh, Gw, d_out, ld_out, d_z_where_local, d_image, aligned,
        1.0f / (float)Gw);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_paste_fwd(const float* image, const float* z_where, int n, int C, int Gh, int Gw, int Oh,
                               int Ow, float* out, void* stream) {
    SPAIR_REQUIRE(image && z_where && out && n > 0 && C > 0 && Gh > 0 && Gw > 0 && Oh > 0 && Ow > 0);
    SPAIR_REQUIRE(((uintptr_t)z_where % 16) == 0 && n <= 65535);
    dim3 grid(grid_for((long long)Oh * Ow, 256), n);
    paste_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(image, z_where, n, C, Gh, Gw, Oh, Ow, out);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_paste_bwd(const float* image, const float* z_where, int n, int C, int Gh, int Gw, int Oh,
                               int Ow, const float* d_out, float* d_image, float* d_z_where, void* stream) {
    SPAIR_REQUIRE(image && z_where && d_out && n > 0 && C > 0 && Gh > 0 && Gw > 0 && Oh > 0 && Ow > 0);
    SPAIR_REQUIRE(((uintptr_t)z_where % 16) == 0 && n <= 65535);
    if (d_z_where) {
        cudaError_t e = cudaMemsetAsync(d_z_where, 0, sizeof(float) * 4 * (size_t)n, (cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid(grid_for((long long)Oh * Ow, 256), n);
    paste_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(image, z_where, n, C, Gh, Gw, Oh, Ow, d_out, d_image,
                                                             d_z_where);
    SPAIR_LAUNCH_CHECK();
}
