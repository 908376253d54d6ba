// R (+E): fused decoder-renderer.
//
// Replaces SPAIR._render after the decoder MLP (reference models.py:481-540) and the BCE of
// SPAIR._build_loss (models.py:547).  The reference materialises a sampling grid [N,I,I,2] and a
// warped stack [N,C+2,I,I] (0.5 GB + 0.76 GB at B=32, 344 GB + 137 GB at config D) and then makes five
// full-size elementwise passes; here nothing per-object and canvas-sized ever exists in HBM.
//
// Forward  — gather formulation, one CTA per 32x16 canvas tile of one image:
//   1. bin: scan the image's HW boxes once, keep (in cell order) those whose footprint meets the tile;
//   2. per group of K kept objects: stage the sub-rectangle of texels the tile can touch into shared
//      memory, applying scale/bias + analytical sigmoid, *z_pres and max(alpha*z_depth, 0.01) ONCE per
//      texel (texels are stored as float4 records [colour.., alpha, importance] so a tap is one
//      LDS.128 for C<=2, two for C<=4);
//   3. every thread composites its pixels: num_c += a~*c~_c*(m~+1e-9), den += m~+1e-9.
//   Output: recon = clamp(num/den), den (for backward) and per-tile BCE partial sums.
// Backward — object formulation, one CTA per object (all gradients of an object are produced by one
//   CTA: no global atomics, bitwise reproducible):
//   0. prep (separate elementwise kernel): gs_c = g_c/S, gs_C = sum_c g_c*out_c/S per canvas pixel,
//      g_c = d_recon + bce_scale * dBCE/dp masked by the clamp;
//   1. pixels of the footprint, band by band: re-sample a~, c~, m~ and their spatial derivatives from
//      the staged texels, form the per-pixel gradients into a shared band buffer and
//      accumulate the box gradient;
//   2. texels gather their gradient from the band buffer through the transposed bilinear weights;
//   3. chain rule through importance / presence / sigmoid, coalesced store of d_logits, block
//      reduction for d_z_where, d_z_depth, d_z_pres.
#include "warp_math.cuh"

namespace spair {

constexpr int kRTileW = 32, kRThreads = 256;
constexpr int kRPPT = 4;                      // canvas rows per thread in the forward kernel
constexpr int kRTileH = (kRThreads / kRTileW) * kRPPT;   // 32
constexpr int kRMaxGroup = 8;
constexpr int kBandPix = 1536;   // pixels per backward band (a 39x39 footprint fits one band)
constexpr int kBandMaxW = 64;    // footprint columns per chunk
constexpr int kBandMaxH = 256;   // footprint rows per band

struct RenderArgs {
    const float* logits;
    const float* z_where;
    const float* z_depth;
    const float* z_pres;
    int B, HW, G, Ih, Iw;
    float obj_scale, alpha_scale, alpha_bias;
    int decoded;      // 1: `logits` holds texel records already through the sigmoids (spair_gemm3x, SPAIR_GEMM_EPI_TEXEL)
    float* recon;
    float* denom;
    const float* target;
    float* bce_partial;
    int group;        // objects staged per round
    int slot_f4;      // float4 records per slot (= G*G*NF4)
    int table_off;    // byte offset of the per-object coordinate tables in dynamic shared memory
};

template <int C>
struct Tex {
    static constexpr int NF4 = (C + 2 + 3) / 4;
    float v[NF4 * 4];
};

// 1/(exp(-x)+1) (modules.py:187) with the SFU exponential and reciprocal: |error| <= 2e-7 for |x| <= 10
// (exp relative error <= 2^-22 + 6e-8|x|), far inside the 1e-5 parity tolerance; saturates to exactly 0 / 1
// like the reference for |x| > ~17 / 88.
__device__ __forceinline__ float sigmoid_analytical(float x) { return __fdividef(1.0f, __expf(-x) + 1.0f); }

// decode one texel: colours, alpha (with presence), importance  (models.py:485-500)
// `decoded`: l already holds sigma(obj_scale * logit) per colour channel and 1 - sigma(alpha_scale * logit + alpha_bias) for the
// alpha channel (the complement keeps sigma'(x) = s (1 - s) accurate where the +5 bias saturates alpha), written by the
// decoder GEMM's epilogue; only the per-object factors remain.
template <int C>
__device__ __forceinline__ void decode_texel(const float* __restrict__ l, float obj_scale, float alpha_scale,
                                             float alpha_bias, float pres, float depth, int decoded, float* out) {
    float s;
    if (decoded) {
#pragma unroll
        for (int c = 0; c < C; ++c) out[c] = l[c];
        s = 1.0f - l[C];
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) out[c] = sigmoid_analytical(__fmul_rn(l[c], obj_scale));
        s = sigmoid_analytical(__fadd_rn(__fmul_rn(l[C], alpha_scale), alpha_bias));
    }
    const float a = s * pres;
    out[C] = a;
    out[C + 1] = fmaxf(__fmul_rn(a, depth), 0.01f);
}

template <int C>
__device__ __forceinline__ void load_logits(const float* __restrict__ p, float* l) {
    if (C == 1) {
        const float2 q = __ldg(reinterpret_cast<const float2*>(p));
        l[0] = q.x; l[1] = q.y;
    } else if (C == 3) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        l[0] = q.x; l[1] = q.y; l[2] = q.z; l[3] = q.w;
    } else {
#pragma unroll
        for (int c = 0; c <= C; ++c) l[c] = __ldg(p + c);
    }
}

// conservative pixel range (inclusive) in which an object can have non-zero bilinear weight
__device__ __forceinline__ void footprint(float t, float s, int I, int G, int& lo, int& hi) {
    const float centre = t * (float)I - 0.5f;
    const float half = fabsf(s) * (float)I * 0.5f * (1.0f + 1.0f / (float)G);
    lo = (int)floorf(centre - half) - 1;
    hi = (int)ceilf(centre + half) + 1;
}

template <int C>
__global__ void __launch_bounds__(kRThreads) render_fwd_kernel(RenderArgs p) {
    constexpr int NF4 = Tex<C>::NF4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* slots = reinterpret_cast<float4*>(smem_raw);
    float4* aff = slots + (size_t)p.group * p.slot_f4;                          // [HW] inverse affine of kept objects
    unsigned short* list = reinterpret_cast<unsigned short*>(aff + p.HW);       // [HW] kept object ids, cell order
    // per staged object: 32 column entries {xa, xb (float4 offsets of the two taps), wx0, wx1} and 32 row entries
    // {ya, yb, wy0, wy1} of this tile; xa / ya < 0 marks a column / row whose sample falls outside the texture
    float4* colT = reinterpret_cast<float4*>(smem_raw + p.table_off);            // [group][32]
    float4* rowT = colT + (size_t)p.group * kRTileW;                             // [group][32]
    __shared__ float bXs[kRTileW], bYs[kRTileH];
    __shared__ int warp_cnt[kRThreads / 32];
    __shared__ int list_len;
    __shared__ float red[kRThreads / 32];

    const int b = blockIdx.z;
    const int X0 = blockIdx.x * kRTileW, Y0 = blockIdx.y * kRTileH;
    const int X1 = min(X0 + kRTileW, p.Iw) - 1, Y1 = min(Y0 + kRTileH, p.Ih) - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = p.G;

    // ---- 1. bin the image's objects against this tile (ordered compaction keeps cell order) ----
    if (threadIdx.x == 0) list_len = 0;
    __syncthreads();
    for (int k0 = 0; k0 < p.HW; k0 += kRThreads) {
        const int k = k0 + threadIdx.x;
        bool hit = false;
        float4 zw = make_float4(0.f, 0.f, 1.f, 1.f);
        if (k < p.HW) {
            zw = __ldg(reinterpret_cast<const float4*>(p.z_where) + (size_t)b * p.HW + k);
            int xl, xh, yl, yh;
            footprint(zw.x, zw.z, p.Iw, G, xl, xh);
            footprint(zw.y, zw.w, p.Ih, G, yl, yh);
            hit = xl <= X1 && xh >= X0 && yl <= Y1 && yh >= Y0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        int off = list_len;
        for (int w = 0; w < warp; ++w) off += warp_cnt[w];
        if (hit) {
            const int pos = off + __popc(m & ((1u << lane) - 1u));
            const InvAffine A(zw.x, zw.y, zw.z, zw.w);
            list[pos] = (unsigned short)k;
            aff[pos] = make_float4(A.ax, A.cx, A.ay, A.cy);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < kRThreads / 32; ++w) tot += warp_cnt[w];
            list_len += tot;
        }
        __syncthreads();
    }
    const int n_list = list_len;

    // ---- per-thread pixels: column tx, rows ty + 8*h ----
    const int tx = threadIdx.x & (kRTileW - 1), ty = threadIdx.x / kRTileW;
    const int X = X0 + tx;
    const bool px_ok = X < p.Iw;
    bool ok[kRPPT];
#pragma unroll
    for (int h = 0; h < kRPPT; ++h) ok[h] = px_ok && (Y0 + ty + (kRThreads / kRTileW) * h) < p.Ih;
    if (threadIdx.x < kRTileW) bXs[threadIdx.x] = base_coord(min(X0 + (int)threadIdx.x, p.Iw - 1), p.Iw);
    else if (threadIdx.x < kRTileW + kRTileH) bYs[threadIdx.x - kRTileW] = base_coord(min(Y0 + (int)threadIdx.x - kRTileW, p.Ih - 1), p.Ih);
    const float bX0 = base_coord(X0, p.Iw), bX1 = base_coord(X1, p.Iw);
    const float bY0 = base_coord(Y0, p.Ih), bY1 = base_coord(Y1, p.Ih);
    const float hG = 0.5f * (float)G;
    __syncthreads();

    float num[kRPPT][C];
    float den[kRPPT];
    int ncov[kRPPT];
#pragma unroll
    for (int h = 0; h < kRPPT; ++h) {
        den[h] = 0.0f;
        ncov[h] = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) num[h][c] = 0.0f;
    }

    // ---- 2./3. stage K objects, composite, repeat ----
    for (int g0 = 0; g0 < n_list; g0 += p.group) {
        const int gcount = min(p.group, n_list - g0);
        for (int s = 0; s < gcount; ++s) {
            const int k = list[g0 + s];
            const float4 A = aff[g0 + s];
            const size_t n = (size_t)b * p.HW + k;
            // texel sub-rectangle reachable from this tile (coordinates are monotone in X / Y)
            const float ixa = unnormalize(affine_coord(bX0, A.x, A.y), hG), ixb = unnormalize(affine_coord(bX1, A.x, A.y), hG);
            const float iya = unnormalize(affine_coord(bY0, A.z, A.w), hG), iyb = unnormalize(affine_coord(bY1, A.z, A.w), hG);
            const int tx_lo = max(0, (int)floorf(fmaxf(fminf(ixa, ixb), -1.0f)));
            const int tx_hi = min(G - 1, (int)floorf(fminf(fmaxf(ixa, ixb), (float)G)) + 1);
            const int ty_lo = max(0, (int)floorf(fmaxf(fminf(iya, iyb), -1.0f)));
            const int ty_hi = min(G - 1, (int)floorf(fminf(fmaxf(iya, iyb), (float)G)) + 1);
            const int tw = tx_hi - tx_lo + 1, th = ty_hi - ty_lo + 1;
            if (tw <= 0 || th <= 0) continue;
            const float depth = __ldg(p.z_depth + n), pres = __ldg(p.z_pres + n);
            float4* slot = slots + (size_t)s * p.slot_f4;
            const float* base = p.logits + n * (size_t)(G * G * (C + 1));
            const float inv_tw = 1.0f / (float)tw;
            for (int t = threadIdx.x; t < tw * th; t += kRThreads) {
                const int ry = fast_div(t, inv_tw), rx = t - ry * tw;
                const int tex = (ty_lo + ry) * G + tx_lo + rx;
                float l[C + 1];
                load_logits<C>(base + (size_t)tex * (C + 1), l);
                Tex<C> o;
#pragma unroll
                for (int i = 0; i < NF4 * 4; ++i) o.v[i] = 0.0f;
                decode_texel<C>(l, p.obj_scale, p.alpha_scale, p.alpha_bias, pres, depth, p.decoded, o.v);
#pragma unroll
                for (int q = 0; q < NF4; ++q)
                    slot[(size_t)tex * NF4 + q] = make_float4(o.v[4 * q], o.v[4 * q + 1], o.v[4 * q + 2], o.v[4 * q + 3]);
            }
        }
        // coordinate tables of the group: one entry per (object, tile column) and (object, tile row), computed ONCE per CTA
        // (every thread of a column / row would otherwise redo the same affine + floor + clamp arithmetic per object)
        for (int idx = threadIdx.x; idx < gcount * (kRTileW + kRTileH); idx += kRThreads) {
            const int s = idx / (kRTileW + kRTileH), e = idx - s * (kRTileW + kRTileH);
            const float4 A = aff[g0 + s];
            const bool is_col = e < kRTileW;
            const int i = is_col ? e : e - kRTileW;
            const bool inside = is_col ? (X0 + i < p.Iw) : (Y0 + i < p.Ih);
            const float ic = unnormalize(is_col ? affine_coord(bXs[i], A.x, A.y) : affine_coord(bYs[i], A.z, A.w), hG);
            const float f0 = floorf(ic);
            const bool valid = inside && f0 >= -1.0f && f0 <= (float)(G - 1);
            const int c0 = (int)f0;
            const float w1 = (c0 + 1 <= G - 1) ? ic - f0 : 0.0f, w0 = (c0 >= 0) ? f0 + 1.0f - ic : 0.0f;
            const int ca = max(c0, 0), cb = min(c0 + 1, G - 1);
            const int unit = is_col ? NF4 : G * NF4;
            const float4 ent = make_float4(__int_as_float(valid ? ca * unit : -1), __int_as_float(cb * unit), w0, w1);
            if (is_col) colT[s * kRTileW + i] = ent;
            else rowT[s * kRTileH + i] = ent;
        }
        __syncthreads();
        for (int s = 0; s < gcount; ++s) {
            const float4* slot = slots + (size_t)s * p.slot_f4;
            const float4 cT = colT[s * kRTileW + tx];
            const int xa = __float_as_int(cT.x), xb = __float_as_int(cT.y);
            if (xa < 0) continue;
            const float wx0 = cT.z, wx1 = cT.w;
#pragma unroll
            for (int h = 0; h < kRPPT; ++h) {
                const float4 rT = rowT[s * kRTileH + ty + (kRThreads / kRTileW) * h];
                const int ya = __float_as_int(rT.x), yb = __float_as_int(rT.y);
                if (ya < 0) continue;
                const float wy0 = rT.z, wy1 = rT.w;
                const float nw = __fmul_rn(wx0, wy0), ne = __fmul_rn(wx1, wy0), sw = __fmul_rn(wx0, wy1), se = __fmul_rn(wx1, wy1);
                float acc[NF4 * 4];
#pragma unroll
                for (int q = 0; q < NF4; ++q) {
                    const float4 t00 = slot[ya + xa + q], t01 = slot[ya + xb + q];
                    const float4 t10 = slot[yb + xa + q], t11 = slot[yb + xb + q];
                    acc[4 * q + 0] = fmaf(t11.x, se, fmaf(t10.x, sw, fmaf(t01.x, ne, __fmul_rn(t00.x, nw))));
                    acc[4 * q + 1] = fmaf(t11.y, se, fmaf(t10.y, sw, fmaf(t01.y, ne, __fmul_rn(t00.y, nw))));
                    acc[4 * q + 2] = fmaf(t11.z, se, fmaf(t10.z, sw, fmaf(t01.z, ne, __fmul_rn(t00.z, nw))));
                    acc[4 * q + 3] = fmaf(t11.w, se, fmaf(t10.w, sw, fmaf(t01.w, ne, __fmul_rn(t00.w, nw))));
                }
                const float imp = acc[C + 1] + 1e-9f;                      // models.py:527
                const float a = acc[C];
#pragma unroll
                for (int c = 0; c < C; ++c) num[h][c] = fmaf(__fmul_rn(a, acc[c]), imp, num[h][c]);   // models.py:529,535
                den[h] += imp;
                ncov[h] += 1;
            }
        }
        __syncthreads();
    }

    // ---- epilogue: normalise, clamp, store, fused BCE ----
    float bce = 0.0f;
#pragma unroll
    for (int h = 0; h < kRPPT; ++h) {
        if (!ok[h]) continue;
        const int Y = Y0 + ty + (kRThreads / kRTileW) * h;
        const float S = den[h] + (float)(p.HW - ncov[h]) * 1e-9f;          // every object adds 1e-9 (models.py:527,532)
        const size_t pix = (size_t)Y * p.Iw + X;
        if (p.denom) p.denom[(size_t)b * p.Ih * p.Iw + pix] = S;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float r = fminf(fmaxf(num[h][c] / S, 0.0f), 1.0f);        // models.py:540
            const size_t o = ((size_t)b * C + c) * p.Ih * p.Iw + pix;
            p.recon[o] = r;
            if (p.target) {
                const float t = __ldg(p.target + o);
                bce += (t - 1.0f) * fmaxf(log1pf(-r), -100.0f) - t * fmaxf(logf(r), -100.0f);   // F.binary_cross_entropy
            }
        }
    }
    if (p.bce_partial) {
        bce = warp_sum(bce);
        if (lane == 0) red[warp] = bce;
        __syncthreads();
        if (warp == 0) {
            float s = lane < kRThreads / 32 ? red[lane] : 0.0f;
            s = warp_sum(s);
            if (lane == 0) p.bce_partial[((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
template <int C>
__global__ void render_bwd_prep_kernel(const float* __restrict__ recon, const float* __restrict__ denom,
                                       const float* __restrict__ d_recon, const float* __restrict__ target,
                                       const float* __restrict__ bce_scale, int B, int Ih, int Iw,
                                       float* __restrict__ gs) {
    const size_t npix = (size_t)Ih * Iw;
    const size_t total = (size_t)B * npix;
    const float scale = bce_scale ? bce_scale[0] : 1.0f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / npix, pix = i - b * npix;
        const float invS = 1.0f / denom[i];
        float q = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const size_t o = (b * C + c) * npix + pix;
            const float r = recon[o];
            float g = d_recon ? d_recon[o] : 0.0f;
            if (target) g += scale * (r - target[o]) / fmaxf((1.0f - r) * r, 1e-12f);   // binary_cross_entropy_backward
            if (r >= 1.0f) g = 0.0f;   // clamp(max=1) active (r == 1 exactly is treated as clamped)
            gs[(b * (C + 1) + c) * npix + pix] = g * invS;
            q += g * r;
        }
        gs[(b * (C + 1) + C) * npix + pix] = q * invS;
    }
}

struct RenderBwdArgs {
    const float* logits;
    const float* z_where;
    const float* z_depth;
    const float* z_pres;
    int B, HW, G, Ih, Iw;
    float obj_scale, alpha_scale, alpha_bias;
    int decoded;
    const float* gs;
    float* d_logits;
    float* d_z_where;
    float* d_z_depth;
    float* d_z_pres;
};

constexpr int kOutside = -1000000;   // marks a footprint column / row whose sample falls outside the texture

template <int C>
__global__ void __launch_bounds__(kRThreads, (C <= 2) ? 4 : 3) render_bwd_kernel(RenderBwdArgs p) {
    constexpr int NF4 = Tex<C>::NF4;
    constexpr int NCH = C + 2;
    constexpr int NPG = (NCH == 3) ? 4 : NCH;   // stride of a per-pixel gradient record (float4 for C == 1)
    constexpr int QMAX = 4;   // texels owned per thread (G*G <= 1024)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = p.G, GG = G * G, GP = G + 2;
    // decoded texels with a one-texel zero border: zeros padding needs no tap masks
    float4* tex = reinterpret_cast<float4*>(smem_raw);                 // [GP*GP][NF4]
    float* pixg = reinterpret_cast<float*>(tex + (size_t)GP * GP * NF4);   // [kBandPix][NPG] per-pixel gradients
    float* col_fx = pixg + kBandPix * NPG;                             // [kBandMaxW] fractional sample position
    float* col_bx = col_fx + kBandMaxW;                                // [kBandMaxW] normalised canvas coordinate
    int* col_x0 = reinterpret_cast<int*>(col_bx + kBandMaxW);          // [kBandMaxW] floor(sample position)
    float* row_fy = reinterpret_cast<float*>(col_x0 + kBandMaxW);      // [kBandMaxH]
    float* row_by = row_fy + kBandMaxH;
    int* row_y0 = reinterpret_cast<int*>(row_by + kBandMaxH);
    int* tcol_lo = row_y0 + kBandMaxH;                                 // [32] first/last footprint column feeding texel column j
    int* tcol_hi = tcol_lo + 32;
    int* trow_lo = tcol_hi + 32;
    int* trow_hi = trow_lo + 32;
    __shared__ float red[6 * (kRThreads / 32)];

    const size_t n = blockIdx.x;
    const int b = (int)(n / p.HW);
    const float4 zw = __ldg(reinterpret_cast<const float4*>(p.z_where) + n);
    const float depth = __ldg(p.z_depth + n), pres = __ldg(p.z_pres + n);
    const InvAffine A(zw.x, zw.y, zw.z, zw.w);
    const float hG = 0.5f * (float)G;
    const float inv_G = 1.0f / (float)G, inv_GP = 1.0f / (float)GP;
    const float* lbase = p.logits + n * (size_t)(GG * (C + 1));

    // ---- 0. stage decoded texels (zero border) ----
#pragma unroll 3
    for (int t = threadIdx.x; t < GP * GP; t += kRThreads) {
        const int yp = fast_div(t, inv_GP), xp = t - yp * GP;
        Tex<C> o;
#pragma unroll
        for (int i = 0; i < NF4 * 4; ++i) o.v[i] = 0.0f;
        if (yp >= 1 && yp <= G && xp >= 1 && xp <= G) {
            float l[C + 1];
            load_logits<C>(lbase + (size_t)((yp - 1) * G + xp - 1) * (C + 1), l);
            decode_texel<C>(l, p.obj_scale, p.alpha_scale, p.alpha_bias, pres, depth, p.decoded, o.v);
        }
#pragma unroll
        for (int q = 0; q < NF4; ++q) tex[(size_t)t * NF4 + q] = make_float4(o.v[4 * q], o.v[4 * q + 1], o.v[4 * q + 2], o.v[4 * q + 3]);
    }

    int Xlo, Xhi, Ylo, Yhi;
    footprint(zw.x, zw.z, p.Iw, G, Xlo, Xhi);
    footprint(zw.y, zw.w, p.Ih, G, Ylo, Yhi);
    Xlo = max(Xlo, 0); Ylo = max(Ylo, 0); Xhi = min(Xhi, p.Iw - 1); Yhi = min(Yhi, p.Ih - 1);

    float dT[QMAX][NCH];
#pragma unroll
    for (int q = 0; q < QMAX; ++q)
#pragma unroll
        for (int c = 0; c < NCH; ++c) dT[q][c] = 0.0f;
    float acc[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};   // sum dgx, sum dgy, sum dgx*bX, sum dgy*bY, d_depth, d_pres
    const size_t npix = (size_t)p.Ih * p.Iw;
    const float* gs_b = p.gs + (size_t)b * (C + 1) * npix;
    // texel coordinates of this thread's texels
    int t_i[QMAX], t_j[QMAX];
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
        const int t = threadIdx.x + q * kRThreads;
        t_i[q] = fast_div(t, inv_G);
        t_j[q] = t - t_i[q] * G;
    }
    __syncthreads();

    for (int Xc = Xlo; Xc <= Xhi; Xc += kBandMaxW) {
        const int cw = min(kBandMaxW, Xhi - Xc + 1);
        const float inv_cw = 1.0f / (float)cw;
        // balanced bands of at most min(kBandPix / cw, kBandMaxH) rows
        const int n_rows = Yhi - Ylo + 1;
        const int max_rows = min(fast_div(kBandPix, inv_cw), kBandMaxH);
        const int n_bands = fast_div(n_rows + max_rows - 1, 1.0f / (float)max_rows);
        const int rows_per_band = n_bands > 0 ? fast_div(n_rows + n_bands - 1, 1.0f / (float)n_bands) : 1;
        // ---- column tables of this chunk ----
        if (threadIdx.x < 32) { tcol_lo[threadIdx.x] = 1 << 30; tcol_hi[threadIdx.x] = -1; }
        __syncthreads();
        for (int j = threadIdx.x; j < cw; j += kRThreads) {
            const float bX = base_coord(Xc + j, p.Iw);
            const float ix = unnormalize(affine_coord(bX, A.ax, A.cx), hG);
            const float f0 = floorf(ix);
            const bool in = f0 >= -1.0f && f0 <= (float)(G - 1);
            const int x0 = in ? (int)f0 : kOutside;
            col_x0[j] = x0;
            col_fx[j] = ix - f0;
            col_bx[j] = bX;
            if (in) {
                if (x0 >= 0) { atomicMin(&tcol_lo[x0], j); atomicMax(&tcol_hi[x0], j); }
                if (x0 + 1 <= G - 1) { atomicMin(&tcol_lo[x0 + 1], j); atomicMax(&tcol_hi[x0 + 1], j); }
            }
        }
        for (int Yb = Ylo; Yb <= Yhi; Yb += rows_per_band) {
            const int rh = min(rows_per_band, Yhi - Yb + 1);
            if (threadIdx.x < 32) { trow_lo[threadIdx.x] = 1 << 30; trow_hi[threadIdx.x] = -1; }
            __syncthreads();
            for (int i = threadIdx.x; i < rh; i += kRThreads) {
                const float bY = base_coord(Yb + i, p.Ih);
                const float iy = unnormalize(affine_coord(bY, A.ay, A.cy), hG);
                const float f0 = floorf(iy);
                const bool in = f0 >= -1.0f && f0 <= (float)(G - 1);
                const int y0 = in ? (int)f0 : kOutside;
                row_y0[i] = y0;
                row_fy[i] = iy - f0;
                row_by[i] = bY;
                if (in) {
                    if (y0 >= 0) { atomicMin(&trow_lo[y0], i); atomicMax(&trow_hi[y0], i); }
                    if (y0 + 1 <= G - 1) { atomicMin(&trow_lo[y0 + 1], i); atomicMax(&trow_hi[y0 + 1], i); }
                }
            }
            __syncthreads();
            // ---- 1. per-pixel gradients ----
            for (int pp = threadIdx.x; pp < cw * rh; pp += kRThreads) {
                const int py = fast_div(pp, inv_cw), px = pp - py * cw;
                const int x0 = col_x0[px], y0 = row_y0[py];
                float out[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) out[c] = 0.0f;
                if (x0 != kOutside && y0 != kOutside) {
                    const size_t pix = (size_t)(Yb + py) * p.Iw + (Xc + px);
                    float gsv[C + 1];                                          // issued early: L2 latency overlaps the taps
#pragma unroll
                    for (int c = 0; c <= C; ++c) gsv[c] = __ldg(gs_b + (size_t)c * npix + pix);
                    const float wx1 = col_fx[px], wy1 = row_fy[py];
                    const float wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;       // == (x0+1) - ix, exactly
                    const float nw = __fmul_rn(wx0, wy0), ne = __fmul_rn(wx1, wy0), sw = __fmul_rn(wx0, wy1), se = __fmul_rn(wx1, wy1);
                    const float4* tp = tex + (size_t)((y0 + 1) * GP + x0 + 1) * NF4;   // padded coordinates
                    float v[NF4 * 4], dvx[NF4 * 4], dvy[NF4 * 4];
#pragma unroll
                    for (int q = 0; q < NF4; ++q) {
                        const float4 t00 = tp[q], t01 = tp[NF4 + q], t10 = tp[(size_t)GP * NF4 + q], t11 = tp[(size_t)(GP + 1) * NF4 + q];
                        const float a00[4] = {t00.x, t00.y, t00.z, t00.w}, a01[4] = {t01.x, t01.y, t01.z, t01.w};
                        const float a10[4] = {t10.x, t10.y, t10.z, t10.w}, a11[4] = {t11.x, t11.y, t11.z, t11.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            v[4 * q + e] = fmaf(a11[e], se, fmaf(a10[e], sw, fmaf(a01[e], ne, __fmul_rn(a00[e], nw))));
                            // grid_sampler_2d_backward, zeros padding (border texels are 0)
                            dvx[4 * q + e] = (a01[e] - a00[e]) * wy0 + (a11[e] - a10[e]) * wy1;
                            dvy[4 * q + e] = (a10[e] - a00[e]) * wx0 + (a11[e] - a01[e]) * wx1;
                        }
                    }
                    const float a = v[C], m = v[C + 1] + 1e-9f;
                    float T = 0.0f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        T = fmaf(gsv[c], v[c], T);
                        out[c] = gsv[c] * a * m;                             // dL/dc~
                    }
                    out[C] = m * T;                                          // dL/da~
                    out[C + 1] = a * T - gsv[C];                             // dL/dm~
                    float gix = 0.0f, giy = 0.0f;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        gix = fmaf(out[c], dvx[c], gix);
                        giy = fmaf(out[c], dvy[c], giy);
                    }
                    const float dgx = gix * hG, dgy = giy * hG;
                    acc[0] += dgx;
                    acc[1] += dgy;
                    acc[2] = fmaf(dgx, col_bx[px], acc[2]);
                    acc[3] = fmaf(dgy, row_by[py], acc[3]);
                }
                if (NPG == 4) {
                    *reinterpret_cast<float4*>(pixg + (size_t)pp * 4) = make_float4(out[0], out[1], out[2], 0.0f);
                } else {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) pixg[(size_t)pp * NPG + c] = out[c];
                }
            }
            __syncthreads();
            // ---- 2. texels gather through the transposed bilinear weights (exact ranges) ----
#pragma unroll
            for (int q = 0; q < QMAX; ++q) {
                if (threadIdx.x + q * kRThreads >= GG) break;
                const int ti = t_i[q], tj = t_j[q];
                const int pya = trow_lo[ti], pyb = trow_hi[ti], pxa = tcol_lo[tj], pxb = tcol_hi[tj];
                for (int py = pya; py <= pyb; ++py) {
                    const float fy = row_fy[py];
                    const float wy = (row_y0[py] == ti) ? 1.0f - fy : fy;
                    const float* grow = pixg + (size_t)py * cw * NPG;
                    for (int px = pxa; px <= pxb; ++px) {
                        const float fx = col_fx[px];
                        const float w = wy * ((col_x0[px] == tj) ? 1.0f - fx : fx);
                        if (NPG == 4) {
                            const float4 g4 = *reinterpret_cast<const float4*>(grow + px * 4);
                            dT[q][0] = fmaf(w, g4.x, dT[q][0]);
                            dT[q][1] = fmaf(w, g4.y, dT[q][1]);
                            dT[q][2] = fmaf(w, g4.z, dT[q][2]);
                        } else {
#pragma unroll
                            for (int c = 0; c < NCH; ++c) dT[q][c] = fmaf(w, grow[px * NPG + c], dT[q][c]);
                        }
                    }
                }
            }
            __syncthreads();
        }
    }

    // ---- 3. chain rule per texel, store d_logits ----
    float* dl_base = p.d_logits + n * (size_t)(GG * (C + 1));
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
        const int t = threadIdx.x + q * kRThreads;
        if (t >= GG) break;
        float l[C + 1];
        load_logits<C>(lbase + (size_t)t * (C + 1), l);
        float dl[C + 1];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float ds;   // sigma'(x) = e s^2 = s (1 - s)
            if (p.decoded) {
                ds = l[c] * (1.0f - l[c]);
            } else {
                const float e = __expf(-__fmul_rn(l[c], p.obj_scale));
                const float s = __fdividef(1.0f, e + 1.0f);
                ds = e * s * s;
            }
            dl[c] = dT[q][c] * ds * p.obj_scale;
        }
        float s, ds;
        if (p.decoded) {
            s = 1.0f - l[C];
            ds = s * l[C];
        } else {
            const float e = __expf(-__fadd_rn(__fmul_rn(l[C], p.alpha_scale), p.alpha_bias));
            s = __fdividef(1.0f, e + 1.0f);
            ds = e * s * s;
        }
        const float a = s * pres;
        float d_a = dT[q][C];
        if (__fmul_rn(a, depth) >= 0.01f) {                                  // clamp(min=0.01) passes the gradient (models.py:500)
            d_a = fmaf(dT[q][C + 1], depth, d_a);
            acc[4] = fmaf(dT[q][C + 1], a, acc[4]);
        }
        acc[5] = fmaf(d_a, s, acc[5]);                                       // alpha = sigmoid * z_pres (models.py:496)
        dl[C] = d_a * pres * ds * p.alpha_scale;
        if (C == 1) {
            *reinterpret_cast<float2*>(dl_base + (size_t)t * 2) = make_float2(dl[0], dl[1]);
        } else if (C == 3) {
            *reinterpret_cast<float4*>(dl_base + (size_t)t * 4) = make_float4(dl[0], dl[1], dl[2], dl[3]);
        } else {
#pragma unroll
            for (int c = 0; c <= C; ++c) dl_base[(size_t)t * (C + 1) + c] = dl[c];
        }
    }
    block_sum<6>(acc, red);
    if (threadIdx.x == 0) {
        const float xtp = 2.0f * zw.x - 1.0f, ytp = 2.0f * zw.y - 1.0f;
        // gx = bX * (1/xs) - (2 xt - 1)/xs   (theta after Tensor.inverse(), modules.py:256-262)
        p.d_z_where[n * 4 + 0] = -2.0f * acc[0] / zw.z;
        p.d_z_where[n * 4 + 1] = -2.0f * acc[1] / zw.w;
        p.d_z_where[n * 4 + 2] = (-acc[2] + acc[0] * xtp) / (zw.z * zw.z);
        p.d_z_where[n * 4 + 3] = (-acc[3] + acc[1] * ytp) / (zw.w * zw.w);
        p.d_z_depth[n] = acc[4];
        p.d_z_pres[n] = acc[5];
    }
}

template <int C>
static int launch_fwd(const RenderArgs& a, cudaStream_t st) {
    RenderArgs p = a;
    const int slot_f4 = p.G * p.G * Tex<C>::NF4;
    const size_t slot_bytes = (size_t)slot_f4 * sizeof(float4);
    int group = (int)((48 * 1024) / slot_bytes);
    if (group > kRMaxGroup) group = kRMaxGroup;
    if (group < 1) group = 1;
    p.group = group;
    p.slot_f4 = slot_f4;
    const size_t table_off = group * slot_bytes + (size_t)p.HW * sizeof(float4) +
                             (((size_t)p.HW * sizeof(unsigned short) + 15) & ~(size_t)15);
    p.table_off = (int)table_off;
    const size_t smem = table_off + (size_t)group * (kRTileW + kRTileH) * sizeof(float4);
    if (smem > 200 * 1024) return SPAIR_ERR_INVALID;
    static size_t smem_set[kMaxDevices] = {0};
    if (ensure_dynamic_smem(render_fwd_kernel<C>, smem, smem_set) != cudaSuccess) return (int)cudaGetLastError();
    dim3 grid((p.Iw + kRTileW - 1) / kRTileW, (p.Ih + kRTileH - 1) / kRTileH, p.B);
    render_fwd_kernel<C><<<grid, kRThreads, smem, st>>>(p);
    SPAIR_LAUNCH_CHECK();
}

template <int C>
static int launch_bwd(const RenderBwdArgs& p, const float* recon, const float* denom, const float* d_recon,
                      const float* target, const float* bce_scale, float* gs_ws, cudaStream_t st) {
    const size_t total = (size_t)p.B * p.Ih * p.Iw;
    int grid = grid_for((long long)total, 256);
    if (grid > kSMs * 16) grid = kSMs * 16;
    render_bwd_prep_kernel<C><<<grid, 256, 0, st>>>(recon, denom, d_recon, target, bce_scale, p.B, p.Ih, p.Iw, gs_ws);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const size_t smem = (size_t)(p.G + 2) * (p.G + 2) * Tex<C>::NF4 * sizeof(float4) +
                        sizeof(float) * ((size_t)kBandPix * ((C + 2) == 3 ? 4 : (C + 2)) + 3 * kBandMaxW + 3 * kBandMaxH + 4 * 32);
    if (smem > 200 * 1024) return SPAIR_ERR_INVALID;
    static size_t smem_set[kMaxDevices] = {0};
    if (ensure_dynamic_smem(render_bwd_kernel<C>, smem, smem_set) != cudaSuccess) return (int)cudaGetLastError();
    render_bwd_kernel<C><<<(unsigned)((size_t)p.B * p.HW), kRThreads, smem, st>>>(p);
    SPAIR_LAUNCH_CHECK();
}

}  // namespace spair

using namespace spair;

extern "C" int spair_render_num_tiles(int B, int Ih, int Iw) {
    return B * ((Ih + kRTileH - 1) / kRTileH) * ((Iw + kRTileW - 1) / kRTileW);
}

extern "C" int spair_render_fwd(const float* logits, const float* z_where, const float* z_depth, const float* z_pres,
                                int B, int HW, int C, int G, int Ih, int Iw, float obj_scale, float alpha_scale,
                                float alpha_bias, int decoded, float* recon, float* denom, const float* target,
                                float* bce_partial, void* stream) {
    SPAIR_REQUIRE(logits && z_where && z_depth && z_pres && recon);
    SPAIR_REQUIRE(B > 0 && B <= 65535 && HW > 0 && HW <= 65535 && G > 0 && G * G <= 1024 && Ih > 0 && Iw > 0);
    SPAIR_REQUIRE((target == nullptr) == (bce_partial == nullptr));
    SPAIR_REQUIRE(((uintptr_t)z_where % 16) == 0 && ((uintptr_t)logits % 16) == 0);
    RenderArgs a{logits, z_where, z_depth, z_pres, B, HW, G, Ih, Iw, obj_scale, alpha_scale, alpha_bias, decoded,
                 recon, denom, target, bce_partial, 0, 0, 0};
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
        case 1: return launch_fwd<1>(a, st);
        case 2: return launch_fwd<2>(a, st);
        case 3: return launch_fwd<3>(a, st);
        case 4: return launch_fwd<4>(a, st);
        default: return SPAIR_ERR_INVALID;
    }
}

extern "C" int spair_render_bwd(const float* logits, const float* z_where, const float* z_depth, const float* z_pres,
                                int B, int HW, int C, int G, int Ih, int Iw, float obj_scale, float alpha_scale,
                                float alpha_bias, int decoded, const float* recon, const float* denom, const float* d_recon,
                                const float* target, const float* bce_scale, float* gs_ws, float* d_logits,
                                float* d_z_where, float* d_z_depth, float* d_z_pres, void* stream) {
    SPAIR_REQUIRE(logits && z_where && z_depth && z_pres && recon && denom && gs_ws);
    SPAIR_REQUIRE(d_logits && d_z_where && d_z_depth && d_z_pres && (d_recon || target));
    SPAIR_REQUIRE(B > 0 && HW > 0 && G > 0 && G * G <= 1024 && Ih > 0 && Iw > 0);
    SPAIR_REQUIRE(((uintptr_t)z_where % 16) == 0 && ((uintptr_t)logits % 16) == 0 && ((uintptr_t)d_logits % 16) == 0);
    RenderBwdArgs a{logits, z_where, z_depth, z_pres, B, HW, G, Ih, Iw, obj_scale, alpha_scale, alpha_bias, decoded,
                    gs_ws, d_logits, d_z_where, d_z_depth, d_z_pres};
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
        case 1: return launch_bwd<1>(a, recon, denom, d_recon, target, bce_scale, gs_ws, st);
        case 2: return launch_bwd<2>(a, recon, denom, d_recon, target, bce_scale, gs_ws, st);
        case 3: return launch_bwd<3>(a, recon, denom, d_recon, target, bce_scale, gs_ws, st);
        case 4: return launch_bwd<4>(a, recon, denom, d_recon, target, bce_scale, gs_ws, st);
        default: return SPAIR_ERR_INVALID;
    }
}
