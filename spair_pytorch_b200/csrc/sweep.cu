// Fused forward sweep: the whole autoregressive cell loop (reference models.py:68-117) in ONE persistent launch.
//
// Every per-cell computation of an image depends only on that image (the lateral context reads cells of the SAME
// image), so the sweep is partitioned by image: a CTA owns `ipc` images and walks all Wc+2(Hc-1) wavefronts for
// them with block-level barriers only — no kernel boundary, no grid-wide synchronisation.  Per wavefront a CTA
// handles R = n_cells * ipc <= 16 rows through context gather -> box MLP -> box head -> glimpse -> encoder MLP ->
// attr head -> z MLP -> depth head -> obj MLP -> presence head.  The MLP layers are register-tiled SIMT dot
// products: a thread owns two output columns, up to 16 rows and a slice of the reduction index; it streams PACKED
// weights (four consecutive reduction indices of one column = one float4, spair_sweep_pack_weights) from L2 with
// coalesced 128-bit loads and reads the activations as 128-bit shared-memory broadcasts; each weight element is
// fetched once per CTA and wavefront and reused for all rows.  All activations are also written to the same
// wavefront-major global buffers the unfused path uses, so the backward pass (and the weight-gradient GEMMs) are
// unchanged.  This replaces ~27 launches per wavefront (12 cuBLAS GEMMs of <= 1536 rows, 8 elementwise, 6 head
// kernels) whose issue-to-issue latency, not their arithmetic, bounded the step.
#include <stdlib.h>

#include "warp_math.cuh"

namespace spair {

constexpr int kSwThreads = 512;     // 16 warps: the weight stream from L2 is latency bound, warps hide it
constexpr int kSwRows = 16;        // padded rows per CTA and wavefront
constexpr int kSwKC = 512;         // K-chunk of a layer whose input comes from global memory
constexpr int kSwKCP = kSwKC + 4;  // padded chunk row (keeps float4 alignment, staggers banks)
constexpr int kSwPart = 2 * kSwThreads * kSwRows;   // split-K partial sums: [K-split][row][column] = 16384 floats
constexpr int kSwHP = 256 + 4;     // padded hidden row
constexpr int kSwMaxG = 64;

#ifdef SW_TIMING
__device__ long long g_sw_timing[64];
#define SW_T0() long long sw_t_ = clock64()
#define SW_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long n_ = clock64(); g_sw_timing[i] += n_ - sw_t_; sw_t_ = n_; } } while (0)
#else
#define SW_T0()
#define SW_MARK(i)
#endif

}  // namespace spair
#include "sweep_tc.cuh"
namespace spair {

// block barrier of the sweep's 512 threads: the whole CTA in the SIMT kernels, named barrier 1 in the tensor-core kernels
// (their producer / MMA-issuer warps never join it)
template <bool TC>
__device__ __forceinline__ void sw_sync() {
    if (TC) tc::workers_sync();
    else __syncthreads();
}

struct SweepLayer {
    const float4* Wp;  // [ceil(K/4)][N] packed weight: Wp[g][n] = W[n][4g .. 4g+3], zero padded (tensor-core sweep: the
                       // UNPACKED nn.Linear weight [N][K], read only to re-evaluate ReLU pre-activations near zero)
    const float* b;    // [N]
    int K, N;
};

struct SweepMLP {
    SweepLayer l[3];
    float* X; int ldX;     // [HW*B, ldX] input rows (wavefront-major)
    float* H0; float* H1;  // [HW*B, N0], [HW*B, N1] post-ReLU activations
    float* Y;              // [HW*B, N2]
};

struct SweepFwdArgs {
    int B, HW, Hc, Wc, F, A, P, C, Ih, Iw, G, ipc, n_wavefronts;
    NeighbourList nb;
    const int* order;      // [HW] cells in wavefront-major order
    const int* starts;     // [n_wavefronts + 1]
    const float* image; const float* feat; const float* edge;
    const float* eps_where; const float* eps_attr; const float* eps_depth; const float* u_pres;
    spair_box_geom geom;
    SweepMLP box, enc, z, obj;
    float* out_box; float* z_where; float* attr; float* depth; float* pres; float* dmean; float* dstd;
};

// ---- split-K dense layer --------------------------------------------------------------------------------------
// The 512 threads are NCOLP / 2 column pairs x KS = 1024 / NCOLP slices of the reduction index.  A thread accumulates
// NR (<= 16) rows x 2 columns over its slice: a weight element is loaded ONCE per CTA and used for every row, an
// activation (128-bit shared-memory broadcast) feeds 8 FMAs, and the dependent load->FMA chain of a layer is KS
// times shorter than with one thread per column (the sweep is latency bound: the CTA is alone on its SM).  Partial
// sums meet in shared memory; the epilogue adds the bias, applies ReLU / the ReLU mask and stores to shared + global
// memory.

// acc[r] += w . x[r][0..3] for all rows and both columns of the thread: one 128-bit shared-memory broadcast feeds 8 FMAs
template <int NR>
__device__ __forceinline__ void fma_group2(float (&a)[NR], float (&b)[NR], const float* __restrict__ xp, int stride,
                                           const float4 wa, const float4 wb) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(xp + r * stride);
        a[r] = fmaf(wa.x, x.x, a[r]);
        b[r] = fmaf(wb.x, x.x, b[r]);
        a[r] = fmaf(wa.y, x.y, a[r]);
        b[r] = fmaf(wb.y, x.y, b[r]);
        a[r] = fmaf(wa.z, x.z, a[r]);
        b[r] = fmaf(wb.z, x.z, b[r]);
        a[r] = fmaf(wa.w, x.w, a[r]);
        b[r] = fmaf(wb.w, x.w, b[r]);
    }
}

// a[r] += sum over groups g in [0, ng) of P[g0 + g][colA] . xs[r][4g .. 4g+3] (b: colB)
template <int NR>
__device__ __forceinline__ void accumulate_packed2(float (&a)[NR], float (&b)[NR], const float4* __restrict__ P, int ncols,
                                                   int colA, int colB, int g0, int ng, const float* __restrict__ xs, int stride) {
    const float4* wa = P + (size_t)g0 * ncols + colA;
    const int dB = colB - colA;
    const float* xp = xs;
    int g = 0;
#pragma unroll 1
    for (; g + 4 <= ng; g += 4) {
        const float4 a0 = __ldg(wa), a1 = __ldg(wa + ncols), a2 = __ldg(wa + 2 * ncols), a3 = __ldg(wa + 3 * ncols);
        const float4 b0 = __ldg(wa + dB), b1 = __ldg(wa + ncols + dB), b2 = __ldg(wa + 2 * ncols + dB), b3 = __ldg(wa + 3 * ncols + dB);
        wa += 4 * ncols;
        fma_group2<NR>(a, b, xp, stride, a0, b0);
        fma_group2<NR>(a, b, xp + 4, stride, a1, b1);
        fma_group2<NR>(a, b, xp + 8, stride, a2, b2);
        fma_group2<NR>(a, b, xp + 12, stride, a3, b3);
        xp += 16;
    }
#pragma unroll 1
    for (; g < ng; ++g) {
        const float4 a0 = __ldg(wa), b0 = __ldg(wa + dB);
        wa += ncols;
        fma_group2<NR>(a, b, xp, stride, a0, b0);
        xp += 4;
    }
}

// Thread mapping of one pass over NCOLP output columns: thread -> column pair (n, n + NCOLP/2), K-slice ks of KS
template <int NCOLP>
struct SliceMap {
    static constexpr int kHalf = NCOLP / 2;
    static constexpr int KS = kSwThreads / kHalf;
    static_assert(KS * kSwRows * NCOLP <= kSwPart, "partial-sum buffer");
};

// this thread's slice of `ng` reduction groups (xs points at group 0, P row g0 is group 0); c0 = first column of the pass
template <int NR, int NCOLP>
__device__ __forceinline__ void accumulate_slice(float (&a)[NR], float (&b)[NR], const float4* __restrict__ P, int ncols, int c0,
                                                 int g0, int ng, const float* __restrict__ xs, int stride) {
    using M = SliceMap<NCOLP>;
    const int n = threadIdx.x % M::kHalf, ks = threadIdx.x / M::kHalf;
    const int colA = c0 + n;
    if (colA >= ncols) return;
    const int colB = (colA + M::kHalf < ncols) ? colA + M::kHalf : colA;     // no second column: recompute the first
    const int per = (ng + M::KS - 1) / M::KS;
    const int gb = ks * per, ge = min(ng, gb + per);
    if (gb < ge) accumulate_packed2<NR>(a, b, P, ncols, colA, colB, g0 + gb, ge - gb, xs + 4 * gb, stride);
}

template <int NR, int NCOLP>
__device__ __forceinline__ void store_partials(const float (&a)[NR], const float (&b)[NR], float* part) {
    using M = SliceMap<NCOLP>;
    const int n = threadIdx.x % M::kHalf, ks = threadIdx.x / M::kHalf;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        part[(ks * kSwRows + r) * NCOLP + n] = a[r];
        part[(ks * kSwRows + r) * NCOLP + n + M::kHalf] = b[r];
    }
}

// Epilogue of one pass over NCOLP columns.  A thread finalises column n = tid % NCOLP for the rows r0, r0 + RS, ...
// (r0 = tid / NCOLP, RS = 512 / NCOLP).  Its bias word and the ReLU-mask words of its rows are loaded BEFORE the
// accumulation (`preload`), so their L2 / HBM latency overlaps the weight stream instead of following the barrier.
template <int NR, int NCOLP>
struct Epilogue {
    static constexpr int RS = kSwThreads / NCOLP;          // row stride of a thread: 4 or 2
    static constexpr int NV = (NR + RS - 1) / RS;          // rows per thread
    float bias;
    float mask[NV];

    __device__ __forceinline__ void preload(const float* __restrict__ b, const float* __restrict__ Hmask, int ncols_total,
                                            int c0, int nrows, const int* __restrict__ grow) {
        const int col = c0 + threadIdx.x % NCOLP, r0 = threadIdx.x / NCOLP;
        const bool colv = col < ncols_total;
        bias = (b && colv) ? __ldg(b + col) : 0.0f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int r = r0 + j * RS;
            mask[j] = (Hmask && colv && r < nrows) ? Hmask[(size_t)grow[r] * ncols_total + col] : 1.0f;
        }
    }

    // out[r][c0 + n] = act(bias + sum over the K-slices) (* mask); rows in [nrows, NR) are written as zero to shared memory
    __device__ __forceinline__ void finalize(const float* part, int ncols_total, int c0, int nrows, bool relu,
                                             const int* __restrict__ grow, float* out_smem, float* __restrict__ out_glob,
                                             int ld_out) const {
        constexpr int KS = SliceMap<NCOLP>::KS;
        const int n = threadIdx.x % NCOLP, col = c0 + n, r0 = threadIdx.x / NCOLP;
        if (col >= ncols_total) return;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int r = r0 + j * RS;
            if (r >= NR) break;
            float v = bias;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) v += part[(ks * kSwRows + r) * NCOLP + n];
            if (relu) v = fmaxf(v, 0.0f);
            if (!(mask[j] > 0.0f)) v = 0.0f;
            if (r < nrows) out_glob[(size_t)grow[r] * ld_out + col] = v;
            else v = 0.0f;
            if (out_smem) out_smem[r * kSwHP + col] = v;
        }
    }
};

// forward layer: out = act(in . W^T + b).  in_smem != nullptr: activations in shared memory [r][kSwHP]; otherwise the
// input rows live in global memory (Xg) and are staged through `chunk` in slabs of kSwKC columns.
template <int NR, int NCOLP, bool GLOBAL_IN>
__device__ __noinline__ void dense_layer_t(const SweepLayer L, const float* in_smem, const float* __restrict__ Xg, int ldX,
                              const int* __restrict__ grow, int nrows, float* chunk, float* part, float* out_smem,
                              float* __restrict__ out_glob, bool relu) {
    SW_T0();
    Epilogue<NR, NCOLP> epi;
    epi.preload(L.b, nullptr, L.N, 0, nrows, grow);
    float a[NR], b[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) a[r] = b[r] = 0.0f;
    if (!GLOBAL_IN) {
        accumulate_slice<NR, NCOLP>(a, b, L.Wp, L.N, 0, 0, (L.K + 3) >> 2, in_smem, kSwHP);
    } else {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;       // 16 warps == kSwRows: one warp stages one row
        for (int k0 = 0; k0 < L.K; k0 += kSwKC) {
            const int kc = min(kSwKC, L.K - k0);
            const int kcp = (kc + 3) & ~3;
            __syncthreads();                 // previous slab fully consumed
            {
                const float* src = Xg + (size_t)grow[warp] * ldX + k0;
                float* dst = chunk + warp * kSwKCP;
                for (int kk = lane; kk < kcp; kk += 32) dst[kk] = (warp < nrows && kk < kc) ? src[kk] : 0.0f;
            }
            __syncthreads();
            accumulate_slice<NR, NCOLP>(a, b, L.Wp, L.N, 0, k0 >> 2, kcp >> 2, chunk, kSwKCP);
        }
    }
    SW_MARK(20);
    store_partials<NR, NCOLP>(a, b, part);
    __syncthreads();
    SW_MARK(21);
    epi.finalize(part, L.N, 0, nrows, relu, grow, out_smem, out_glob, L.N);
    __syncthreads();
    SW_MARK(22);
}

__device__ __forceinline__ void dense_layer(const SweepLayer& L, const float* in_smem, const float* Xg, int ldX,
                                            const int* grow, int nrows, float* chunk, float* part, float* out_smem,
                                            float* out_glob, bool relu) {
    // zero the padding columns the next layer's float4 reads may touch
    for (int idx = threadIdx.x; idx < kSwRows * 4; idx += kSwThreads) out_smem[(idx >> 2) * kSwHP + L.N + (idx & 3)] = 0.0f;
    const int q = (nrows + 3) >> 2;         // rows actually computed, in groups of 4
#define SW_DISPATCH(NRV, NCV)                                                                                              \
    do {                                                                                                                   \
        if (in_smem) dense_layer_t<NRV, NCV, false>(L, in_smem, Xg, ldX, grow, nrows, chunk, part, out_smem, out_glob, relu); \
        else dense_layer_t<NRV, NCV, true>(L, in_smem, Xg, ldX, grow, nrows, chunk, part, out_smem, out_glob, relu);         \
    } while (0)
    if (L.N <= 128) {
        if (q <= 1) SW_DISPATCH(4, 128);
        else if (q == 2) SW_DISPATCH(8, 128);
        else if (q == 3) SW_DISPATCH(12, 128);
        else SW_DISPATCH(16, 128);
    } else {
        if (q <= 1) SW_DISPATCH(4, 256);
        else if (q == 2) SW_DISPATCH(8, 256);
        else if (q == 3) SW_DISPATCH(12, 256);
        else SW_DISPATCH(16, 256);
    }
#undef SW_DISPATCH
}

// three-layer MLP: X (global) -> H0 -> H1 -> Y ; result left in shared memory `y`
__device__ __forceinline__ void mlp3(const SweepMLP& M, const int* grow, int nrows, float* chunk, float* part, float* ha,
                                     float* hb, float* y) {
    SW_T0();
    dense_layer(M.l[0], nullptr, M.X, M.ldX, grow, nrows, chunk, part, ha, M.H0, true);
    SW_MARK(10);
    dense_layer(M.l[1], ha, nullptr, 0, grow, nrows, chunk, part, hb, M.H1, true);
    SW_MARK(11);
    dense_layer(M.l[2], hb, nullptr, 0, grow, nrows, chunk, part, y, M.Y, false);
    SW_MARK(12);
}

// the same MLP on the tensor cores: the input rows go from global memory straight into operand tiles, the hidden
// activations from the epilogue registers into the next layer's operand tiles (sweep_tc.cuh)
__device__ __forceinline__ void mlp3_tc(tc::Worker& W, const SweepMLP& M, const int* grow, int nrows, float* y) {
    SW_T0();
    tc::workers_sync();                                 // the input rows written by other threads are visible
    const int row = threadIdx.x >> 5;
    tc::stage_rows(W, M.X + (size_t)grow[row] * M.ldX, row < nrows, M.l[0].K);
    SW_MARK(10);
    const tc::Kink k0{reinterpret_cast<const float*>(M.l[0].Wp), M.X, M.ldX};
    const tc::Kink k1{reinterpret_cast<const float*>(M.l[1].Wp), M.H0, M.l[0].N};
    const tc::Kink none{nullptr, nullptr, 0};
    tc::epilogue(W, M.l[0].K, M.l[0].N, M.l[0].b, true, nullptr, grow, nrows, M.H0, M.l[0].N, true, nullptr, 0, k0, 20);
    SW_MARK(11);
    tc::epilogue(W, M.l[1].K, M.l[1].N, M.l[1].b, true, nullptr, grow, nrows, M.H1, M.l[1].N, true, nullptr, 0, k1, 20);
    SW_MARK(13);
    tc::epilogue(W, M.l[2].K, M.l[2].N, M.l[2].b, false, nullptr, grow, nrows, M.Y, M.l[2].N, false, y, kSwHP, none, 20);
    tc::workers_sync();
    SW_MARK(12);
}

__device__ __forceinline__ int sw_box_slot(int k) { return k == 0 ? 1 : (k == 1 ? 0 : (k == 2 ? 3 : 2)); }

// TC = false: SIMT dense layers (512 threads).  TC = true: tensor-core dense layers (sweep_tc.cuh; 576 threads: the same
// 16 worker warps + a bulk-copy producer warp + an MMA-issuer warp); everything between the MLPs is the same code.
template <bool TC>
__global__ void __launch_bounds__(TC ? tc::kThreads : kSwThreads, 1)
sweep_fwd_kernel(const SweepFwdArgs p, const tc::Plan plan, const float* __restrict__ wstream, int stages_per_wavefront) {
    static_assert(kSwThreads / 32 == kSwRows, "one warp stages one input row");
    static_assert(kSwThreads == tc::kWorkers && kSwRows == tc::kRows, "worker layout of the tensor-core path");
    extern __shared__ __align__(16) float sm[];
    float *chunk = nullptr, *part = nullptr, *ha = nullptr, *hb = nullptr, *y;
    tc::Worker W;
    uint32_t wring = 0;
    if (TC) {
        // [weight ring | activation ring] (1024-byte aligned for the 128-byte swizzle) | barriers | row-major buffers
        const uint32_t raw = tc::smem_u32(sm), base = (raw + 1023u) & ~1023u;
        wring = base;
        W.xring = base + tc::kWStages * tc::kStageBytes;
        W.b.base = base + tc::kRingBytes;
        W.x_prod = 0; W.acc_cnt = 0;
        y = reinterpret_cast<float*>(reinterpret_cast<char*>(sm) + (base - raw) + tc::kRingBytes + tc::kBarBytes);
    } else {
        chunk = sm;                                     // [kSwRows][kSwKCP]
        part = chunk + kSwRows * kSwKCP;                // [kSwPart] split-K partial sums
        ha = part + kSwPart;                            // [kSwRows][kSwHP]
        hb = ha + kSwRows * kSwHP;
        y = hb + kSwRows * kSwHP;
    }
    float* base_g = y + kSwRows * kSwHP;                // [kSwMaxG] normalised base grid of the glimpse
    float* zw_s = base_g + kSwMaxG;                     // [kSwRows][4] boxes of the current rows
    int* grow = reinterpret_cast<int*>(zw_s + kSwRows * 4);   // [kSwRows] global row of each local row
    int* rcell = grow + kSwRows;                        // [kSwRows] cell id
    int* rimg = rcell + kSwRows;                        // [kSwRows] image

    const int b0 = blockIdx.x * p.ipc;
    const int n_img = min(p.ipc, p.B - b0);
    if (n_img <= 0) return;
    if (TC) {
        const int warp = threadIdx.x >> 5;
        if (threadIdx.x == 0) tc::init_barriers(W.b);
        if (warp == tc::kMmaWarp) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(W.b.tmem_slot()), "n"(tc::kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        tc::fence_before();
        __syncthreads();
        tc::fence_after();
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(W.tmem) : "r"(W.b.tmem_slot()));
        if (warp >= tc::kProducerWarp) {
            if (threadIdx.x == tc::kProducerWarp * 32) tc::producer_loop(wring, W.b, wstream, stages_per_wavefront, p.n_wavefronts);
            // (an exact thread index: the compiler then keeps the descriptors in uniform registers; with a per-warp lane test
            // it wraps every tcgen05.mma in a broadcast loop and the issue rate drops by a third)
            if (threadIdx.x == tc::kMmaWarp * 32) tc::mma_loop<0>(wring, W.xring, W.b, W.tmem, plan, p.n_wavefronts);
            if (threadIdx.x == (tc::kMmaWarp + 1) * 32) tc::mma_loop<1>(wring, W.xring, W.b, W.tmem, plan, p.n_wavefronts);
            __syncwarp();
            tc::fence_before();
            __syncthreads();                            // the workers are done: every MMA has been consumed
            if (warp == tc::kMmaWarp) {
                __syncwarp();
                tc::fence_after();
                asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(W.tmem), "n"(tc::kTmemCols));
            }
            return;
        }
    }
    const int E = p.A + 6, D = 4 + p.A + 1;
    const int CTX = p.nb.n * E;
    const int c_pt = p.F + CTX, c_box = c_pt + p.P, c_attr = c_box + 4, c_depth = c_attr + p.A;
    const int GG = p.G * p.G;
    for (int j = threadIdx.x; j < p.G; j += kSwThreads) base_g[j] = base_coord(j, p.G);

    SW_T0();
    for (int t = 0; t < p.n_wavefronts; ++t) {
        const int s0 = p.starts[t], n_cells = p.starts[t + 1] - s0;
        const int nrows = n_cells * n_img;
        sw_sync<TC>();                                // previous wavefront's latents are visible; row tables free
        if (threadIdx.x < kSwRows) {
            const int r = threadIdx.x;
            if (r < nrows) {
                const int k = r / n_img, li = r - k * n_img;
                grow[r] = (s0 + k) * p.B + b0 + li;
                rcell[r] = p.order[s0 + k];
                rimg[r] = b0 + li;
            } else {
                grow[r] = 0; rcell[r] = 0; rimg[r] = b0;
            }
        }
        sw_sync<TC>();

        // ---- L0: lateral context -> input columns [0, F+CTX) of the three networks (models.py:73,76) ----
        const int width = p.F + CTX;
#pragma unroll 4
        for (int idx = threadIdx.x; idx < nrows * width; idx += kSwThreads) {
            const int r = idx / width, col = idx - r * width;
            const int b = rimg[r], cell = rcell[r];
            const int h = cell / p.Wc, w = cell - h * p.Wc;
            float v;
            if (col < p.F) {
                v = __ldg(p.feat + ((size_t)(b * p.F + col) * p.Hc + h) * p.Wc + w);
            } else {
                const int s = (col - p.F) / E, j = (col - p.F) - s * E;
                const int nh = h + p.nb.dh[s], nw = w + p.nb.dw[s];
                if (nh >= 0 && nh < p.Hc && nw >= 0 && nw < p.Wc) {
                    const size_t o = (size_t)b * p.HW + nh * p.Wc + nw;   // written by this CTA in an earlier wavefront
                    if (j < 4) v = p.out_box[o * 4 + j];
                    else if (j < 4 + p.A) v = p.attr[o * p.A + (j - 4)];
                    else if (j == 4 + p.A) v = p.depth[o];
                    else v = p.pres[o];
                } else {
                    v = __ldg(p.edge + j);
                }
            }
            const size_t g = grow[r];
            p.box.X[g * p.box.ldX + col] = v;
            p.z.X[g * p.z.ldX + col] = v;
            p.obj.X[g * p.obj.ldX + col] = v;
        }
        // (dense_layer starts with a barrier before it reads the rows back)
        SW_MARK(0);

        // ---- z_where: box network + box head (models.py:76-79, 322-381) ----
        if (TC) mlp3_tc(W, p.box, grow, nrows, y);
        else mlp3(p.box, grow, nrows, chunk, part, ha, hb, y);
        SW_MARK(1);
        for (int idx = threadIdx.x; idx < nrows * (4 + p.P); idx += kSwThreads) {
            const int r = idx / (4 + p.P), k = idx - r * (4 + p.P);
            const size_t g = grow[r];
            const float* yr = y + r * kSwHP;
            if (k >= 4) {
                p.z.X[g * p.z.ldX + c_pt + (k - 4)] = yr[8 + (k - 4)];
                continue;
            }
            const int cell = rcell[r];
            const size_t o = (size_t)rimg[r] * p.HW + cell;
            const float mean = yr[k];
            const float std_ = sigmoid_f(clamp10(yr[4 + k])) * 2.0f;
            const float zl = mean + __ldg(p.eps_where + o * 4 + k) * std_;
            const float sg = sigmoid_f(clamp10(zl));
            float bval, zval;
            if (k < 2) {
                bval = p.geom.yx_scale * sg + p.geom.yx_min;
                const float pos = (k == 0) ? (float)(cell / p.Wc) : (float)(cell % p.Wc);
                zval = (k == 0 ? p.geom.cell_ratio_y : p.geom.cell_ratio_x) * (bval + pos);
            } else {
                bval = p.geom.hw_scale * sg + p.geom.hw_min;
                zval = bval * p.geom.anchor / (k == 2 ? p.geom.img_h : p.geom.img_w);
            }
            const int slot = sw_box_slot(k);
            p.out_box[o * 4 + slot] = bval;
            p.z_where[o * 4 + slot] = zval;
            zw_s[r * 4 + slot] = zval;
            p.dmean[o * D + k] = mean;
            p.dstd[o * D + k] = std_;
            p.z.X[g * p.z.ldX + c_box + slot] = bval;
            p.obj.X[g * p.obj.ldX + c_box + slot] = bval;
        }
        sw_sync<TC>();
        SW_MARK(2);

        // ---- z_what: glimpse (modules.py:216-273, border padding) -> encoder input rows ----
#pragma unroll 4
        for (int idx = threadIdx.x; idx < nrows * GG; idx += kSwThreads) {
            const int r = idx / GG, tt = idx - r * GG;
            const int i = tt / p.G, j = tt - i * p.G;
            const FwdAffine A(zw_s[r * 4 + 0], zw_s[r * 4 + 1], zw_s[r * 4 + 2], zw_s[r * 4 + 3]);
            float ix = unnormalize(affine_coord(base_g[j], A.ax, A.cx), 0.5f * (float)p.Iw);
            float iy = unnormalize(affine_coord(base_g[i], A.ay, A.cy), 0.5f * (float)p.Ih);
            ix = fminf(fmaxf(ix, 0.0f), (float)(p.Iw - 1));
            iy = fminf(fmaxf(iy, 0.0f), (float)(p.Ih - 1));
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            const int x0 = (int)fx0, y0 = (int)fy0;
            const int x1 = min(x0 + 1, p.Iw - 1), y1 = min(y0 + 1, p.Ih - 1);     // weight is exactly 0 when clamped
            const float wx1 = ix - fx0, wx0 = fx0 + 1.0f - ix, wy1 = iy - fy0, wy0 = fy0 + 1.0f - iy;
            const float nw = __fmul_rn(wx0, wy0), ne = __fmul_rn(wx1, wy0), sw = __fmul_rn(wx0, wy1), se = __fmul_rn(wx1, wy1);
            float* orow = p.enc.X + (size_t)grow[r] * p.enc.ldX;
            for (int c = 0; c < p.C; ++c) {
                const float* pl = p.image + ((size_t)rimg[r] * p.C + c) * p.Ih * p.Iw;
                float acc = __fmul_rn(__ldg(pl + (size_t)y0 * p.Iw + x0), nw);
                acc = fmaf(__ldg(pl + (size_t)y0 * p.Iw + x1), ne, acc);
                acc = fmaf(__ldg(pl + (size_t)y1 * p.Iw + x0), sw, acc);
                acc = fmaf(__ldg(pl + (size_t)y1 * p.Iw + x1), se, acc);
                orow[c * GG + tt] = acc;
            }
        }
        SW_MARK(3);
        if (TC) mlp3_tc(W, p.enc, grow, nrows, y);
        else mlp3(p.enc, grow, nrows, chunk, part, ha, hb, y);
        SW_MARK(4);
        for (int idx = threadIdx.x; idx < nrows * p.A; idx += kSwThreads) {
            const int r = idx / p.A, k = idx - r * p.A;
            const size_t g = grow[r];
            const size_t o = (size_t)rimg[r] * p.HW + rcell[r];
            const float* yr = y + r * kSwHP;
            const float mean = yr[k];
            const float std_ = sigmoid_f(clamp10(yr[p.A + k])) * 2.0f;
            const float zv = mean + __ldg(p.eps_attr + o * p.A + k) * std_;
            p.attr[o * p.A + k] = zv;
            p.dmean[o * D + 4 + k] = mean;
            p.dstd[o * D + 4 + k] = std_;
            p.z.X[g * p.z.ldX + c_attr + k] = zv;
            p.obj.X[g * p.obj.ldX + c_attr + k] = zv;
        }

        // ---- z_depth (models.py:88-97) ----
        SW_MARK(5);
        if (TC) mlp3_tc(W, p.z, grow, nrows, y);
        else mlp3(p.z, grow, nrows, chunk, part, ha, hb, y);
        SW_MARK(6);
        for (int idx = threadIdx.x; idx < nrows * (1 + p.P); idx += kSwThreads) {
            const int r = idx / (1 + p.P), k = idx - r * (1 + p.P);
            const size_t g = grow[r];
            const float* yr = y + r * kSwHP;
            if (k >= 1) {
                p.obj.X[g * p.obj.ldX + c_pt + (k - 1)] = yr[2 + (k - 1)];
                continue;
            }
            const size_t o = (size_t)rimg[r] * p.HW + rcell[r];
            const float mean = yr[0];
            const float std_ = sigmoid_f(clamp10(yr[1])) * 2.0f;
            const float zv = 4.0f * sigmoid_f(clamp10(mean + __ldg(p.eps_depth + o) * std_));
            p.depth[o] = zv;
            p.dmean[o * D + D - 1] = mean;
            p.dstd[o * D + D - 1] = std_;
            p.obj.X[g * p.obj.ldX + c_depth] = zv;
        }

        // ---- z_pres (models.py:100-105, 393-411) ----
        SW_MARK(7);
        if (TC) mlp3_tc(W, p.obj, grow, nrows, y);
        else mlp3(p.obj, grow, nrows, chunk, part, ha, hb, y);
        SW_MARK(8);
        if (threadIdx.x < nrows) {
            const int r = threadIdx.x;
            const size_t o = (size_t)rimg[r] * p.HW + rcell[r];
            const float u = __ldg(p.u_pres + o);
            const float noise = logf(u + 10e-10f) - logf(1.0f - u + 10e-10f);
            p.pres[o] = sigmoid_f(clamp10(y[r * kSwHP]) + noise);
        }
    }
    if (TC) {
        tc::fence_before();
        __syncthreads();                                // pairs with the producer / MMA warps' closing barrier
    }
}


// =================================================================================================================
// Fused backward sweep: the hand-written backward of the cell loop (ops.CellSweepFunction.backward) in one launch.
// Same image partition as the forward: a CTA walks the wavefronts of its images in REVERSE order.  The gradient that
// arrives through the lateral context comes from cells of later wavefronts of the SAME image, i.e. from rows this
// CTA has already written.  Per wavefront: context-gradient gather -> presence head -> obj MLP (dX chain) -> depth
// head -> z MLP -> attr head -> encoder MLP -> glimpse (d z_where) -> box head -> box MLP.  A backward layer
// dX = dY . W is the same register-tiled dot product as the forward with the weight packed along n (reduction
// over n, coalesced over k) and a ReLU-mask epilogue.  dY / dH / dX of every row are written to the wavefront-major
// global buffers: the 13 weight gradients stay one large cuBLAS GEMM each over all rows, after the sweep.
// =================================================================================================================
struct SweepMLPBwd {
    const float4* W[3];    // [ceil(N/4)][K] packed weights: W[g][k] = (w[4g][k] .. w[4g+3][k]), zero padded
    int K[3], N[3];
    const float* H0; const float* H1; const float* Y;   // forward activations
    float* dX; int ldX; float* dH0; float* dH1; float* dY;
};

struct SweepBwdArgs {
    int B, HW, Hc, Wc, F, A, P, C, Ih, Iw, G, ipc, n_wavefronts;
    NeighbourList nb;
    const int* order; const int* starts; const int* wf_pos;
    const float* image; const float* z_where;
    const float* eps_where; const float* eps_attr; const float* eps_depth; const float* u_pres;
    const float* wheel;
    spair_box_geom geom;
    SweepMLPBwd box, enc, z, obj;
    const float* encX; int ld_encX;      // unused (glimpse is re-sampled), kept for symmetry
    const float* d_zw; const float* d_attr; const float* d_depth; const float* d_pres;   // image-major upstream (may be null)
    const float* d_dmean; const float* d_dstd;
};

// out[r][c] = sum_n g[r][n] * W[n][c] (* (Hmask[r][c] > 0)), output columns in passes of 128, reduction split 8 ways
template <int NR, int NCOLP>
__device__ __noinline__ void dense_bwd_layer_t(const float4* __restrict__ W, int Nred, int Kout, const float* g_smem,
                                  const float* __restrict__ Hmask, const int* __restrict__ grow, int nrows, float* part,
                                  float* out_smem, float* __restrict__ out_glob, int ld_out) {
    const int ng = (Nred + 3) >> 2;
    for (int c0 = 0; c0 < Kout; c0 += NCOLP) {
        SW_T0();
        Epilogue<NR, NCOLP> epi;
        epi.preload(nullptr, Hmask, Kout, c0, nrows, grow);
        float a[NR], b[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) a[r] = b[r] = 0.0f;
        accumulate_slice<NR, NCOLP>(a, b, W, Kout, c0, 0, ng, g_smem, kSwHP);
        SW_MARK(40);
        store_partials<NR, NCOLP>(a, b, part);
        __syncthreads();
        SW_MARK(41);
        epi.finalize(part, Kout, c0, nrows, false, grow, out_smem, out_glob, ld_out);
        __syncthreads();
        SW_MARK(42);
    }
}

__device__ __forceinline__ void dense_bwd_layer(const float4* W, int Nred, int Kout, const float* g_smem, const float* Hmask,
                                                const int* grow, int nrows, float* part, float* out_smem, float* out_glob,
                                                int ld_out) {
    if (out_smem)
        for (int idx = threadIdx.x; idx < kSwRows * 4; idx += kSwThreads) out_smem[(idx >> 2) * kSwHP + Kout + (idx & 3)] = 0.0f;
    const int q = (nrows + 3) >> 2;
#ifndef SW_BWD_WIDE
#define SW_BWD_WIDE 1
#endif
    if (SW_BWD_WIDE && Kout > 128) {       // wide outputs: 256 columns per pass, reduction split 4 ways
        if (q <= 1) dense_bwd_layer_t<4, 256>(W, Nred, Kout, g_smem, Hmask, grow, nrows, part, out_smem, out_glob, ld_out);
        else if (q == 2) dense_bwd_layer_t<8, 256>(W, Nred, Kout, g_smem, Hmask, grow, nrows, part, out_smem, out_glob, ld_out);
        else if (q == 3) dense_bwd_layer_t<12, 256>(W, Nred, Kout, g_smem, Hmask, grow, nrows, part, out_smem, out_glob, ld_out);
        else dense_bwd_layer_t<16, 256>(W, Nred, Kout, g_smem, Hmask, grow, nrows, part, out_smem, out_glob, ld_out);
    } else {                                // 128 columns per pass, reduction split 8 ways
        if (q <= 1) dense_bwd_layer_t<4, 128>(W, Nred, Kout, g_smem, Hmask, grow, nrows, part, out_smem, out_glob, ld_out);
        else if (q == 2) dense_bwd_layer_t<8, 128>(W, Nred, Kout, g_smem, Hmask, grow, nrows, part, out_smem, out_glob, ld_out);
        else if (q == 3) dense_bwd_layer_t<12, 128>(W, Nred, Kout, g_smem, Hmask, grow, nrows, part, out_smem, out_glob, ld_out);
        else dense_bwd_layer_t<16, 128>(W, Nred, Kout, g_smem, Hmask, grow, nrows, part, out_smem, out_glob, ld_out);
    }
}

// dY (already in shared memory `gy` and in global M.dY) -> dH1 -> dH0 -> dX
__device__ __forceinline__ void mlp3_bwd(const SweepMLPBwd& M, const int* grow, int nrows, float* part, float* gy, float* ga,
                                         float* gb) {
    dense_bwd_layer(M.W[2], M.N[2], M.K[2], gy, M.H1, grow, nrows, part, ga, M.dH1, M.K[2]);
    dense_bwd_layer(M.W[1], M.N[1], M.K[1], ga, M.H0, grow, nrows, part, gb, M.dH0, M.K[1]);
    dense_bwd_layer(M.W[0], M.N[0], M.K[0], gb, nullptr, grow, nrows, part, nullptr, M.dX, M.ldX);
}

// the same chain on the tensor cores (sweep_tc.cuh): gy (row-major, written by the head phase) -> operand tiles -> dH1 -> dH0 -> dX
__device__ __forceinline__ void mlp3_bwd_tc(tc::Worker& W, const SweepMLPBwd& M, const int* grow, int nrows, const float* gy) {
    SW_T0();
    const int row = threadIdx.x >> 5;
    tc::stage_rows(W, gy + row * kSwHP, row < nrows, M.N[2]);
    SW_MARK(43);
    const tc::Kink none{nullptr, nullptr, 0};
    tc::epilogue(W, M.N[2], M.K[2], nullptr, false, M.H1, grow, nrows, M.dH1, M.K[2], true, nullptr, 0, none, 40);
    SW_MARK(44);
    tc::epilogue(W, M.N[1], M.K[1], nullptr, false, M.H0, grow, nrows, M.dH0, M.K[1], true, nullptr, 0, none, 40);
    SW_MARK(45);
    tc::epilogue(W, M.N[0], M.K[0], nullptr, false, nullptr, grow, nrows, M.dX, M.ldX, false, nullptr, 0, none, 40);
    tc::workers_sync();                                 // dX rows are read back from global memory by the head phases
    SW_MARK(46);
}

template <bool TC>
__global__ void __launch_bounds__(TC ? tc::kThreads : kSwThreads, 1)
sweep_bwd_kernel(const SweepBwdArgs p, const tc::Plan plan, const float* __restrict__ wstream, int stages_per_wavefront) {
    static_assert(kSwThreads / 32 == kSwRows, "the glimpse gradient uses one warp per row");
    extern __shared__ __align__(16) float sm[];
    float *part = nullptr, *gy, *ga = nullptr, *gb = nullptr, *dcell;
    tc::Worker W;
    uint32_t wring = 0;
    if (TC) {
        const uint32_t raw = tc::smem_u32(sm), base = (raw + 1023u) & ~1023u;
        wring = base;
        W.xring = base + tc::kWStages * tc::kStageBytes;
        W.b.base = base + tc::kRingBytes;
        W.x_prod = 0; W.acc_cnt = 0;
        gy = reinterpret_cast<float*>(reinterpret_cast<char*>(sm) + (base - raw) + tc::kRingBytes + tc::kBarBytes);
        dcell = gy + kSwRows * kSwHP;
    } else {
        part = sm;                                      // [kSwPart] split-K partial sums
        gy = part + kSwPart;                            // [kSwRows][kSwHP] dY of the current network
        ga = gy + kSwRows * kSwHP;
        gb = ga + kSwRows * kSwHP;
        dcell = gb + kSwRows * kSwHP;                   // [kSwRows][64] gradient through the lateral context
    }
    float* base_g = dcell + kSwRows * 64;               // [kSwMaxG]
    float* dzw = base_g + kSwMaxG;                      // [kSwRows][4] glimpse gradient wrt z_where
    int* grow = reinterpret_cast<int*>(dzw + kSwRows * 4);
    int* rcell = grow + kSwRows;
    int* rimg = rcell + kSwRows;

    const int b0 = blockIdx.x * p.ipc;
    const int n_img = min(p.ipc, p.B - b0);
    if (n_img <= 0) return;
    if (TC) {
        const int warp = threadIdx.x >> 5;
        if (threadIdx.x == 0) tc::init_barriers(W.b);
        if (warp == tc::kMmaWarp) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(W.b.tmem_slot()), "n"(tc::kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        tc::fence_before();
        __syncthreads();
        tc::fence_after();
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(W.tmem) : "r"(W.b.tmem_slot()));
        if (warp >= tc::kProducerWarp) {
            if (threadIdx.x == tc::kProducerWarp * 32) tc::producer_loop(wring, W.b, wstream, stages_per_wavefront, p.n_wavefronts);
            // (an exact thread index: the compiler then keeps the descriptors in uniform registers; with a per-warp lane test
            // it wraps every tcgen05.mma in a broadcast loop and the issue rate drops by a third)
            if (threadIdx.x == tc::kMmaWarp * 32) tc::mma_loop<0>(wring, W.xring, W.b, W.tmem, plan, p.n_wavefronts);
            if (threadIdx.x == (tc::kMmaWarp + 1) * 32) tc::mma_loop<1>(wring, W.xring, W.b, W.tmem, plan, p.n_wavefronts);
            __syncwarp();
            tc::fence_before();
            __syncthreads();
            if (warp == tc::kMmaWarp) {
                __syncwarp();
                tc::fence_after();
                asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(W.tmem), "n"(tc::kTmemCols));
            }
            return;
        }
    }
    const int E = p.A + 6, D = 4 + p.A + 1;
    const int CTX = p.nb.n * E;
    const int c_pt = p.F + CTX, c_box = c_pt + p.P, c_attr = c_box + 4, c_depth = c_attr + p.A;
    const int GG = p.G * p.G;
    const float keep = 1.0f - p.wheel[0];
    for (int j = threadIdx.x; j < p.G; j += kSwThreads) base_g[j] = base_coord(j, p.G);

    SW_T0();
    for (int t = p.n_wavefronts - 1; t >= 0; --t) {
        const int s0 = p.starts[t], n_cells = p.starts[t + 1] - s0;
        const int nrows = n_cells * n_img;
        sw_sync<TC>();
        if (threadIdx.x < kSwRows) {
            const int r = threadIdx.x;
            if (r < nrows) {
                const int k = r / n_img, li = r - k * n_img;
                grow[r] = (s0 + k) * p.B + b0 + li;
                rcell[r] = p.order[s0 + k];
                rimg[r] = b0 + li;
            } else {
                grow[r] = 0; rcell[r] = 0; rimg[r] = b0;
            }
        }
        sw_sync<TC>();

        // ---- gradient arriving through the lateral context of later cells (models.py:106 -> 73) ----
        for (int idx = threadIdx.x; idx < nrows * E; idx += kSwThreads) {
            const int r = idx / E, j = idx - r * E;
            const int b = rimg[r], cell = rcell[r];
            const int h = cell / p.Wc, w = cell - h * p.Wc;
            float acc = 0.0f;
            for (int s = 0; s < p.nb.n; ++s) {
                const int ch = h - p.nb.dh[s], cw = w - p.nb.dw[s];
                if (ch < 0 || ch >= p.Hc || cw < 0 || cw >= p.Wc) continue;
                const size_t cr = (size_t)p.wf_pos[ch * p.Wc + cw] * p.B + b;      // row written by this CTA earlier
                const int c = p.F + s * E + j;
                acc += p.box.dX[cr * p.box.ldX + c] + p.z.dX[cr * p.z.ldX + c] + p.obj.dX[cr * p.obj.ldX + c];
            }
            dcell[r * 64 + j] = acc;
        }
        sw_sync<TC>();

        // ---- z_pres (models.py:393-411) ----
        if (threadIdx.x < kSwRows) {
            const int r = threadIdx.x;
            float dy = 0.0f;
            if (r < nrows) {
                const size_t o = (size_t)rimg[r] * p.HW + rcell[r];
                const float logit = p.obj.Y[grow[r]];
                const float u = __ldg(p.u_pres + o);
                const float pr = sigmoid_f(clamp10(logit) + (logf(u + 10e-10f) - logf(1.0f - u + 10e-10f)));
                const float d = dcell[r * 64 + E - 1] + (p.d_pres ? __ldg(p.d_pres + o) : 0.0f);
                dy = keep * d * pr * (1.0f - pr) * clamp10_mask(logit);
                p.obj.dY[grow[r]] = dy;
            }
            gy[r * kSwHP + 0] = dy;
            gy[r * kSwHP + 1] = gy[r * kSwHP + 2] = gy[r * kSwHP + 3] = 0.0f;
        }
        sw_sync<TC>();
        SW_MARK(30);
        if (TC) mlp3_bwd_tc(W, p.obj, grow, nrows, gy);
        else mlp3_bwd(p.obj, grow, nrows, part, gy, ga, gb);
        SW_MARK(31);

        // ---- z_depth (models.py:88-97) ----
        for (int idx = threadIdx.x; idx < kSwRows * (2 + p.P); idx += kSwThreads) {
            const int r = idx / (2 + p.P), k = idx - r * (2 + p.P);
            float v = 0.0f;
            if (r < nrows) {
                const size_t g = grow[r];
                if (k >= 2) {
                    v = p.obj.dX[g * p.obj.ldX + c_pt + (k - 2)];               // passthrough features fed the obj network
                } else {
                    const size_t o = (size_t)rimg[r] * p.HW + rcell[r];
                    const float* yr = p.z.Y + g * p.z.N[2];
                    const float ls = yr[1], sg = sigmoid_f(clamp10(ls)), e = __ldg(p.eps_depth + o);
                    const float zl = yr[0] + e * (sg * 2.0f);
                    const float sq = sigmoid_f(clamp10(zl));
                    const float d_o = p.obj.dX[g * p.obj.ldX + c_depth] + dcell[r * 64 + E - 2] + (p.d_depth ? __ldg(p.d_depth + o) : 0.0f);
                    const float d_z = d_o * 4.0f * sq * (1.0f - sq) * clamp10_mask(zl);
                    if (k == 0) v = keep * (d_z + (p.d_dmean ? __ldg(p.d_dmean + o * D + D - 1) : 0.0f));
                    else v = keep * ((d_z * e + (p.d_dstd ? __ldg(p.d_dstd + o * D + D - 1) : 0.0f)) * 2.0f * sg * (1.0f - sg) * clamp10_mask(ls));
                }
                p.z.dY[g * p.z.N[2] + k] = v;
            }
            gy[r * kSwHP + k] = v;
        }
        for (int idx = threadIdx.x; idx < kSwRows * 4; idx += kSwThreads) gy[(idx >> 2) * kSwHP + 2 + p.P + (idx & 3)] = 0.0f;
        sw_sync<TC>();
        SW_MARK(32);
        if (TC) mlp3_bwd_tc(W, p.z, grow, nrows, gy);
        else mlp3_bwd(p.z, grow, nrows, part, gy, ga, gb);
        SW_MARK(33);

        // ---- z_what (models.py:83-85) ----
        for (int idx = threadIdx.x; idx < kSwRows * p.A; idx += kSwThreads) {
            const int r = idx / p.A, k = idx - r * p.A;
            float vm = 0.0f, vs = 0.0f;
            if (r < nrows) {
                const size_t g = grow[r];
                const size_t o = (size_t)rimg[r] * p.HW + rcell[r];
                const float* yr = p.enc.Y + g * p.enc.N[2];
                const float ls = yr[p.A + k], sg = sigmoid_f(clamp10(ls)), e = __ldg(p.eps_attr + o * p.A + k);
                const float d_o = p.z.dX[g * p.z.ldX + c_attr + k] + p.obj.dX[g * p.obj.ldX + c_attr + k] + dcell[r * 64 + 4 + k] +
                                  (p.d_attr ? __ldg(p.d_attr + o * p.A + k) : 0.0f);
                vm = d_o + (p.d_dmean ? __ldg(p.d_dmean + o * D + 4 + k) : 0.0f);
                vs = (d_o * e + (p.d_dstd ? __ldg(p.d_dstd + o * D + 4 + k) : 0.0f)) * 2.0f * sg * (1.0f - sg) * clamp10_mask(ls);
                p.enc.dY[g * p.enc.N[2] + k] = vm;
                p.enc.dY[g * p.enc.N[2] + p.A + k] = vs;
            }
            gy[r * kSwHP + k] = vm;
            gy[r * kSwHP + p.A + k] = vs;
        }
        for (int idx = threadIdx.x; idx < kSwRows * 4; idx += kSwThreads) gy[(idx >> 2) * kSwHP + 2 * p.A + (idx & 3)] = 0.0f;
        sw_sync<TC>();
        SW_MARK(34);
        if (TC) mlp3_bwd_tc(W, p.enc, grow, nrows, gy);
        else mlp3_bwd(p.enc, grow, nrows, part, gy, ga, gb);
        SW_MARK(35);

        // ---- glimpse: d z_where (modules.py:216-273; the image has no gradient in the model) ----
        // one warp per row (kSwThreads / 32 == kSwRows): lanes stride over the texels, one butterfly reduction per row —
        // deterministic, no atomics
        {
            const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
            if (r < nrows) {
                const size_t o = (size_t)rimg[r] * p.HW + rcell[r];
                const float4 zw = *reinterpret_cast<const float4*>(p.z_where + o * 4);
                const FwdAffine A(zw.x, zw.y, zw.z, zw.w);
                const float* grow_e = p.enc.dX + (size_t)grow[r] * p.enc.ldX;
#pragma unroll 4
                for (int tt = lane; tt < GG; tt += 32) {
                    const int i = tt / p.G, j = tt - i * p.G;
                    float ix = unnormalize(affine_coord(base_g[j], A.ax, A.cx), 0.5f * (float)p.Iw);
                    float iy = unnormalize(affine_coord(base_g[i], A.ay, A.cy), 0.5f * (float)p.Ih);
                    const float mx = (ix > 0.0f && ix < (float)(p.Iw - 1)) ? 1.0f : 0.0f;
                    const float my = (iy > 0.0f && iy < (float)(p.Ih - 1)) ? 1.0f : 0.0f;
                    ix = fminf(fmaxf(ix, 0.0f), (float)(p.Iw - 1));
                    iy = fminf(fmaxf(iy, 0.0f), (float)(p.Ih - 1));
                    const float fx0 = floorf(ix), fy0 = floorf(iy);
                    const int x0 = (int)fx0, y0 = (int)fy0;
                    const int x1 = min(x0 + 1, p.Iw - 1), y1 = min(y0 + 1, p.Ih - 1);
                    const float wx1 = ix - fx0, wx0 = fx0 + 1.0f - ix, wy1 = iy - fy0, wy0 = fy0 + 1.0f - iy;
                    float gix = 0.0f, giy = 0.0f;
                    for (int c = 0; c < p.C; ++c) {
                        const float* pl = p.image + ((size_t)rimg[r] * p.C + c) * p.Ih * p.Iw;
                        const float v00 = __ldg(pl + (size_t)y0 * p.Iw + x0), v01 = __ldg(pl + (size_t)y0 * p.Iw + x1);
                        const float v10 = __ldg(pl + (size_t)y1 * p.Iw + x0), v11 = __ldg(pl + (size_t)y1 * p.Iw + x1);
                        const float g = grow_e[c * GG + tt];
                        gix += g * ((v01 - v00) * wy0 + (v11 - v10) * wy1);
                        giy += g * ((v10 - v00) * wx0 + (v11 - v01) * wx1);
                    }
                    const float dgx = gix * mx, dgy = giy * my;
                    a0 += dgx;
                    a1 += dgy;
                    a2 = fmaf(dgx, base_g[j], a2);
                    a3 = fmaf(dgy, base_g[i], a3);
                }
            }
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
            if (lane == 0) { dzw[r * 4 + 0] = a0; dzw[r * 4 + 1] = a1; dzw[r * 4 + 2] = a2; dzw[r * 4 + 3] = a3; }
        }
        sw_sync<TC>();
        SW_MARK(38);

        // ---- z_where: box head (models.py:322-381) ----
        for (int idx = threadIdx.x; idx < kSwRows * (8 + p.P); idx += kSwThreads) {
            const int r = idx / (8 + p.P), k = idx - r * (8 + p.P);
            float v = 0.0f;
            if (r < nrows) {
                const size_t g = grow[r];
                if (k >= 8) {
                    v = p.z.dX[g * p.z.ldX + c_pt + (k - 8)];                    // passthrough features fed the z network
                } else {
                    const int kk = k & 3;                                        // component: cy, cx, height, width
                    const int slot = sw_box_slot(kk);
                    const size_t o = (size_t)rimg[r] * p.HW + rcell[r];
                    const float* yr = p.box.Y + g * p.box.N[2];
                    const float ls = yr[4 + kk], sg = sigmoid_f(clamp10(ls)), e = __ldg(p.eps_where + o * 4 + kk);
                    const float zl = yr[kk] + e * (sg * 2.0f);
                    const float sq = sigmoid_f(clamp10(zl));
                    float d_b = p.z.dX[g * p.z.ldX + c_box + slot] + p.obj.dX[g * p.obj.ldX + c_box + slot] + dcell[r * 64 + slot];
                    // glimpse gradient (d gx sums) -> d z_where: gx = xs*base + (2xt-1), ix = (gx+1)*I/2 - 0.5
                    const float hI = 0.5f * ((slot == 0 || slot == 2) ? (float)p.Iw : (float)p.Ih);
                    const float d_zw_glimpse = (slot == 0) ? 2.0f * hI * dzw[r * 4 + 0] : (slot == 1) ? 2.0f * hI * dzw[r * 4 + 1]
                                             : (slot == 2) ? hI * dzw[r * 4 + 2] : hI * dzw[r * 4 + 3];
                    const float d_zwv = d_zw_glimpse + (p.d_zw ? __ldg(p.d_zw + o * 4 + slot) : 0.0f);
                    float d_s;
                    if (kk < 2) {
                        d_b += d_zwv * (kk == 0 ? p.geom.cell_ratio_y : p.geom.cell_ratio_x);
                        d_s = d_b * p.geom.yx_scale;
                    } else {
                        d_b += (d_zwv / (kk == 2 ? p.geom.img_h : p.geom.img_w)) * p.geom.anchor;
                        d_s = d_b * p.geom.hw_scale;
                    }
                    const float d_zl = d_s * sq * (1.0f - sq) * clamp10_mask(zl);
                    if (k < 4) v = keep * (d_zl + (p.d_dmean ? __ldg(p.d_dmean + o * D + kk) : 0.0f));
                    else v = keep * ((d_zl * e + (p.d_dstd ? __ldg(p.d_dstd + o * D + kk) : 0.0f)) * 2.0f * sg * (1.0f - sg) * clamp10_mask(ls));
                }
                p.box.dY[g * p.box.N[2] + k] = v;
            }
            gy[r * kSwHP + k] = v;
        }
        for (int idx = threadIdx.x; idx < kSwRows * 4; idx += kSwThreads) gy[(idx >> 2) * kSwHP + 8 + p.P + (idx & 3)] = 0.0f;
        sw_sync<TC>();
        SW_MARK(36);
        if (TC) mlp3_bwd_tc(W, p.box, grow, nrows, gy);
        else mlp3_bwd(p.box, grow, nrows, part, gy, ga, gb);
        SW_MARK(37);
    }
    if (TC) {
        tc::fence_before();
        __syncthreads();                                // pairs with the producer / MMA warps' closing barrier
    }
}

// ---- weight packing for the two sweeps ----------------------------------------------------------------------------
constexpr int kPackMaxLayers = 16;
struct PackArgs {
    const float* W[kPackMaxLayers];
    float4* fwd[kPackMaxLayers];
    float4* bwd[kPackMaxLayers];
    int N[kPackMaxLayers], K[kPackMaxLayers];
};

__global__ void __launch_bounds__(256) sweep_pack_kernel(PackArgs p) {
    const int l = blockIdx.y;
    const float* __restrict__ W = p.W[l];
    const int N = p.N[l], K = p.K[l];
    const int nf = ((K + 3) >> 2) * N, nbk = ((N + 3) >> 2) * K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf + nbk; i += gridDim.x * blockDim.x) {
        float v[4];
        if (i < nf) {
            const int g = i / N, n = i - g * N;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (4 * g + j < K) ? __ldg(W + (size_t)n * K + 4 * g + j) : 0.0f;
            if (p.fwd[l]) p.fwd[l][i] = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            const int q = i - nf, g = q / K, k = q - g * K;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (4 * g + j < N) ? __ldg(W + (size_t)(4 * g + j) * K + k) : 0.0f;
            if (p.bwd[l]) p.bwd[l][q] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// ---- weight stream of the tensor-core sweeps --------------------------------------------------------------------------
// One CTA per stage (= 128 output features x 32 reduction indices, hi tile then lo tile, each in the 128-byte-swizzled
// K-major operand layout), stages in the order tc::mma_loop consumes them: layers in execution order; inside a layer
// chunk (256 reduction indices) -> feature tile -> k-block.  backward stream: execution order obj, z, enc, box, layers
// 2, 1, 0, operand = W^T (features = the layer's inputs, reduction over its outputs).
struct TcPackArgs {
    const float* W[tc::kMaxLayers];
    int N[tc::kMaxLayers], K[tc::kMaxLayers];    // nn.Linear weight [N][K], model order box0..box2, enc0.., z0.., obj0..
    float* fwd; float* bwd;
    int stages_fwd;
};

__global__ void __launch_bounds__(256) tc_pack_kernel(const TcPackArgs a) {
    int s = blockIdx.x;
    const bool bwd = s >= a.stages_fwd;
    if (bwd) s -= a.stages_fwd;
    float* dst = (bwd ? a.bwd : a.fwd) + (size_t)s * tc::kStageFloats;
    if (dst == nullptr || (bwd ? a.bwd : a.fwd) == nullptr) return;
    int l = 0, Kred = 0, Mout = 0;
    for (int e = 0; e < tc::kMaxLayers; ++e) {
        l = bwd ? 3 * (3 - e / 3) + (2 - e % 3) : e;
        Kred = bwd ? a.N[l] : a.K[l];
        Mout = bwd ? a.K[l] : a.N[l];
        const int cnt = tc::stages_of(Kred, Mout);
        if (s < cnt) break;
        s -= cnt;
    }
    const int KB = (Kred + 31) >> 5, tiles = (Mout + 127) >> 7, chunks = (Kred + tc::kChunkK - 1) / tc::kChunkK;
    int mt = 0, kb = 0;
    for (int g0 = 0, rem = s, found = 0; g0 < tiles && !found; g0 += tc::kTileGroup) {      // the loops of tc::mma_loop
        const int gt = min(tc::kTileGroup, tiles - g0);
        for (int c = 0; c < chunks; ++c) {
            const int kbn = min(tc::kChunkKB, KB - tc::kChunkKB * c);
            if (rem < gt * kbn) {
                mt = g0 + rem / kbn;
                kb = tc::kChunkKB * c + rem % kbn;
                found = 1;
                break;
            }
            rem -= gt * kbn;
        }
    }
    const float* __restrict__ W = a.W[l];
    const int ldw = a.K[l];
    for (int i = threadIdx.x; i < 128 * 32; i += 256) {
        const int fl = bwd ? (i & 127) : (i >> 5), kl = bwd ? (i >> 7) : (i & 31);     // coalesced along W's contiguous index
        const int f = mt * 128 + fl, k = kb * 32 + kl;
        float v = 0.0f;
        if (f < Mout && k < Kred) v = bwd ? __ldg(W + (size_t)k * ldw + f) : __ldg(W + (size_t)f * ldw + k);
        uint32_t hi, lo;
        tc::split_tf32(v, hi, lo);
        const int o = fl * 32 + ((((kl >> 2) ^ fl) & 7) << 2) + (kl & 3);
        dst[o] = __uint_as_float(hi);
        dst[o + 4096] = __uint_as_float(lo);
    }
}

}  // namespace spair

using namespace spair;

static int tc_stages(const int* n, const int* k, int backward) {
    int total = 0;
    for (int l = 0; l < tc::kMaxLayers; ++l) total += backward ? tc::stages_of(n[l], k[l]) : tc::stages_of(k[l], n[l]);
    return total;
}

extern "C" int spair_sweep_tc_stream_floats(const int* n, const int* k, int n_layers, int backward) {
    if (!n || !k || n_layers != tc::kMaxLayers) return -1;
    return tc_stages(n, k, backward) * tc::kStageFloats;
}

extern "C" int spair_sweep_tc_pack(const float* const* w, const int* n, const int* k, int n_layers, float* fwd_stream,
                                   float* bwd_stream, void* stream) {
    SPAIR_REQUIRE(w && n && k && n_layers == tc::kMaxLayers && (fwd_stream || bwd_stream));
    SPAIR_REQUIRE(((uintptr_t)fwd_stream % 16) == 0 && ((uintptr_t)bwd_stream % 16) == 0);
    TcPackArgs a;
    for (int l = 0; l < tc::kMaxLayers; ++l) {
        SPAIR_REQUIRE(w[l] && n[l] > 0 && k[l] > 0);
        a.W[l] = w[l]; a.N[l] = n[l]; a.K[l] = k[l];
    }
    a.fwd = fwd_stream; a.bwd = bwd_stream;
    a.stages_fwd = tc_stages(n, k, 0);
    const int total = a.stages_fwd + tc_stages(n, k, 1);
    tc_pack_kernel<<<total, 256, 0, (cudaStream_t)stream>>>(a);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_sweep_pack_weights(const spair_sweep_pack* layers, int n_layers, void* stream) {
    SPAIR_REQUIRE(layers && n_layers >= 1 && n_layers <= kPackMaxLayers);
    PackArgs a;
    int most = 0;
    for (int l = 0; l < n_layers; ++l) {
        const spair_sweep_pack& L = layers[l];
        SPAIR_REQUIRE(L.w && L.n > 0 && L.k > 0 && (L.fwd || L.bwd));
        SPAIR_REQUIRE(((uintptr_t)L.fwd % 16) == 0 && ((uintptr_t)L.bwd % 16) == 0);
        a.W[l] = L.w; a.fwd[l] = reinterpret_cast<float4*>(L.fwd); a.bwd[l] = reinterpret_cast<float4*>(L.bwd);
        a.N[l] = L.n; a.K[l] = L.k;
        most = max(most, ((L.k + 3) / 4) * L.n + ((L.n + 3) / 4) * L.k);
    }
    dim3 grid(min((most + 255) / 256, 64), n_layers);
    sweep_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    SPAIR_LAUNCH_CHECK();
}

// Host-side description of one MLP for spair_sweep_fwd (mirrors the C struct in include/spair_b200.h)
static bool to_mlp(const spair_sweep_mlp* m, SweepMLP& out, bool packed_w) {
    if (!m || !m->x || !m->h0 || !m->h1 || !m->y) return false;
    for (int i = 0; i < 3; ++i) {
        if (!m->wt[i] || !m->b[i] || m->k[i] <= 0 || m->n[i] <= 0 || m->n[i] > 256) return false;
        if (packed_w && (uintptr_t)m->wt[i] % 16) return false;
        out.l[i] = SweepLayer{reinterpret_cast<const float4*>(m->wt[i]), m->b[i], m->k[i], m->n[i]};
    }
    if (m->k[1] != m->n[0] || m->k[2] != m->n[1] || m->ld_x < m->k[0]) return false;
    out.X = m->x; out.ldX = m->ld_x; out.H0 = m->h0; out.H1 = m->h1; out.Y = m->y;
    return true;
}

extern "C" int spair_sweep_max_rows(void) { return kSwRows; }

#ifdef SW_TIMING
// debug builds only (-DSW_TIMING): per-phase clock64 totals of CTA 0, read and cleared
extern "C" int spair_debug_sweep_timing(long long* out64) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out64, g_sw_timing, sizeof(long long) * 64);
    long long zero[64] = {0};
    cudaMemcpyToSymbol(g_sw_timing, zero, sizeof(zero));
    return 0;
}
#endif

static int sweep_fwd_impl(const spair_sweep_dims* d, const int* order, const int* starts, const int* nb_offsets,
                          const float* image, const float* feat, const float* edge, const float* eps_where,
                          const float* eps_attr, const float* eps_depth, const float* u_pres,
                          const spair_box_geom* geom, const spair_sweep_mlp* box_mlp, const spair_sweep_mlp* enc_mlp,
                          const spair_sweep_mlp* z_mlp, const spair_sweep_mlp* obj_mlp, float* box, float* z_where,
                          float* attr, float* depth, float* pres, float* dmean, float* dstd, const float* wstream, bool use_tc,
                          void* stream) {
    SPAIR_REQUIRE(d && order && starts && nb_offsets && image && feat && edge && eps_where && eps_attr && eps_depth && u_pres);
    SPAIR_REQUIRE(geom && box && z_where && attr && depth && pres && dmean && dstd);
    SPAIR_REQUIRE(d->B > 0 && d->HW == d->Hc * d->Wc && d->G > 0 && d->G <= kSwMaxG && d->ipc >= 1 && d->n_wavefronts > 0);
    SPAIR_REQUIRE(d->max_cells * d->ipc <= kSwRows && d->n_nb >= 1 && d->n_nb <= SPAIR_MAX_NEIGHBOURS);
    SweepFwdArgs a;
    a.B = d->B; a.HW = d->HW; a.Hc = d->Hc; a.Wc = d->Wc; a.F = d->F; a.A = d->A; a.P = d->P; a.C = d->C;
    a.Ih = d->Ih; a.Iw = d->Iw; a.G = d->G; a.ipc = d->ipc; a.n_wavefronts = d->n_wavefronts;
    a.nb.n = d->n_nb;
    for (int i = 0; i < d->n_nb; ++i) { a.nb.dh[i] = nb_offsets[2 * i]; a.nb.dw[i] = nb_offsets[2 * i + 1]; }
    a.order = order; a.starts = starts; a.image = image; a.feat = feat; a.edge = edge;
    a.eps_where = eps_where; a.eps_attr = eps_attr; a.eps_depth = eps_depth; a.u_pres = u_pres; a.geom = *geom;
    SPAIR_REQUIRE(to_mlp(box_mlp, a.box, !use_tc) && to_mlp(enc_mlp, a.enc, !use_tc) && to_mlp(z_mlp, a.z, !use_tc) &&
                  to_mlp(obj_mlp, a.obj, !use_tc));
    const int E = d->A + 6, CTX = d->n_nb * E;
    SPAIR_REQUIRE(a.box.l[0].K == d->F + CTX && a.box.l[2].N == 8 + d->P);
    SPAIR_REQUIRE(a.enc.l[0].K == d->C * d->G * d->G && a.enc.l[2].N == 2 * d->A);
    SPAIR_REQUIRE(a.z.l[0].K == d->F + CTX + d->P + 4 + d->A && a.z.l[2].N == 2 + d->P);
    SPAIR_REQUIRE(a.obj.l[0].K == a.z.l[0].K + 1 && a.obj.l[2].N == 1);
    a.out_box = box; a.z_where = z_where; a.attr = attr; a.depth = depth; a.pres = pres; a.dmean = dmean; a.dstd = dstd;
    const int grid = (d->B + d->ipc - 1) / d->ipc;
    tc::Plan plan;
    plan.n_layers = 0;
    if (use_tc) {
        SPAIR_REQUIRE(wstream && ((uintptr_t)wstream % 16) == 0);
        if (getenv("SPAIR_SWEEP_TC_NO_RELU_FIXUP")) {      // diagnostic: what the exact-ReLU re-evaluation costs
            SweepMLP* mm[4] = {&a.box, &a.enc, &a.z, &a.obj};
            for (int m = 0; m < 4; ++m)
                for (int i = 0; i < 3; ++i) mm[m]->l[i].Wp = nullptr;
        }
        const SweepMLP* ms[4] = {&a.box, &a.enc, &a.z, &a.obj};
        int stages = 0;
        for (int m = 0; m < 4; ++m)
            for (int i = 0; i < 3; ++i) {
                plan.K[plan.n_layers] = ms[m]->l[i].K; plan.M[plan.n_layers] = ms[m]->l[i].N;
                stages += tc::stages_of(ms[m]->l[i].K, ms[m]->l[i].N);
                ++plan.n_layers;
            }
        const size_t smem = 1024 + tc::kRingBytes + tc::kBarBytes + sizeof(float) * (size_t)(kSwRows * kSwHP + kSwMaxG + kSwRows * 4) +
                            sizeof(int) * 3 * kSwRows;
        static size_t smem_set_tc[kMaxDevices] = {0};
        if (ensure_dynamic_smem(sweep_fwd_kernel<true>, smem, smem_set_tc) != cudaSuccess) return (int)cudaGetLastError();
        sweep_fwd_kernel<true><<<grid, tc::kThreads, smem, (cudaStream_t)stream>>>(a, plan, wstream, stages);
        SPAIR_LAUNCH_CHECK();
    }
    const size_t smem = sizeof(float) * (size_t)(kSwRows * kSwKCP + kSwPart + 3 * kSwRows * kSwHP + kSwMaxG + kSwRows * 4) +
                        sizeof(int) * 3 * kSwRows;
    static size_t smem_set[kMaxDevices] = {0};
    if (ensure_dynamic_smem(sweep_fwd_kernel<false>, smem, smem_set) != cudaSuccess) return (int)cudaGetLastError();
    sweep_fwd_kernel<false><<<grid, kSwThreads, smem, (cudaStream_t)stream>>>(a, plan, nullptr, 0);
    SPAIR_LAUNCH_CHECK();
}

#ifdef SW_TIMING
extern "C" int spair_debug_sweep_tc_flags(int flags) {
    return (int)cudaMemcpyToSymbol(tc::g_tc_debug, &flags, sizeof(int));
}
#endif

extern "C" int spair_sweep_fwd(const spair_sweep_dims* d, const int* order, const int* starts, const int* nb_offsets,
                               const float* image, const float* feat, const float* edge, const float* eps_where,
                               const float* eps_attr, const float* eps_depth, const float* u_pres,
                               const spair_box_geom* geom, const spair_sweep_mlp* box_mlp, const spair_sweep_mlp* enc_mlp,
                               const spair_sweep_mlp* z_mlp, const spair_sweep_mlp* obj_mlp, float* box, float* z_where,
                               float* attr, float* depth, float* pres, float* dmean, float* dstd, void* stream) {
    return sweep_fwd_impl(d, order, starts, nb_offsets, image, feat, edge, eps_where, eps_attr, eps_depth, u_pres, geom, box_mlp,
                          enc_mlp, z_mlp, obj_mlp, box, z_where, attr, depth, pres, dmean, dstd, nullptr, false, stream);
}

extern "C" int spair_sweep_fwd_tc(const spair_sweep_dims* d, const int* order, const int* starts, const int* nb_offsets,
                                  const float* image, const float* feat, const float* edge, const float* eps_where,
                                  const float* eps_attr, const float* eps_depth, const float* u_pres,
                                  const spair_box_geom* geom, const spair_sweep_mlp* box_mlp, const spair_sweep_mlp* enc_mlp,
                                  const spair_sweep_mlp* z_mlp, const spair_sweep_mlp* obj_mlp, float* box, float* z_where,
                                  float* attr, float* depth, float* pres, float* dmean, float* dstd, const float* wstream,
                                  void* stream) {
    return sweep_fwd_impl(d, order, starts, nb_offsets, image, feat, edge, eps_where, eps_attr, eps_depth, u_pres, geom, box_mlp,
                          enc_mlp, z_mlp, obj_mlp, box, z_where, attr, depth, pres, dmean, dstd, wstream, true, stream);
}

static bool to_mlp_bwd(const spair_sweep_mlp_bwd* m, SweepMLPBwd& out, bool need_w) {
    if (!m || !m->h0 || !m->h1 || !m->y || !m->dx || !m->dh0 || !m->dh1 || !m->dy) return false;
    for (int i = 0; i < 3; ++i) {
        if ((need_w && !m->w[i]) || m->k[i] <= 0 || m->n[i] <= 0 || m->n[i] > 256) return false;
        if ((uintptr_t)m->w[i] % 16) return false;
        out.W[i] = reinterpret_cast<const float4*>(m->w[i]); out.K[i] = m->k[i]; out.N[i] = m->n[i];
    }
    if (m->k[1] != m->n[0] || m->k[2] != m->n[1] || m->ld_dx < m->k[0]) return false;
    out.H0 = m->h0; out.H1 = m->h1; out.Y = m->y; out.dX = m->dx; out.ldX = m->ld_dx;
    out.dH0 = m->dh0; out.dH1 = m->dh1; out.dY = m->dy;
    return true;
}

static int sweep_bwd_impl(const spair_sweep_dims* d, const int* order, const int* starts, const int* wf_pos,
                          const int* nb_offsets, const float* image, const float* z_where, const float* eps_where,
                          const float* eps_attr, const float* eps_depth, const float* u_pres, const float* wheel,
                          const spair_box_geom* geom, const spair_sweep_mlp_bwd* box_mlp,
                          const spair_sweep_mlp_bwd* enc_mlp, const spair_sweep_mlp_bwd* z_mlp,
                          const spair_sweep_mlp_bwd* obj_mlp, const float* d_zw, const float* d_attr,
                          const float* d_depth, const float* d_pres, const float* d_dmean, const float* d_dstd,
                          const float* wstream, bool use_tc, void* stream) {
    SPAIR_REQUIRE(d && order && starts && wf_pos && nb_offsets && image && z_where && eps_where && eps_attr && eps_depth && u_pres);
    SPAIR_REQUIRE(wheel && geom && ((uintptr_t)z_where % 16) == 0 && (d_dmean == nullptr) == (d_dstd == nullptr));
    SPAIR_REQUIRE(d->B > 0 && d->HW == d->Hc * d->Wc && d->G > 0 && d->G <= kSwMaxG && d->ipc >= 1 && d->n_wavefronts > 0);
    SPAIR_REQUIRE(d->max_cells * d->ipc <= kSwRows && d->n_nb >= 1 && d->n_nb <= SPAIR_MAX_NEIGHBOURS && d->A + 6 <= 64);
    SweepBwdArgs a;
    a.B = d->B; a.HW = d->HW; a.Hc = d->Hc; a.Wc = d->Wc; a.F = d->F; a.A = d->A; a.P = d->P; a.C = d->C;
    a.Ih = d->Ih; a.Iw = d->Iw; a.G = d->G; a.ipc = d->ipc; a.n_wavefronts = d->n_wavefronts;
    a.nb.n = d->n_nb;
    for (int i = 0; i < d->n_nb; ++i) { a.nb.dh[i] = nb_offsets[2 * i]; a.nb.dw[i] = nb_offsets[2 * i + 1]; }
    a.order = order; a.starts = starts; a.wf_pos = wf_pos; a.image = image; a.z_where = z_where;
    a.eps_where = eps_where; a.eps_attr = eps_attr; a.eps_depth = eps_depth; a.u_pres = u_pres; a.wheel = wheel;
    a.geom = *geom;
    SPAIR_REQUIRE(to_mlp_bwd(box_mlp, a.box, !use_tc) && to_mlp_bwd(enc_mlp, a.enc, !use_tc) && to_mlp_bwd(z_mlp, a.z, !use_tc) &&
                  to_mlp_bwd(obj_mlp, a.obj, !use_tc));
    const int E = d->A + 6, CTX = d->n_nb * E;
    SPAIR_REQUIRE(a.box.K[0] == d->F + CTX && a.box.N[2] == 8 + d->P && a.enc.K[0] == d->C * d->G * d->G && a.enc.N[2] == 2 * d->A);
    SPAIR_REQUIRE(a.z.K[0] == d->F + CTX + d->P + 4 + d->A && a.z.N[2] == 2 + d->P && a.obj.K[0] == a.z.K[0] + 1 && a.obj.N[2] == 1);
    a.encX = nullptr; a.ld_encX = 0;
    a.d_zw = d_zw; a.d_attr = d_attr; a.d_depth = d_depth; a.d_pres = d_pres; a.d_dmean = d_dmean; a.d_dstd = d_dstd;
    const int grid = (d->B + d->ipc - 1) / d->ipc;
    tc::Plan plan;
    plan.n_layers = 0;
    if (use_tc) {
        SPAIR_REQUIRE(wstream && ((uintptr_t)wstream % 16) == 0);
        const SweepMLPBwd* ms[4] = {&a.obj, &a.z, &a.enc, &a.box};
        int stages = 0;
        for (int m = 0; m < 4; ++m)
            for (int i = 2; i >= 0; --i) {
                plan.K[plan.n_layers] = ms[m]->N[i]; plan.M[plan.n_layers] = ms[m]->K[i];
                // more feature tiles than one accumulator group: the layer's activation chunks must all stay resident
                SPAIR_REQUIRE((ms[m]->K[i] + 127) / 128 <= tc::kTileGroup || (ms[m]->N[i] + tc::kChunkK - 1) / tc::kChunkK <= tc::kXSlots);
                stages += tc::stages_of(ms[m]->N[i], ms[m]->K[i]);
                ++plan.n_layers;
            }
        const size_t smem = 1024 + tc::kRingBytes + tc::kBarBytes +
                            sizeof(float) * (size_t)(kSwRows * kSwHP + kSwRows * 64 + kSwMaxG + kSwRows * 4) + sizeof(int) * 3 * kSwRows;
        static size_t smem_set_tc[kMaxDevices] = {0};
        if (ensure_dynamic_smem(sweep_bwd_kernel<true>, smem, smem_set_tc) != cudaSuccess) return (int)cudaGetLastError();
        sweep_bwd_kernel<true><<<grid, tc::kThreads, smem, (cudaStream_t)stream>>>(a, plan, wstream, stages);
        SPAIR_LAUNCH_CHECK();
    }
    const size_t smem = sizeof(float) * (size_t)(kSwPart + 3 * kSwRows * kSwHP + kSwRows * 64 + kSwMaxG + kSwRows * 4) +
                        sizeof(int) * 3 * kSwRows;
    static size_t smem_set[kMaxDevices] = {0};
    if (ensure_dynamic_smem(sweep_bwd_kernel<false>, smem, smem_set) != cudaSuccess) return (int)cudaGetLastError();
    sweep_bwd_kernel<false><<<grid, kSwThreads, smem, (cudaStream_t)stream>>>(a, plan, nullptr, 0);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_sweep_bwd(const spair_sweep_dims* d, const int* order, const int* starts, const int* wf_pos,
                               const int* nb_offsets, const float* image, const float* z_where, const float* eps_where,
                               const float* eps_attr, const float* eps_depth, const float* u_pres, const float* wheel,
                               const spair_box_geom* geom, const spair_sweep_mlp_bwd* box_mlp,
                               const spair_sweep_mlp_bwd* enc_mlp, const spair_sweep_mlp_bwd* z_mlp,
                               const spair_sweep_mlp_bwd* obj_mlp, const float* d_zw, const float* d_attr,
                               const float* d_depth, const float* d_pres, const float* d_dmean, const float* d_dstd,
                               void* stream) {
    return sweep_bwd_impl(d, order, starts, wf_pos, nb_offsets, image, z_where, eps_where, eps_attr, eps_depth, u_pres, wheel,
                          geom, box_mlp, enc_mlp, z_mlp, obj_mlp, d_zw, d_attr, d_depth, d_pres, d_dmean, d_dstd, nullptr, false,
                          stream);
}

extern "C" int spair_sweep_bwd_tc(const spair_sweep_dims* d, const int* order, const int* starts, const int* wf_pos,
                                  const int* nb_offsets, const float* image, const float* z_where, const float* eps_where,
                                  const float* eps_attr, const float* eps_depth, const float* u_pres, const float* wheel,
                                  const spair_box_geom* geom, const spair_sweep_mlp_bwd* box_mlp,
                                  const spair_sweep_mlp_bwd* enc_mlp, const spair_sweep_mlp_bwd* z_mlp,
                                  const spair_sweep_mlp_bwd* obj_mlp, const float* d_zw, const float* d_attr,
                                  const float* d_depth, const float* d_pres, const float* d_dmean, const float* d_dstd,
                                  const float* wstream, void* stream) {
    return sweep_bwd_impl(d, order, starts, wf_pos, nb_offsets, image, z_where, eps_where, eps_attr, eps_depth, u_pres, wheel,
                          geom, box_mlp, enc_mlp, z_mlp, obj_mlp, d_zw, d_attr, d_depth, d_pres, d_dmean, d_dstd, wstream, true,
                          stream);
}
