// tcgen05 GEMM for the path's own dense contractions (the decoder MLP, reference models.py:165,474-500, and the weight
// gradients of the five per-object MLPs, modules.py:124-165), fp32 in / fp32 out at fp32 accuracy.
//
// The 5th-generation tensor cores have no fp32 input type, and plain TF32 (10 mantissa bits) misses the parity tolerance
// (rtol 1e-4) by an order of magnitude.  Each operand is therefore split in shared memory into x = hi + lo, hi = x rounded
// to TF32 and lo = x - hi (exact in fp32), and every k-step issues three MMAs into the same TMEM accumulator:
// hi*hi + hi*lo + lo*hi ("3xTF32"; the dropped lo*lo term is 2^-22 relative).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer: raw fp32 tiles global -> shared (128-byte swizzle), one mbarrier per ring stage
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 (M = 128, N = BN, K = 8), accumulators in TMEM,
//               double-buffered (2 x 256 columns) so that the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2-9   splitters: raw tile -> hi (in place) and lo (second buffer), fence.proxy.async, signal the MMA warp
//   warps 10-17 epilogue: tcgen05.ld TMEM -> registers, bias / ReLU / texel decode, transpose through shared memory,
//               coalesced 128-bit global stores (two warps per TMEM lane quarter, each takes half of the columns)
// Either operand may be K-major (reduction index contiguous in memory) or MN-major, so that y = x W^T, dx = dy W and
// dW = dy^T x all run on the stored tensors without a transposed copy.  A reduction that is much longer than the output
// is wide (the weight gradients: K = B*HW rows) is split over CTAs into a workspace and summed in a fixed order by a
// second kernel (deterministic; no atomics).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace spair {
namespace gemm {

constexpr int BM = 128;          // rows of one output tile = TMEM lanes
constexpr int BK = 32;           // floats per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;        // k per tcgen05.mma.kind::tf32
constexpr int kSplitWarps = 8, kEpiWarps = 8;
constexpr int kSplitWarp0 = 2, kEpiWarp0 = kSplitWarp0 + kSplitWarps;
constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);   // 18 warps
constexpr int kStgPitch = 20;     // floats per row of the epilogue transpose buffer (32 rows x 16 columns per pass)
constexpr int kAccumCols = 256;  // TMEM columns per accumulator buffer (two buffers = the whole 512-column TMEM)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
// TMA im2col load (4-D NHWC tensor map created by cuTensorMapEncodeIm2col): `pixelsPerColumn` consecutive output pixels
// starting at input-space base (w, h) of image n, traversed with the convolution stride inside the map's bounding box,
// 32 channels from channel c, at filter tap offset (off_w, off_h).  Lands as [pixel][32 channels] rows of 128 bytes.
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                int off_w, int off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"((unsigned short)off_w), "h"((unsigned short)off_h)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (PTX ISA "tcgen05 shared memory descriptor"; same bit layout as cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in
// [46,48), layout type in [61,64): 2 = SWIZZLE_128B (16-byte chunks; K-major operands), 1 = SWIZZLE_128B_BASE32B (32-byte
// chunks, Swizzle<2,5,2>: the only swizzled layout the hardware accepts for MN-major 32-bit operands).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), majors at bits
// 15 / 16 (0 = K-major), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t instr_desc(int n, bool a_kmajor, bool b_kmajor) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_kmajor ? 0u : 1u) << 15) | ((b_kmajor ? 0u : 1u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct Params {
    int M, N, K;          // output rows / columns, reduction length
    int splits, k_per_split;   // k_per_split is a multiple of BK
    float* C;             // [M, ldc] when splits == 1, else workspace [splits][M][N] (ldc = N)
    int ldc;
    const float* bias;    // [N] or nullptr (applied here only when splits == 1)
    int epilogue;         // SPAIR_GEMM_EPI_*
    int period;           // texel decode: channels per texel (C + 1)
    float s_colour, s_alpha, b_alpha;
    int debug;            // diagnostics (SPAIR_GEMM_DEBUG): 1 = splitters skip their work, 2 = one MMA per k-step, 4 = no stores
    int b_presplit;       // the B operand (a weight matrix) arrives already split: map_b = its TF32 hi plane, map_b2 = its lo plane;
                          // the producer loads both and the splitters only convert the A tile (half of their work)
    int acc_split;        // 1, or 4 (BN <= 128): k-block j accumulates into TMEM accumulator j % 4, summed in the epilogue
    // implicit-GEMM convolution on a channels-last input (reference Backbone, modules.py:44-66): 0 = plain GEMM;
    // 1 = the A operand is the patch matrix of x, read by TMA im2col (forward, rows = output pixels);
    // 2 = the B operand is the patch matrix (weight gradient, reduction over the output pixels)
    int conv, cv_Ho, cv_Wo, cv_k, cv_s, cv_cpb;   // output grid, kernel side, stride, 32-channel blocks per tap
    int cv_base;          // input-space coordinate of output pixel 0's window (0; -(T-1) for the transposed convolution)
    // input gradient of a strided convolution as s*s stride-1 sub-convolutions over dy (one per output parity class): GEMM
    // row m = (b*cv_Ho + i)*cv_Wo + j is pixel (s*i + py, s*j + px) of dx [B, om_H, om_W, N]; om_s = 0: rows are stored as is
    int om_s, om_H, om_W, om_py, om_px;
    unsigned* kink_ws;    // ReLU sign fix-up list: [0] = counter, [1 .. kink_cap] = row * N + col of uncertain outputs
    int kink_cap;
};

// A ReLU output whose pre-activation is within the GEMM's own rounding error of zero may take the other branch than the
// fp32 reference does; the gradient of every layer below then differs by ~1/sqrt(#activations) (a "kink").  Outputs with
// |v| < kKinkTol * (largest |v| of the 32 neighbouring columns) are therefore listed and re-evaluated by
// relu_fixup_kernel with float64 accumulation, so the sign decision is the exact one.
constexpr float kKinkTol = 4.8828125e-4f;   // 2^-11: > 30x the measured worst-case error of a K = 2048 reduction

template <int BN>
struct Cfg {
    static constexpr int kStageA = BM * BK * 4, kStageB = BN * BK * 4;
    static constexpr int kStage = 2 * (kStageA + kStageB);   // hi + lo of both operands
    static constexpr int kStages = (200 * 1024) / kStage < 4 ? (200 * 1024) / kStage : 4;
    static constexpr int kSmem = kStages * kStage + 1024 /* alignment slack */ + 256 /* barriers */ + kEpiWarps * 32 * kStgPitch * 4;
};

template <int BN, bool A_K, bool B_K>
__global__ void __launch_bounds__(kThreads, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              const __grid_constant__ CUtensorMap map_b2, const Params p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // stage s: [A hi | B hi | A lo | B lo]
    auto a_hi = [&](int s) { return base + s * C::kStage; };
    auto b_hi = [&](int s) { return base + s * C::kStage + C::kStageA; };
    auto a_lo = [&](int s) { return base + s * C::kStage + C::kStageA + C::kStageB; };
    auto b_lo = [&](int s) { return base + s * C::kStage + 2 * C::kStageA + C::kStageB; };
    const uint32_t bars = base + C::kStages * C::kStage;
    auto bar_full = [&](int s) { return bars + 8 * s; };
    auto bar_ready = [&](int s) { return bars + 8 * (C::kStages + s); };
    auto bar_empty = [&](int s) { return bars + 8 * (2 * C::kStages + s); };
    auto bar_acc_full = [&](int a) { return bars + 8 * (3 * C::kStages + a); };
    auto bar_acc_empty = [&](int a) { return bars + 8 * (3 * C::kStages + 2 + a); };
    const uint32_t tmem_slot = bars + 8 * (3 * C::kStages + 4);
    const uint32_t stage_out = bars + 256;   // epilogue transpose buffers: 4 warps x 32 rows x kStgPitch floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
    const int n_work = m_tiles * n_tiles * p.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_ready(s), kSplitWarps);
            mbar_init(bar_empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_acc_full(a), 1);
            mbar_init(bar_acc_empty(a), kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * kAccumCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // work item -> (m tile, n tile, split); n fastest so that concurrently running CTAs share the A rows in L2
    auto decode = [&](int w, int& mt, int& nt, int& sp) {
        nt = w % n_tiles;
        int r = w / n_tiles;
        mt = r % m_tiles;
        sp = r / m_tiles;
    };
    auto k_blocks_of = [&](int sp) {
        int k0 = sp * p.k_per_split;
        int k1 = min(p.K, k0 + p.k_per_split);
        return (k1 - k0 + BK - 1) / BK;
    };

    if (warp == 0) {
        // ------------------------------------------------ TMA producer ------------------------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                int mt, nt, sp;
                decode(w, mt, nt, sp);
                const int kb_n = k_blocks_of(sp), k_base = sp * p.k_per_split;
                for (int kb = 0; kb < kb_n; ++kb) {
                    mbar_wait(bar_empty(stage), phase ^ 1);
                    // (weight gradient of a convolution: a last column tile may hold fewer than BN / 32 patch atoms)
                    const int b_atoms = p.conv == 2 ? min(BN / 32, (p.N - nt * BN + 31) / 32) : BN / 32;
                    mbar_expect_tx(bar_full(stage), C::kStageA + (p.conv == 2 ? b_atoms * 4096 : C::kStageB) + (p.b_presplit ? C::kStageB : 0));
                    const int k0 = k_base + kb * BK;
                    if (p.conv == 1) {
                        // k-block -> (tap, 32-channel block); tile row 0 -> (image, oy, ox); 128 pixels in one instruction
                        const int kbg = k0 / BK, tap = kbg / p.cv_cpb, cg = kbg - tap * p.cv_cpb;
                        const int kh = tap / p.cv_k, kw = tap - kh * p.cv_k;
                        const int m0 = mt * BM, ox = m0 % p.cv_Wo, t = m0 / p.cv_Wo, oy = t % p.cv_Ho, n = t / p.cv_Ho;
                        tma_load_im2col(a_hi(stage), &map_a, bar_full(stage), cg * 32, ox * p.cv_s + p.cv_base, oy * p.cv_s + p.cv_base, n, kw, kh);
                    } else if (A_K) {
                        tma_load_2d(a_hi(stage), &map_a, bar_full(stage), k0, mt * BM);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 32; ++j) tma_load_2d(a_hi(stage) + j * 4096, &map_a, bar_full(stage), mt * BM + 32 * j, k0);
                    }
                    if (p.conv == 2) {
                        // 32 output pixels (the k-block) x BN / 32 (tap, channel block) atoms of the patch matrix
                        const int ox = k0 % p.cv_Wo, t = k0 / p.cv_Wo, oy = t % p.cv_Ho, n = t / p.cv_Ho;
#pragma unroll
                        for (int j = 0; j < BN / 32; ++j) {
                            if (j >= b_atoms) break;
                            const int colg = nt * (BN / 32) + j, tap = colg / p.cv_cpb, cg = colg - tap * p.cv_cpb;
                            const int kh = tap / p.cv_k, kw = tap - kh * p.cv_k;
                            tma_load_im2col(b_hi(stage) + j * 4096, &map_b, bar_full(stage), cg * 32, ox * p.cv_s, oy * p.cv_s, n, kw, kh);
                        }
                    } else if (B_K) {
                        tma_load_2d(b_hi(stage), &map_b, bar_full(stage), k0, nt * BN);
                        if (p.b_presplit) tma_load_2d(b_lo(stage), &map_b2, bar_full(stage), k0, nt * BN);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 32; ++j) tma_load_2d(b_hi(stage) + j * 4096, &map_b, bar_full(stage), nt * BN + 32 * j, k0);
                        if (p.b_presplit) {
#pragma unroll
                            for (int j = 0; j < BN / 32; ++j) tma_load_2d(b_lo(stage) + j * 4096, &map_b2, bar_full(stage), nt * BN + 32 * j, k0);
                        }
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer --------------------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc(BN, A_K, B_K);
            // K-major (SWIZZLE_128B): 8-row groups of 1024 B (SBO), k-step of 8 floats = +32 B inside the swizzle row.
            // MN-major (SWIZZLE_128B_BASE32B): a TMA box is 32 k-rows x 32 columns (128 B per row): 32-column atoms of
            // 4096 B (LBO), 4-k groups of 512 B (SBO), k-step of 8 rows = +1024 B.
            constexpr uint32_t a_lbo = A_K ? 16 : 4096, b_lbo = B_K ? 16 : 4096;
            constexpr uint32_t a_sbo = A_K ? 1024 : 512, b_sbo = B_K ? 1024 : 512;
            constexpr uint32_t a_lay = A_K ? 2 : 1, b_lay = B_K ? 2 : 1;
            constexpr uint32_t a_step = A_K ? UMMA_K * 4 : 1024, b_step = B_K ? UMMA_K * 4 : 1024;
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
                int mt, nt, sp;
                decode(w, mt, nt, sp);
                const int kb_n = k_blocks_of(sp);
                // The tensor core adds every MMA into the fp32 accumulator with truncation, so the error of a long reduction
                // grows with the number of accumulations.  acc_split = 4 deals the k-blocks over four accumulators (the whole
                // TMEM, no double buffering) and the epilogue sums them with round-to-nearest adds: ~8x smaller error,
                // used for the K = 2048 forward convolutions whose ReLU decisions the reference's gradients depend on.
                const bool split4 = p.acc_split == 4;
                const int acc = split4 ? 0 : (it & 1);
                const uint32_t acc_phase = split4 ? (it & 1) : ((it >> 1) & 1);
                mbar_wait(bar_acc_empty(acc), acc_phase ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < kb_n; ++kb) {
                    const uint32_t d = tmem_base + (split4 ? (kb & 3) * 128 : acc * kAccumCols);
                    const int first = split4 ? (kb < 4) : (kb == 0);
                    mbar_wait(bar_ready(stage), phase);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                        const uint64_t dah = smem_desc(a_hi(stage) + kk * a_step, a_lbo, a_sbo, a_lay);
                        const uint64_t dal = smem_desc(a_lo(stage) + kk * a_step, a_lbo, a_sbo, a_lay);
                        const uint64_t dbh = smem_desc(b_hi(stage) + kk * b_step, b_lbo, b_sbo, b_lay);
                        const uint64_t dbl = smem_desc(b_lo(stage) + kk * b_step, b_lbo, b_sbo, b_lay);
                        if (!(p.debug & 2)) {
                            umma_tf32(d, dal, dbh, idesc, !(first && kk == 0));
                            umma_tf32(d, dah, dbl, idesc, 1);
                            umma_tf32(d, dah, dbh, idesc, 1);
                        } else {
                            umma_tf32(d, dah, dbh, idesc, !(first && kk == 0));
                        }
                    }
                    umma_commit(bar_empty(stage));   // the stage may be refilled once these MMAs have read it
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(bar_acc_full(acc));
            }
        }
    } else if (warp < kEpiWarp0) {
        // ------------------------------------------------ splitters ---------------------------------------------------
        const int t = threadIdx.x - kSplitWarp0 * 32;
        constexpr int kChunks = (C::kStageA + C::kStageB) / 16;
        constexpr uint32_t kLoOff = C::kStageA + C::kStageB;
        int stage = 0;
        uint32_t phase = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            int mt, nt, sp;
            decode(w, mt, nt, sp);
            const int kb_n = k_blocks_of(sp);
            for (int kb = 0; kb < kb_n; ++kb) {
                mbar_wait(bar_full(stage), phase);
                const uint32_t s0 = a_hi(stage);
#pragma unroll 4
                for (int c = t; c < ((p.debug & 1) ? 0 : (p.b_presplit ? C::kStageA / 16 : kChunks)); c += kSplitWarps * 32) {
                    const uint32_t addr = s0 + c * 16;
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
                    // debug 8: hi = the raw word (valid only if the tensor core TRUNCATES the low 13 mantissa bits of a tf32
                    // operand), so only lo is written
                    const uint32_t rnd = (p.debug & 8) ? 0u : 0x1000u;
                    uint32_t h0 = (__float_as_uint(v.x) + rnd) & 0xffffe000u, h1 = (__float_as_uint(v.y) + rnd) & 0xffffe000u;
                    uint32_t h2 = (__float_as_uint(v.z) + rnd) & 0xffffe000u, h3 = (__float_as_uint(v.w) + rnd) & 0xffffe000u;
                    // lo = x - hi is exact in fp32 (13 significant bits); rounded to TF32 here (the tensor core would truncate it)
                    const uint32_t l0 = (__float_as_uint(v.x - __uint_as_float(h0)) + 0x1000u) & 0xffffe000u;
                    const uint32_t l1 = (__float_as_uint(v.y - __uint_as_float(h1)) + 0x1000u) & 0xffffe000u;
                    const uint32_t l2 = (__float_as_uint(v.z - __uint_as_float(h2)) + 0x1000u) & 0xffffe000u;
                    const uint32_t l3 = (__float_as_uint(v.w - __uint_as_float(h3)) + 0x1000u) & 0xffffe000u;
                    if (!(p.debug & 8)) asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr + kLoOff), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's async reads
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ready(stage));
                if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------ epilogue ----------------------------------------------------
        const int q = warp & 3;   // TMEM lane quarter this warp may access
        const int e = warp - kEpiWarp0;
        constexpr int kChunks32 = BN / 32, kHalf = (kChunks32 + 1) / 2;
        const int c_begin = (e < 4) ? 0 : kHalf * 32, c_end = (e < 4) ? kHalf * 32 : BN;   // two warps share a lane quarter
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            int mt, nt, sp;
            decode(w, mt, nt, sp);
            const bool split4 = p.acc_split == 4;
            const int acc = split4 ? 0 : (it & 1);
            const uint32_t acc_phase = split4 ? (it & 1) : ((it >> 1) & 1);
            const int n_acc = split4 ? min(4, k_blocks_of(sp)) : 1;
            mbar_wait(bar_acc_full(acc), acc_phase);
            tc_fence_after();
            const int row0 = mt * BM + q * 32;
            const int n0 = nt * BN;
            float* cbase = p.C + (size_t)sp * p.M * p.N;
            const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(cbase) & 15) == 0);
            const uint32_t stg = stage_out + e * (32 * kStgPitch * 4);
#pragma unroll 1
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccumCols + c0, r);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                for (int a = 1; a < n_acc; ++a) {      // acc_split: the other accumulators of this tile (128 columns apart)
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * 128 + c0, r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
                }
                if (n0 + c0 >= p.N || (p.debug & 4)) continue;   // warp-uniform
                if (p.splits == 1) {
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + c0 + j < p.N) v[j] += __ldg(p.bias + n0 + c0 + j);
                    }
                    if (p.epilogue == SPAIR_GEMM_EPI_RELU) {
                        if (p.kink_ws && row0 + lane < p.M) {
                            float vmax = 0.0f;
#pragma unroll
                            for (int j = 0; j < 32; ++j) vmax = fmaxf(vmax, fabsf(v[j]));
                            const float tol = vmax * kKinkTol;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (fabsf(v[j]) <= tol && n0 + c0 + j < p.N) {
                                    const unsigned slot = atomicAdd(p.kink_ws, 1u);
                                    if (slot < (unsigned)p.kink_cap) p.kink_ws[1 + slot] = (unsigned)(row0 + lane) * (unsigned)p.N + (unsigned)(n0 + c0 + j);
                                }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                    } else if (p.epilogue == SPAIR_GEMM_EPI_TEXEL) {
                        // reference models.py:485-493: colour = sigma(OBJ_LOGIT_SCALE * l), alpha = sigma(ALPHA_LOGIT_SCALE * l
                        // + ALPHA_LOGIT_BIAS), sigma(x) = 1 / (exp(-x) + 1) (modules.py:186-187)
                        int ch = (n0 + c0) % p.period;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const bool is_alpha = ch == p.period - 1;
                            const float x = is_alpha ? __fadd_rn(__fmul_rn(v[j], p.s_alpha), p.b_alpha) : __fmul_rn(v[j], p.s_colour);
                            // the alpha channel stores the complement 1 - sigma(x) = e / (e + 1): the +5 bias saturates alpha
                            // near 1, and the renderer's backward needs sigma'(x) = s (1 - s) to full relative accuracy
                            const float e = __expf(-x);
                            v[j] = __fdividef(is_alpha ? e : 1.0f, e + 1.0f);
                            if (++ch == p.period) ch = 0;
                        }
                    }
                }
                // thread = row in TMEM; transpose through shared memory (two passes of 16 columns) so that every store
                // instruction writes whole 64-byte row segments instead of 16 bytes in each of 32 rows
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stg + (lane * kStgPitch + j) * 4), "f"(v[16 * h + j]),
                                     "f"(v[16 * h + j + 1]), "f"(v[16 * h + j + 2]), "f"(v[16 * h + j + 3])
                                     : "memory");
                    __syncwarp();
                    const int cc = (lane & 3) * 4, col = n0 + c0 + 16 * h + cc;
                    const int r_in = (lane >> 3) + 4 * ((lane >> 2) & 1);   // rows r and r+4 of a quarter-warp: distinct banks
#pragma unroll
                    for (int itr = 0; itr < 4; ++itr) {
                        const int rr = itr * 8 + r_in;
                        float4 o;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(stg + (rr * kStgPitch + cc) * 4));
                        const int grow = row0 + rr;
                        if (grow < p.M) {
                            size_t orow = (size_t)grow;
                            if (p.om_s) {      // parity-class row -> pixel of the full-resolution gradient map
                                const int j = grow % p.cv_Wo, t = grow / p.cv_Wo, i = t % p.cv_Ho, b = t / p.cv_Ho;
                                orow = ((size_t)b * p.om_H + (size_t)p.om_s * i + p.om_py) * p.om_W + (size_t)p.om_s * j + p.om_px;
                            }
                            float* dst = cbase + orow * p.ldc + col;
                            if (vec_ok && col + 4 <= p.N) {
                                *reinterpret_cast<float4*>(dst) = o;
                            } else {
                                if (col < p.N) dst[0] = o.x;
                                if (col + 1 < p.N) dst[1] = o.y;
                                if (col + 2 < p.N) dst[2] = o.z;
                                if (col + 3 < p.N) dst[3] = o.w;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(acc));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * kAccumCols));
    }
}

// One warp per listed output: v = sum_k A[m][k] B[n][k] + bias[n] accumulated in float64 (exact to ~1e-16 relative, so the
// sign is the true one), C[m][n] = max(v, 0).  K-major operands only (the forward layers).
__global__ void __launch_bounds__(256) relu_fixup_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                         const float* __restrict__ bias, float* __restrict__ C, int ldc, int N, int K,
                                                         const unsigned* __restrict__ kink_ws, int cap) {
    const unsigned count = min(kink_ws[0], (unsigned)cap);
    const int lane = threadIdx.x & 31;
    const unsigned warps = gridDim.x * (blockDim.x >> 5);
    for (unsigned e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < count; e += warps) {
        const unsigned idx = kink_ws[1 + e];
        const unsigned m = idx / (unsigned)N, n = idx - m * (unsigned)N;
        const float* a = A + (size_t)m * lda;
        const float* b = B + (size_t)n * ldb;
        double acc = 0.0;
        for (int k = lane; k < K; k += 32) acc += (double)a[k] * (double)b[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            if (bias) acc += (double)bias[n];
            C[(size_t)m * ldc + n] = acc > 0.0 ? (float)acc : 0.0f;
        }
    }
}

// The same for an implicit-GEMM convolution: the patch row of output pixel m is gathered from the channels-last input.
__global__ void __launch_bounds__(256) relu_fixup_conv_kernel(const float* __restrict__ x, int H, int W, int Cin, int k, int s, int Ho, int Wo,
                                                              const float* __restrict__ w, int ldw, const float* __restrict__ bias,
                                                              float* __restrict__ C, int ldc, int N, const unsigned* __restrict__ kink_ws, int cap) {
    const unsigned count = min(kink_ws[0], (unsigned)cap);
    const int lane = threadIdx.x & 31;
    const unsigned warps = gridDim.x * (blockDim.x >> 5);
    for (unsigned e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < count; e += warps) {
        const unsigned idx = kink_ws[1 + e];
        const unsigned m = idx / (unsigned)N, n = idx - m * (unsigned)N;
        const int ox = (int)(m % (unsigned)Wo), t = (int)(m / (unsigned)Wo), oy = t % Ho, b = t / Ho;
        const float* wr = w + (size_t)n * ldw;
        double acc = 0.0;
        for (int tap = 0; tap < k * k; ++tap) {
            const int kh = tap / k, kw = tap - kh * k;
            const float* xp = x + (((size_t)b * H + (size_t)s * oy + kh) * W + (size_t)s * ox + kw) * Cin;
            for (int c = lane; c < Cin; c += 32) acc += (double)xp[c] * (double)wr[tap * Cin + c];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            if (bias) acc += (double)bias[n];
            C[(size_t)m * ldc + n] = acc > 0.0 ? (float)acc : 0.0f;
        }
    }
}

// C[m][n] = sum_s ws[s][m][n] (+ bias[n]), fixed order.
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long mn, int N, const float* __restrict__ bias,
                                     float* __restrict__ C, int ldc) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < mn; i += (long long)gridDim.x * blockDim.x) {
        // the partial sums are added in split order (reproducible); the loads of eight splits are issued together — a thread
        // that waits for one L2 round trip per split made every one of these 21 launches per step a 15 us latency chain
        float acc = ws[i];
        int s = 1;
        for (; s + 8 <= splits; s += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldcs(ws + (long long)(s + j) * mn + i);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[j];
        }
        for (; s < splits; ++s) acc += ws[(long long)s * mn + i];
        const int n = (int)(i % N);
        const long long m = i / N;
        if (bias) acc += bias[n];
        C[m * ldc + n] = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// 2-D fp32 tensor [outer][ld] with `inner` valid contiguous elements per row; box = {box_inner (32 floats = 128 B), box_outer}.
static bool make_map(CUtensorMap* map, const float* ptr, int inner, int outer, int ld, int box_inner, int box_outer, bool kmajor) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              kmajor ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn encode_im2col_fn() {
    static EncodeIm2colFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeIm2colFn>(f);
    }();
    return fn;
}

// channels-last x [B][H][W][C]; k x k window, stride s, no padding: base pixels live in [0, W - (k - 1)) x [0, H - (k - 1)),
// the filter tap is added by the instruction's offsets.  32 channels x `pixels` output pixels per load.
static bool make_im2col_map(CUtensorMap* map, const float* x, int B, int H, int W, int C, int k, int s, int pixels, bool kmajor_tile,
                            int lower_w = 0, int lower_h = 0, int upper_w = 1 << 30, int upper_h = 1 << 30) {
    EncodeIm2colFn fn = encode_im2col_fn();
    if (!fn) return false;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    int lower[2] = {lower_w, lower_h}, upper[2] = {upper_w == (1 << 30) ? -(k - 1) : upper_w, upper_h == (1 << 30) ? -(k - 1) : upper_h};
    cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, lower, upper, 32, (cuuint32_t)pixels, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, kmajor_tile ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, bool A_K, bool B_K>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2, const Params& p, cudaStream_t stream) {
    static size_t cache[kMaxDevices] = {};
    auto kern = gemm3x_kernel<BN, A_K, B_K>;
    cudaError_t e = ensure_dynamic_smem(kern, Cfg<BN>::kSmem, cache);
    if (e != cudaSuccess) return (int)e;
    const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
    const int n_work = m_tiles * n_tiles * p.splits;
    kern<<<n_work < kSMs ? n_work : kSMs, kThreads, Cfg<BN>::kSmem, stream>>>(ma, mb, mb2, p);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

template <int BN>
static int launch_major(bool a_k, bool b_k, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2, const Params& p,
                        cudaStream_t st) {
    if (a_k) return b_k ? launch<BN, true, true>(ma, mb, mb2, p, st) : launch<BN, true, false>(ma, mb, mb2, p, st);
    return b_k ? launch<BN, false, true>(ma, mb, mb2, p, st) : launch<BN, false, false>(ma, mb, mb2, p, st);
}

static int launch_bn(int bn, bool a_k, bool b_k, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2, const Params& p,
                     cudaStream_t st) {
    switch (bn) {
        case 224: return launch_major<224>(a_k, b_k, ma, mb, mb2, p, st);
        case 256: return launch_major<256>(a_k, b_k, ma, mb, mb2, p, st);
        case 128: return launch_major<128>(a_k, b_k, ma, mb, mb2, p, st);
        default: return launch_major<64>(a_k, b_k, ma, mb, mb2, p, st);
    }
}

// x -> TF32 hi (nearest, ties away) and lo = TF32(x - hi), exactly what the splitter warps of gemm3x_kernel compute
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = src[i];
        const uint32_t h = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
        const uint32_t l = (__float_as_uint(x - __uint_as_float(h)) + 0x1000u) & 0xffffe000u;
        hi[i] = __uint_as_float(h);
        lo[i] = __uint_as_float(l);
    }
}

}  // namespace gemm
}  // namespace spair

using namespace spair;

extern "C" int spair_gemm_block_n(int N, int b_kmajor) {
    // the N tile: a divisor-friendly width for the decoder's G*G*(C+1) = 1568 = 7 * 224 columns, otherwise the widest tile
    // whose padding waste stays small.  MN-major B tiles are built from 32-column swizzle atoms.
    if (N % 224 == 0) return 224;
    if (N > 128 && (N % 256 == 0 || N % 256 > 128)) return 256;
    if (N > 64) return 128;
    (void)b_kmajor;
    return 64;
}

extern "C" int spair_gemm_splits(int M, int N, int K) {
    const int bn = spair_gemm_block_n(N, 1);
    const long long tiles = (long long)((M + gemm::BM - 1) / gemm::BM) * ((N + bn - 1) / bn);
    if (tiles >= kSMs / 2 || K <= 4 * gemm::BK) return 1;
    long long s = (2 * kSMs + tiles - 1) / tiles;          // about two work items per SM
    const long long max_s = (K + 8 * gemm::BK - 1) / (8 * gemm::BK);   // at least 8 k-blocks per split
    if (s > max_s) s = max_s;
    return (int)(s < 1 ? 1 : s);
}

extern "C" int spair_split_tf32(const float* src, float* hi, float* lo, int n, void* stream) {
    SPAIR_REQUIRE(src && hi && lo && n > 0);
    const int grid = (int)((n + 255) / 256 < 4 * kSMs ? (n + 255) / 256 : 4 * kSMs);
    gemm::split_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, hi, lo, n);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_gemm3x(const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor, float* C, int ldc, int M,
                            int N, int K, const float* bias, int epilogue, int period, float s_colour, float s_alpha, float b_alpha,
                            float* workspace, int splits, unsigned* kink_ws, int kink_cap, const float* B_hi, const float* B_lo, void* stream) {
    SPAIR_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && splits >= 1);
    SPAIR_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0);                                   // TMA: 16-byte row pitch
    SPAIR_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0);
    SPAIR_REQUIRE(epilogue >= SPAIR_GEMM_EPI_NONE && epilogue <= SPAIR_GEMM_EPI_TEXEL);
    SPAIR_REQUIRE(epilogue != SPAIR_GEMM_EPI_TEXEL || period >= 2);
    SPAIR_REQUIRE(splits == 1 || (workspace != nullptr && epilogue == SPAIR_GEMM_EPI_NONE));
    const int bn = spair_gemm_block_n(N, b_kmajor);
    SPAIR_REQUIRE(((uintptr_t)B_hi & 15) == 0 && ((uintptr_t)B_lo & 15) == 0 && (B_hi == nullptr) == (B_lo == nullptr));
    CUtensorMap ma, mb, mb2;
    const float* Bm = B_hi ? B_hi : B;       // B itself stays the fp32 matrix (the exact-ReLU re-evaluation reads it)
    bool ok = a_kmajor ? gemm::make_map(&ma, A, K, M, lda, gemm::BK, gemm::BM, true) : gemm::make_map(&ma, A, M, K, lda, 32, gemm::BK, false);
    ok = ok && (b_kmajor ? gemm::make_map(&mb, Bm, K, N, ldb, gemm::BK, bn, true) : gemm::make_map(&mb, Bm, N, K, ldb, 32, gemm::BK, false));
    mb2 = mb;
    if (B_lo) ok = ok && (b_kmajor ? gemm::make_map(&mb2, B_lo, K, N, ldb, gemm::BK, bn, true) : gemm::make_map(&mb2, B_lo, N, K, ldb, 32, gemm::BK, false));
    SPAIR_REQUIRE(ok);
    gemm::Params p;
    p.b_presplit = B_lo != nullptr;
    p.M = M; p.N = N; p.K = K;
    // the caller sizes the workspace for `splits`; the k-blocks are dealt out evenly and splits that would be empty are dropped
    const int kb_total = (K + gemm::BK - 1) / gemm::BK;
    const int kb_per = (kb_total + splits - 1) / splits;
    splits = (kb_total + kb_per - 1) / kb_per;
    p.splits = splits;
    p.k_per_split = kb_per * gemm::BK;
    p.C = splits == 1 ? C : workspace;
    p.ldc = splits == 1 ? ldc : N;
    p.bias = bias;
    p.epilogue = epilogue;
    p.period = period;
    p.s_colour = s_colour; p.s_alpha = s_alpha; p.b_alpha = b_alpha;
    p.conv = 0; p.cv_Ho = p.cv_Wo = p.cv_k = p.cv_s = p.cv_cpb = 1;
    p.cv_base = 0; p.om_s = p.om_H = p.om_W = p.om_py = p.om_px = 0;
    // forward layers with a long reduction whose ReLU decisions matter: split accumulators (see the MMA warp)
    p.acc_split = (epilogue == SPAIR_GEMM_EPI_RELU && bn <= 128 && splits == 1 && K >= 512) ? 4 : 1;
    const bool fixup = kink_ws != nullptr && kink_cap > 0 && epilogue == SPAIR_GEMM_EPI_RELU;
    SPAIR_REQUIRE(!fixup || (a_kmajor && b_kmajor && splits == 1 && (long long)M * N < (1ll << 32)));
    p.kink_ws = fixup ? kink_ws : nullptr;
    p.kink_cap = kink_cap;
    const char* dbg = getenv("SPAIR_GEMM_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (fixup) {
        cudaError_t e = cudaMemsetAsync(kink_ws, 0, sizeof(unsigned), st);
        if (e != cudaSuccess) return (int)e;
    }
    int rc = gemm::launch_bn(bn, a_kmajor, b_kmajor, ma, mb, mb2, p, st);
    if (rc == 0 && fixup) {
        gemm::relu_fixup_kernel<<<2 * kSMs, 256, 0, st>>>(A, lda, B, ldb, bias, C, ldc, N, K, kink_ws, kink_cap);
        SPAIR_LAUNCH_CHECK();
    }
    if (rc != 0 || splits == 1) return rc;
    const long long mn = (long long)M * N;
    gemm::splitk_reduce_kernel<<<grid_for(mn, 256) < 4 * kSMs ? grid_for(mn, 256) : 4 * kSMs, 256, 0, st>>>(workspace, splits, mn, N, bias, C, ldc);
    SPAIR_LAUNCH_CHECK();
}

// Implicit-GEMM convolution of the backbone tail (reference modules.py:44-66) on a channels-last input, no patch matrix in HBM:
//   mode 1 (forward)         y[B*Ho*Wo, Cout] = act(patches(x) . w^T + bias),  w = [Cout][(kh, kw, c)]
//   mode 2 (weight gradient) dw[Cout][(kh, kw, c)] = dy^T . patches(x),        dy = [B*Ho*Wo, Cout]
// patches(x)[m][(kh*k + kw)*C + c] = x[b][s*oy + kh][s*ox + kw][c] is read tile by tile with TMA im2col loads.
extern "C" int spair_conv_gemm3x(const float* x, int B, int H, int W, int C, int k, int stride, int mode, const float* other,
                                 int ld_other, float* out, int ldc, int Cout, const float* bias, int epilogue, float* workspace,
                                 int splits, unsigned* kink_ws, int kink_cap, const float* w_hi, const float* w_lo, void* stream) {
    SPAIR_REQUIRE(x && other && out && B > 0 && H >= k && W >= k && C > 0 && (C & 31) == 0 && k > 0 && k <= 8 && stride > 0 && stride <= 8);
    SPAIR_REQUIRE((mode == 1 || mode == 2) && Cout > 0 && splits >= 1 && (ld_other & 3) == 0);
    SPAIR_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)other & 15) == 0);
    const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
    const long long pixels = (long long)B * Ho * Wo;
    SPAIR_REQUIRE(pixels < (1ll << 31));
    const int KK = k * k * C;
    gemm::Params p;
    CUtensorMap ma, mb, mb2;
    int bn;
    bool ok;
    p.b_presplit = 0;
    if (mode == 1) {
        SPAIR_REQUIRE(epilogue == SPAIR_GEMM_EPI_NONE || epilogue == SPAIR_GEMM_EPI_RELU);
        SPAIR_REQUIRE(splits == 1 && ((uintptr_t)w_hi & 15) == 0 && ((uintptr_t)w_lo & 15) == 0 && (w_hi == nullptr) == (w_lo == nullptr));
        p.M = (int)pixels; p.N = Cout; p.K = KK;
        bn = spair_gemm_block_n(p.N, 1);
        // w_hi / w_lo: the TF32 planes of the weights (spair_split_tf32); `other` stays the fp32 matrix for the exact-ReLU pass
        ok = gemm::make_im2col_map(&ma, x, B, H, W, C, k, stride, gemm::BM, true) &&
             gemm::make_map(&mb, w_hi ? w_hi : other, KK, Cout, ld_other, gemm::BK, bn, true);
        if (w_lo) {
            ok = ok && gemm::make_map(&mb2, w_lo, KK, Cout, ld_other, gemm::BK, bn, true);
            p.b_presplit = 1;
        }
    } else {
        SPAIR_REQUIRE(w_hi == nullptr && w_lo == nullptr);
        SPAIR_REQUIRE(epilogue == SPAIR_GEMM_EPI_NONE && bias == nullptr && (splits == 1 || workspace != nullptr));
        p.M = Cout; p.N = KK; p.K = (int)pixels;
        bn = spair_gemm_block_n(p.N, 0);
        ok = gemm::make_map(&ma, other, Cout, (int)pixels, ld_other, 32, gemm::BK, false) && gemm::make_im2col_map(&mb, x, B, H, W, C, k, stride, 32, false);
    }
    SPAIR_REQUIRE(ok);
    if (!p.b_presplit) mb2 = mb;
    const int kb_total = (p.K + gemm::BK - 1) / gemm::BK;
    const int kb_per = (kb_total + splits - 1) / splits;
    splits = (kb_total + kb_per - 1) / kb_per;
    p.splits = splits;
    p.k_per_split = kb_per * gemm::BK;
    p.C = splits == 1 ? out : workspace;
    p.ldc = splits == 1 ? ldc : p.N;
    p.bias = bias;
    p.epilogue = epilogue;
    p.period = 2; p.s_colour = p.s_alpha = 1.0f; p.b_alpha = 0.0f;
    p.conv = mode; p.cv_Ho = Ho; p.cv_Wo = Wo; p.cv_k = k; p.cv_s = stride; p.cv_cpb = C / 32;
    p.cv_base = 0; p.om_s = p.om_H = p.om_W = p.om_py = p.om_px = 0;
    p.acc_split = (mode == 1 && epilogue == SPAIR_GEMM_EPI_RELU && bn <= 128 && p.K >= 512) ? 4 : 1;
    const bool fixup = mode == 1 && kink_ws != nullptr && kink_cap > 0 && epilogue == SPAIR_GEMM_EPI_RELU && (long long)p.M * p.N < (1ll << 32);
    p.kink_ws = fixup ? kink_ws : nullptr;
    p.kink_cap = kink_cap;
    const char* dbg = getenv("SPAIR_GEMM_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool a_k = mode == 1, b_k = mode == 1;
    if (fixup) {
        cudaError_t e = cudaMemsetAsync(kink_ws, 0, sizeof(unsigned), st);
        if (e != cudaSuccess) return (int)e;
    }
    int rc = gemm::launch_bn(bn, a_k, b_k, ma, mb, mb2, p, st);
    if (rc == 0 && fixup) {
        gemm::relu_fixup_conv_kernel<<<2 * kSMs, 256, 0, st>>>(x, H, W, C, k, stride, Ho, Wo, other, ld_other, bias, out, ldc, Cout, kink_ws, kink_cap);
        SPAIR_LAUNCH_CHECK();
    }
    if (rc != 0 || splits == 1) return rc;
    const long long mn = (long long)p.M * p.N;
    gemm::splitk_reduce_kernel<<<grid_for(mn, 256) < 4 * kSMs ? grid_for(mn, 256) : 4 * kSMs, 256, 0, st>>>(workspace, splits, mn, p.N, nullptr, out, ldc);
    SPAIR_LAUNCH_CHECK();
}

// Input gradient of the same convolution without a d_col matrix: dx[b][s*i+py][s*j+px][c] = sum over the T x T taps (T = k / s)
// of dy[b][i-a][j-a'][:] . w[:, c, py + s*a, px + s*a'] — s*s stride-1 sub-convolutions over dy (zero padded by T-1 through the
// TMA bounding box), one GEMM per output parity class, whose rows the epilogue scatters to their pixels of dx.
//   dy [B,Ho,Wo,Cout] channels-last (Cout % 32 == 0);  wc: s*s packed class weights [Cin][(a', a, co) = T*T*Cout] with
//   wc[cls = py*s+px][c][((T-1-a)*T + (T-1-a'))*Cout + co] = w[co][c][py + s*a][px + s*a'];  dx [B,H,W,Cin] (fully overwritten).
extern "C" int spair_conv_dgrad3x(const float* dy, int B, int H, int W, int Cin, int k, int stride, int Cout, const float* wc,
                                  float* dx, const float* wc_hi, const float* wc_lo, void* stream) {
    SPAIR_REQUIRE(dy && wc && dx && B > 0 && H >= k && W >= k && Cin > 0 && Cout > 0 && (Cout & 31) == 0 && k > 0 && stride > 0);
    SPAIR_REQUIRE(k % stride == 0 && k / stride <= 8 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)wc & 15) == 0);
    SPAIR_REQUIRE(((uintptr_t)wc_hi & 15) == 0 && ((uintptr_t)wc_lo & 15) == 0 && (wc_hi == nullptr) == (wc_lo == nullptr));
    const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1, T = k / stride, KK = T * T * Cout;
    // rows / columns of x beyond the last window receive no gradient: the caller gets zeros there from the class loop below
    // only if every pixel belongs to some class row, i.e. Hp covers all of H
    cudaStream_t st = (cudaStream_t)stream;
    const int bn = spair_gemm_block_n(Cin, 1);
    for (int py = 0; py < stride; ++py)
        for (int px = 0; px < stride; ++px) {
            const int Hp = (H - py + stride - 1) / stride, Wp = (W - px + stride - 1) / stride;
            if (Hp <= 0 || Wp <= 0) continue;
            gemm::Params p;
            p.M = B * Hp * Wp; p.N = Cin; p.K = KK;
            p.splits = 1; p.k_per_split = ((KK + gemm::BK - 1) / gemm::BK) * gemm::BK;
            p.C = dx; p.ldc = Cin; p.bias = nullptr; p.epilogue = SPAIR_GEMM_EPI_NONE;
            p.period = 2; p.s_colour = p.s_alpha = 1.0f; p.b_alpha = 0.0f;
            p.conv = 1; p.cv_Ho = Hp; p.cv_Wo = Wp; p.cv_k = T; p.cv_s = 1; p.cv_cpb = Cout / 32; p.cv_base = -(T - 1);
            p.om_s = stride; p.om_H = H; p.om_W = W; p.om_py = py; p.om_px = px;
            p.acc_split = 1; p.kink_ws = nullptr; p.kink_cap = 0;
            p.b_presplit = wc_lo != nullptr;      // wc_hi / wc_lo: the TF32 planes of the class weights (spair_split_tf32)
            const char* dbg = getenv("SPAIR_GEMM_DEBUG");
            p.debug = dbg ? atoi(dbg) : 0;
            CUtensorMap ma, mb, mb2;
            // base positions of the T-wide window over dy: [-(T-1), Wp - (T-1)) -> upper corner = Wp - Wo - (T-1)
            bool ok = gemm::make_im2col_map(&ma, dy, B, Ho, Wo, Cout, T, 1, gemm::BM, true, -(T - 1), -(T - 1), Wp - Wo - (T - 1), Hp - Ho - (T - 1));
            ok = ok && gemm::make_map(&mb, (wc_hi ? wc_hi : wc) + (size_t)(py * stride + px) * Cin * KK, KK, Cin, KK, gemm::BK, bn, true);
            mb2 = mb;
            if (wc_lo) ok = ok && gemm::make_map(&mb2, wc_lo + (size_t)(py * stride + px) * Cin * KK, KK, Cin, KK, gemm::BK, bn, true);
            SPAIR_REQUIRE(ok);
            const int rc = gemm::launch_bn(bn, true, true, ma, mb, mb2, p, st);
            if (rc != 0) return rc;
        }
    return 0;
}
