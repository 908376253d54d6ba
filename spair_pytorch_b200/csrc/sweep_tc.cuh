// Tensor-core dense layers of the fused cell sweep (csrc/sweep.cu; reference modules.py:124-165 evaluated inside the
// autoregressive loop of models.py:68-117).
//
// The SIMT sweep streams every weight matrix from L2 per wavefront into FMAs that reuse a weight for <= 16 rows.  Here the
// same CTA (16 worker warps, <= 16 rows per wavefront) hands its dense layers to the 5th-generation tensor cores:
//   * the WEIGHTS are the M-side operand A (128 output features per tile), the activation rows the N-side operand B, so
//     a tcgen05.mma costs what its 16 / 32 activation rows cost, not a padded 128-row tile; D[feature][row] lives in TMEM;
//   * fp32 accuracy by operand splitting (x = hi + lo, both TF32): B holds the hi rows (0-15) and the lo rows (16-31) of
//     the activations side by side, so ONE MMA of N = 32 with W_hi yields W_hi x_hi and W_hi x_lo in two column groups and
//     a second MMA of N = 16 with W_lo adds W_lo x_hi — two instructions per k-step instead of three, and the hi*lo terms
//     accumulate apart from the hi*hi sum;
//   * weights are pre-split and pre-swizzled ONCE per step by tc_pack_kernel into the exact shared-memory image of the
//     operand tiles (128 x 32 floats, 128-byte swizzle, hi then lo = 32 KB per stage), laid out in consumption order: a
//     producer thread streams the whole sequence with plain bulk async copies (cp.async.bulk, SASS UBLKCP) through a
//     3-stage mbarrier ring and runs ahead ACROSS layer boundaries, so a layer never starts with a cold L2 round trip;
//   * two MMA-issuer threads walk the static layer plan, alternating over the k-blocks of a layer (the issue interval of a
//     tcgen05.mma, not its execution, is what a stage costs); the workers convert their activations into B tiles (first
//     layer: from the global input rows; hidden layers: straight from the epilogue registers) and hand them over through
//     a 3-slot ring; tcgen05.commit frees weight stages / activation slots and publishes the accumulators.
// Hidden activations never exist in row-major form on chip: the epilogue of layer l (tcgen05.ld -> bias -> ReLU / mask)
// writes layer l+1's swizzled hi / lo operand rows and the global copy the weight-gradient GEMMs need.
#pragma once
#include "common.cuh"

#ifndef SW_MARK      // phase-timing hooks of csrc/sweep.cu (-DSW_TIMING)
#define SW_T0()
#define SW_MARK(i)
#endif

namespace spair {
namespace tc {

#ifdef SW_TIMING
// diagnostics of instrumented builds only (results are garbage): 1 = the producer signals stages without copying,
// 2 = the MMA threads commit without issuing MMAs (spair_debug_sweep_tc_flags; tools/sweep_phase_timing.py TC_DEBUG_FLAGS)
__device__ int g_tc_debug;
#define TC_DEBUG(flag) (g_tc_debug & (flag))
#else
#define TC_DEBUG(flag) 0
#endif

constexpr int kWorkers = 512;            // threads 0..511: the sweep's own 16 warps
constexpr int kIssuers = 2;              // MMA-issuing threads (one per warp): a thread issues one tcgen05.mma per ~67 cycles whatever
                                         // its shape, a second thread raises the rate to one per ~46 (tools/probes/umma_latency.cu).
                                         // Issuer i takes the k-blocks (= weight stages) with index % kIssuers == i of every layer
constexpr int kAccs = kIssuers;          // ... and adds them into its own accumulator of the tile.  The tensor core adds into TMEM with
                                         // truncation, so the error grows with the chain length; the epilogue sums the partial
                                         // accumulators with round-to-nearest adds
constexpr int kProducerWarp = 16, kMmaWarp = 17;   // warp 16: bulk-copy producer; warps 17 .. 17 + kIssuers - 1: MMA issuers
constexpr int kThreads = 32 * (kMmaWarp + kIssuers);
constexpr int kRows = 16;                // activation rows per CTA and wavefront (== kSwRows)
constexpr int kStageBytes = 32768;       // one weight stage: hi tile (128 features x 128 B) + lo tile
constexpr int kStageFloats = kStageBytes / 4;
constexpr int kWStages = 3;
constexpr int kChunkKB = 8;              // k-blocks per activation chunk
constexpr int kSlotBytes = kChunkKB * 4096;   // one activation chunk: 8 k-blocks x (16 hi + 16 lo rows) x 128 B
constexpr int kXSlots = 3;
constexpr int kChunkK = kChunkKB * 32;   // reduction indices per activation chunk
constexpr int kTmemCols = 512;
constexpr int kTileCols = 32 * kAccs;    // every accumulator: 16 columns hi*hi + lo*hi, 16 columns hi*lo
constexpr int kTileGroup = kTmemCols / kTileCols;   // 128-feature tiles per accumulator group (8)
constexpr int kMaxLayers = 12;
constexpr int kRingBytes = kWStages * kStageBytes + kXSlots * kSlotBytes;
constexpr int kBarBytes = 256;
static_assert(8 * (2 * kWStages + 2 * kXSlots + 3) <= kBarBytes, "barrier slots");

// Layers in execution order: K = reduction length, M = output features.
struct Plan {
    int n_layers;
    int K[kMaxLayers], M[kMaxLayers];
};

__host__ __device__ inline int stages_of(int K, int M) { return ((M + 127) / 128) * ((K + 31) / 32); }

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
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
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
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// barrier of the 512 worker threads only (the producer / MMA warps never join it)
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// K-major SWIZZLE_128B operand descriptor (same encoding as csrc/gemm.cu: 8-row groups 1024 B apart, version 1, layout 2)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(16u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// D = f32, A = B = tf32, both K-major, M = 128
__host__ __device__ constexpr uint32_t instr_desc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ uint32_t tf32_round(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_round(x);
    lo = tf32_round(x - __uint_as_float(hi));     // x - hi is exact in fp32
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// barrier slots (8 bytes each) behind the rings
struct Bars {
    uint32_t base;
    __device__ __forceinline__ uint32_t w_full(int s) const { return base + 8 * s; }
    __device__ __forceinline__ uint32_t w_empty(int s) const { return base + 8 * (kWStages + s); }
    __device__ __forceinline__ uint32_t x_full(int s) const { return base + 8 * (2 * kWStages + s); }
    __device__ __forceinline__ uint32_t x_empty(int s) const { return base + 8 * (2 * kWStages + kXSlots + s); }
    __device__ __forceinline__ uint32_t acc_full() const { return base + 8 * (2 * kWStages + 2 * kXSlots); }
    __device__ __forceinline__ uint32_t acc_empty() const { return base + 8 * (2 * kWStages + 2 * kXSlots + 1); }
    __device__ __forceinline__ uint32_t tmem_slot() const { return base + 8 * (2 * kWStages + 2 * kXSlots + 2); }
};

__device__ __forceinline__ void init_barriers(const Bars b) {
    for (int s = 0; s < kWStages; ++s) { mbar_init(b.w_full(s), 1); mbar_init(b.w_empty(s), kIssuers); }
    for (int s = 0; s < kXSlots; ++s) { mbar_init(b.x_full(s), kWorkers / 32); mbar_init(b.x_empty(s), kIssuers); }
    mbar_init(b.acc_full(), kIssuers);
    mbar_init(b.acc_empty(), kWorkers / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// ---- producer: one thread streams the packed weight sequence of a wavefront, once per wavefront --------------------------
__device__ __forceinline__ void producer_loop(uint32_t wring, const Bars b, const float* __restrict__ wstream, int stages_per_wavefront,
                                              int n_wavefronts) {
    int ws = 0;
    uint32_t ph = 0;
    for (int t = 0; t < n_wavefronts; ++t) {
        const float* src = wstream;
        for (int s = 0; s < stages_per_wavefront; ++s, src += kStageFloats) {
            mbar_wait(b.w_empty(ws), ph ^ 1);
            if (TC_DEBUG(1)) {
                mbar_arrive(b.w_full(ws));
            } else {
                mbar_expect_tx(b.w_full(ws), kStageBytes);
                bulk_copy(wring + ws * kStageBytes, src, kStageBytes, b.w_full(ws));
            }
            if (++ws == kWStages) { ws = 0; ph ^= 1; }
        }
    }
}

// ---- MMA issuers: kIssuers threads (one per warp) walk the layer plan of every wavefront ---------------------------------
// The MMAs are short (N = 32 / 16) and a thread can issue one per ~67 cycles whatever its shape, so the issue slots are
// the critical resource: the k-blocks of a layer (= weight stages: 4 k-steps of two MMAs, hi pass N = 32 + lo pass
// N = 16) alternate between the issuers, each adding into its own 32 TMEM columns of the tile.  A stage is waited for and
// released by its issuer only; chunks and accumulator groups are committed by every issuer (barrier counts = kIssuers).
// Descriptors are formed by adding small constants to per-ring-position bases (the start-address field is the low 14 bits of
// the descriptor: + 2 = one k-step of 32 bytes, + 1024 = the lo tile 16 KB behind the hi tile, + 256 = the next 4 KB
// k-block of the activations).
constexpr uint64_t kDescBase = ((uint64_t)(16u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);

template <int ISS>
__device__ __forceinline__ void mma_loop(uint32_t wring, uint32_t xring, const Bars b, uint32_t tmem, const Plan& plan, int n_wavefronts) {
    constexpr uint32_t idesc32 = instr_desc(32), idesc16 = instr_desc(16);
    const uint64_t a_base = kDescBase | (uint64_t)((wring & 0x3FFFFu) >> 4);
    const uint64_t x_base = kDescBase | (uint64_t)((xring & 0x3FFFFu) >> 4);
    uint64_t a_desc = a_base;
    int ws = 0;
    uint32_t wph = 0, acc_it = 0, x_cons = 0;      // x_cons: activation chunks consumed so far, modulo 2 * kXSlots
    for (int t = 0; t < n_wavefronts; ++t) {
        for (int l = 0; l < plan.n_layers; ++l) {
            const int K = plan.K[l], M = plan.M[l];
            const int KB = (K + 31) >> 5, chunks = (K + kChunkK - 1) / kChunkK, tiles = (M + 127) >> 7;
            // one accumulator group: every chunk is released as soon as its MMAs are issued.  Several groups (more than
            // kTileGroup feature tiles; the host guarantees chunks <= kXSlots then): the chunks stay until the last group.
            const bool hold = tiles > kTileGroup;
            for (int g0 = 0; g0 < tiles; g0 += kTileGroup) {
                const int gt = min(kTileGroup, tiles - g0);
                const bool last_group = g0 + kTileGroup >= tiles;
                mbar_wait(b.acc_empty(), (acc_it & 1) ^ 1);      // the workers have read the previous accumulator group
                fence_after();
                for (int c = 0; c < chunks; ++c) {
                    const uint32_t xc = (x_cons + (hold ? c : 0)) % (2 * kXSlots);
                    const uint32_t xs = xc % kXSlots, xph = xc / kXSlots;
                    if (g0 == 0) {
                        mbar_wait(b.x_full(xs), xph);
                        fence_after();
                    }
                    const uint64_t x_desc = x_base + xs * (kSlotBytes >> 4);
                    const int kbn = min(kChunkKB, KB - kChunkKB * c);
                    const int k_left = K - kChunkK * c;          // reduction indices from the start of this chunk
                    for (int mt = 0; mt < gt; ++mt) {
                        const uint32_t d = tmem + mt * kTileCols + 32 * ISS;
                        uint64_t b_desc = x_desc;
                        for (int kbl = 0; kbl < kbn; ++kbl, b_desc += 256) {
                            // (kChunkKB is a multiple of kIssuers: the owner of a k-block follows from its index in the chunk)
                            // every issuer observes AND releases every stage: the phase of an mbarrier is one bit, so a thread that
                            // skipped a round of a slot would mistake another round's completion for the one it waits for
                            mbar_wait(b.w_full(ws), wph);
                            if (kbl % kIssuers == ISS) {
                                fence_after();
                                const int ks_n = (k_left - 32 * kbl + 7) >> 3;     // >= 4: a full k-block
                                if (!TC_DEBUG(2)) {
                                    umma_tf32(d, a_desc, b_desc, idesc32, (c != 0) | (kbl >= kIssuers));
                                    umma_tf32(d, a_desc + 1024, b_desc, idesc16, 1);
                                    if (ks_n > 1) {
                                        umma_tf32(d, a_desc + 2, b_desc + 2, idesc32, 1);
                                        umma_tf32(d, a_desc + 1026, b_desc + 2, idesc16, 1);
                                    }
                                    if (ks_n > 2) {
                                        umma_tf32(d, a_desc + 4, b_desc + 4, idesc32, 1);
                                        umma_tf32(d, a_desc + 1028, b_desc + 4, idesc16, 1);
                                    }
                                    if (ks_n > 3) {
                                        umma_tf32(d, a_desc + 6, b_desc + 6, idesc32, 1);
                                        umma_tf32(d, a_desc + 1030, b_desc + 6, idesc16, 1);
                                    }
                                }
                                umma_commit(b.w_empty(ws));      // stage may be refilled once these MMAs have read it ...
                            } else {
                                mbar_arrive(b.w_empty(ws));      // ... and every other issuer has seen this round of the slot
                            }
                            a_desc += kStageBytes >> 4;
                            if (++ws == kWStages) { ws = 0; wph ^= 1; a_desc = a_base; }
                        }
                    }
                    if (!hold) {
                        umma_commit(b.x_empty(xs));
                        x_cons = (x_cons + 1) % (2 * kXSlots);
                    }
                }
                if (hold && last_group) {
                    for (int c = 0; c < chunks; ++c) umma_commit(b.x_empty((x_cons + c) % kXSlots));
                    x_cons = (x_cons + chunks) % (2 * kXSlots);
                }
                umma_commit(b.acc_full());
                ++acc_it;
            }
        }
    }
}
static_assert(kChunkKB % kIssuers == 0, "k-block ownership by index in the chunk");

// ---- workers ---------------------------------------------------------------------------------------------------------------
struct Worker {
    uint32_t xring;
    Bars b;
    uint32_t tmem;
    uint32_t x_prod;     // activation chunks produced so far
    uint32_t acc_cnt;    // accumulator groups consumed so far

    __device__ __forceinline__ uint32_t acquire_slot() {
        const uint32_t slot = x_prod % kXSlots, ph = (x_prod / kXSlots) & 1;
        mbar_wait(b.x_empty(slot), ph ^ 1);
        return slot;
    }
    __device__ __forceinline__ void publish_slot(uint32_t slot) {
        fence_async_smem();          // generic-proxy writes -> visible to the MMA's async-proxy reads
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(b.x_full(slot));
        ++x_prod;
    }
};

// byte offset of reduction index k (0..31 inside its k-block) of activation row r inside a k-block tile (128-byte swizzle)
__device__ __forceinline__ uint32_t b_offset(int row, int k) { return row * 128 + ((((k >> 2) ^ row) & 7) << 4) + ((k & 3) << 2); }

// First layer: the input rows live in memory (global X rows, or the row-major shared buffer the heads wrote): warp r
// converts row r into hi / lo operand rows, one chunk of 256 reduction indices at a time.  Rows >= nrows and the padding
// up to the next multiple of 32 are written as zero (a NaN there would poison valid columns through 0 * NaN).
__device__ __forceinline__ void stage_rows(Worker& w, const float* __restrict__ src_row, bool valid, int K) {
    const int row = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k0 = 0; k0 < K; k0 += 2 * kChunkK) {       // two chunks per pass: eight loads in flight per thread
        float v[2 * kChunkKB];
#pragma unroll
        for (int i = 0; i < 2 * kChunkKB; ++i) {
            const int k = k0 + 32 * i + lane;
            v[i] = (valid && k < K) ? src_row[k] : 0.0f;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int kc = k0 + h * kChunkK;
            if (kc < K) {
                const int kbn = min(kChunkKB, (K - kc + 31) >> 5);
                const uint32_t slot = w.acquire_slot();
                const uint32_t base = w.xring + slot * kSlotBytes + b_offset(row, lane);
#pragma unroll
                for (int i = 0; i < kChunkKB; ++i) {
                    if (i < kbn) {
                        uint32_t hi, lo;
                        split_tf32(v[h * kChunkKB + i], hi, lo);
                        st_shared_u32(base + i * 4096, hi);
                        st_shared_u32(base + i * 4096 + kRows * 128, lo);
                    }
                }
                w.publish_slot(slot);
            }
        }
    }
}

// fp32 operands of a ReLU layer for the exact re-evaluation of pre-activations near zero: nn.Linear weight [M][K] and the
// layer's input rows in global memory (row pitch ld_x); w == nullptr: no re-evaluation
struct Kink {
    const float* w;
    const float* x;
    int ld_x;
};
constexpr float kKinkTol = 1.0e-5f;   // absolute: 4x the measured split-precision error of these layers (<= 2.5e-6 on
                                      // pre-activations of magnitude ~1)

// Epilogue of one layer with M output features: out[r][f] = act(D[f][r] + D[f][16 + r] + bias[f]) (* mask).  Warp w reads
// TMEM lane quarter w % 4 (features) for the rows 4 (w / 4) .. + 3; a lane owns one feature, so the global stores of a row
// are 128-byte segments and the operand stores of a row hit 32 distinct banks.
//   Hmask    : [rows][M] forward activation whose sign gates the gradient (backward hidden layers), or nullptr
//   next_b   : write the result as the next layer's operand chunk (hidden layers)
//   y_out    : row-major shared copy [r][ld_y] for the head phases (last layer of an MLP), or nullptr
__device__ __forceinline__ void epilogue(Worker& w, int K, int M, const float* __restrict__ bias, bool relu, const float* __restrict__ Hmask,
                                         const int* __restrict__ grow, int nrows, float* __restrict__ out_glob, int ld_out, bool next_b,
                                         float* y_out, int ld_y, const Kink kink, int mark) {
    SW_T0();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3, r0 = (warp >> 2) * 4;
    const int tiles = (M + 127) >> 7, n_acc = min(kAccs, (K + 31) >> 5);
    int grow_r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) grow_r[i] = grow[r0 + i];
    float bz_n = 0.0f, mk_n[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    auto preload = [&](int t) {
        const int f = t * 128 + q * 32 + lane;
        bz_n = (bias && f < M) ? __ldg(bias + f) : 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) mk_n[i] = (Hmask && f < M && r0 + i < nrows) ? Hmask[(size_t)grow_r[i] * M + f] : 1.0f;
    };
    preload(0);
    for (int g0 = 0; g0 < tiles; g0 += kTileGroup) {
        const int gt = min(kTileGroup, tiles - g0);
        mbar_wait(w.b.acc_full(), w.acc_cnt & 1);
        fence_after();
        SW_MARK(mark);
        uint32_t slot = 0;
        if (next_b) slot = w.acquire_slot();
#pragma unroll 1
        for (int tl = 0; tl < gt; ++tl) {
            const int t = g0 + tl;
            const float bz = bz_n;
            float mk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) mk[i] = mk_n[i];
            if (t + 1 < tiles && (bias || Hmask)) preload(t + 1);
            // the accumulators that received a k-block in this layer (K <= 32 (kAccs - 1): fewer), hi*hi and hi*lo column
            // groups of each, added in a fixed order
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            const uint32_t taddr = w.tmem + ((uint32_t)(q * 32) << 16) + tl * kTileCols + r0;
#pragma unroll
            for (int j = 0; j < kAccs; ++j) {
                if (j < n_acc) {
                    uint32_t a[4], bb[4];
                    tmem_ld4(taddr + 32 * j, a);
                    tmem_ld4(taddr + 32 * j + kRows, bb);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i] += __uint_as_float(a[i]) + __uint_as_float(bb[i]);
                }
            }
            const int f = t * 128 + q * 32 + lane;
            const bool fv = f < M;
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += bz;
            if (relu && kink.w != nullptr) {       // warp-uniform
                // A ReLU whose pre-activation is within the split-precision rounding error of zero could take the other branch
                // than fp32 arithmetic does; every gradient through this unit would then differ (the forward value would not:
                // it is ~0 either way).  Such elements — |v| <= kKinkTol, a few per 100,000 — are re-evaluated by the whole
                // warp from the fp32 operands with float64 accumulation.
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    unsigned pending = __ballot_sync(0xffffffffu, fv && r0 + i < nrows && fabsf(acc[i]) <= kKinkTol);
                    while (pending) {
                        const int src = __ffs(pending) - 1;
                        pending &= pending - 1;
                        const int fs = t * 128 + q * 32 + src;
                        const float* __restrict__ wrow = kink.w + (size_t)fs * K;
                        const float* xrow = kink.x + (size_t)grow_r[i] * kink.ld_x;
                        double sum = 0.0;
                        for (int k = lane; k < K; k += 32) sum += (double)__ldg(wrow + k) * (double)__ldcg(xrow + k);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                        if (lane == src) acc[i] = (float)(sum + (double)bz);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + i;
                float v = acc[i];
                if (relu) v = fmaxf(v, 0.0f);
                if (!(mk[i] > 0.0f)) v = 0.0f;
                if (!fv) v = 0.0f;
                if (fv && r < nrows) out_glob[(size_t)grow_r[i] * ld_out + f] = v;
                if (next_b) {
                    uint32_t hi, lo;
                    split_tf32(r < nrows ? v : 0.0f, hi, lo);
                    const uint32_t addr = w.xring + slot * kSlotBytes + (t * 4 + q) * 4096 + b_offset(r, lane);
                    st_shared_u32(addr, hi);
                    st_shared_u32(addr + kRows * 128, lo);
                }
                if (y_out && fv) y_out[r * ld_y + f] = (r < nrows) ? v : 0.0f;
            }
        }
        fence_before();
        if (next_b) w.publish_slot(slot);      // M <= 256 = one chunk of the next layer
        __syncwarp();
        if (lane == 0) mbar_arrive(w.b.acc_empty());
        ++w.acc_cnt;
        SW_MARK(mark + 1);
    }
    (void)mark;
}

}  // namespace tc
}  // namespace spair
