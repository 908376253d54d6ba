// Probe: latency / throughput of short tcgen05.mma.kind::tf32 sequences (M = 128, K = 8, shared-memory operands) as the
// fused sweep issues them.  One CTA, one issuing thread; times clock64 from the first issue to the observed completion
// of the tcgen05.commit mbarrier.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_latency umma_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
constexpr uint64_t kDescBase = ((uint64_t)(16u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
__host__ __device__ constexpr uint32_t idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// mode 0: all MMAs accumulate into the same D; 1: D rotates over 4 column groups; 2: every MMA reads the same A tile too;
// 3: straight-line groups of 8 MMAs (descriptor = base + constant), as csrc/sweep_tc.cuh issues them; 4: mode 3 issued by TWO
// warps at the same time (n_mma each, separate accumulators and barriers) — does the issue rate scale with issuing threads?
__global__ void probe(long long* out, int n_mma, int N, int mode, int reps) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
    const uint32_t bar = base + 160 * 1024, slot = bar + 8;
    for (int i = threadIdx.x; i < 40 * 1024; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.0f;
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (mode >= 3) {
        if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < (mode == 4 ? 2 : 1)) {
            const int w = threadIdx.x >> 5;
            const uint32_t mybar = bar + 16 + 8 * w;
            mbar_init(mybar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint64_t a0 = kDescBase | (((base + w * 65536) & 0x3FFFFu) >> 4), b0 = kDescBase | (((base + 128 * 1024) & 0x3FFFFu) >> 4);
            const uint32_t id = idesc(N), id2 = idesc(N / 2), d = tmem + w * 256;
            long long best = 1ll << 60, best_issue = 1ll << 60;
            uint32_t ph = 0;
            for (int r = 0; r < reps; ++r) {
                const long long t0 = clock64();
                uint64_t a = a0;
                for (int i = 0; i < n_mma; i += 8, a += 2048) {
                    umma(d, a, b0, id, i > 0);
                    umma(d, a + 1024, b0, id2, 1);
                    umma(d, a + 2, b0 + 2, id, 1);
                    umma(d, a + 1026, b0 + 2, id2, 1);
                    umma(d, a + 4, b0 + 4, id, 1);
                    umma(d, a + 1028, b0 + 4, id2, 1);
                    umma(d, a + 6, b0 + 6, id, 1);
                    umma(d, a + 1030, b0 + 6, id2, 1);
                    if (a == a0 + 3 * 2048) a = a0 - 2048;
                }
                commit(mybar);
                const long long t1 = clock64();
                mbar_wait(mybar, ph);
                const long long t2 = clock64();
                ph ^= 1;
                if (t2 - t0 < best) best = t2 - t0;
                if (t1 - t0 < best_issue) best_issue = t1 - t0;
            }
            out[2 * w] = best; out[2 * w + 1] = best_issue;
        }
    } else if (threadIdx.x == 0) {
        const uint64_t a0 = kDescBase | ((base & 0x3FFFFu) >> 4), b0 = kDescBase | (((base + 128 * 1024) & 0x3FFFFu) >> 4);
        const uint32_t id = idesc(N);
        long long best = 1ll << 60, best_issue = 1ll << 60;
        uint32_t ph = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i) {
                const uint64_t a = a0 + (mode == 2 ? 0 : ((i & 3) * 2 + ((i >> 2) & 3) * 2048));   // k-steps of 4 stages
                const uint32_t d = tmem + (mode == 1 ? (i & 3) * 64 : 0);
                umma(d, a, b0 + (i & 3) * 2, id, i > 3);
            }
            commit(bar);
            const long long t1 = clock64();
            mbar_wait(bar, ph);
            const long long t2 = clock64();
            ph ^= 1;
            if (t2 - t0 < best) best = t2 - t0;
            if (t1 - t0 < best_issue) best_issue = t1 - t0;
        }
        out[0] = best; out[1] = best_issue;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

int main() {
    long long* d; cudaMalloc(&d, 32);
    const int smem = 162 * 1024 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    printf("mode N n_mma : cycles issue->complete (issue only)\n");
    for (int mode = 2; mode < 5; ++mode)
        for (int N : {32, 128, 256})
            for (int n : {8, 16, 64}) {
                probe<<<1, 128, smem>>>(d, n, N, mode, 20);
                long long h[4] = {0, 0, 0, 0}; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
                cudaError_t e = cudaGetLastError();
                printf("%d %3d %3d : %6lld (%lld)  second warp %6lld (%lld) %s\n", mode, N, n, h[0], h[1], mode == 4 ? h[2] : 0ll, mode == 4 ? h[3] : 0ll, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
