// Probe: throughput of mma.sync.m16n8k8 TF32 (HMMA.1688.F32.TF32) and of TMA bulk copies from L2 on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CHAINS>
__global__ void mma_kernel(float* out, int iters) {
    float acc[CHAINS][4];
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f900000u, 0x3fa00000u, 0x3fb00000u}, b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[c][e] = 0.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) mma_tf32(acc[c], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// every CTA streams the SAME `total` bytes from global (L2 resident) through a ring of `stages` x `stage_bytes`
__global__ void __launch_bounds__(64) tma_kernel(const float* src, size_t total, int stage_bytes, int stages, float* out, int rotate) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm);
    uint64_t* empty = full + 16;
    unsigned char* ring = sm + 256;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(full + s)), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(empty + s)), "r"(1));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t n_stage = total / stage_bytes;
    const size_t start = rotate ? (blockIdx.x * 7919u) % n_stage : 0;
    float acc = 0.f;
    if (threadIdx.x == 32) {            // producer
        for (size_t it = 0; it < n_stage; ++it) {
            const uint32_t slot = it % stages, ph = (it / stages) & 1;
            uint32_t done = 0;
            while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(empty + slot)), "r"(ph ^ 1) : "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full + slot)), "r"(stage_bytes) : "memory");
            const unsigned char* g = reinterpret_cast<const unsigned char*>(src) + ((it + start) % n_stage) * stage_bytes;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + (size_t)slot * stage_bytes)),
                         "l"(g), "r"(stage_bytes), "r"(smem_u32(full + slot)) : "memory");
        }
    } else if (threadIdx.x < 32) {      // consumer warp: touch one word per stage, release
        for (size_t it = 0; it < n_stage; ++it) {
            const uint32_t slot = it % stages, ph = (it / stages) & 1;
            uint32_t done = 0;
            while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(full + slot)), "r"(ph) : "memory");
            acc += reinterpret_cast<float*>(ring + (size_t)slot * stage_bytes)[threadIdx.x];
            __syncwarp();
            if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + slot)) : "memory");
        }
    }
    if (threadIdx.x < 32) out[blockIdx.x * 32 + threadIdx.x] = acc;
}

int main() {
    float* out;
    cudaMalloc(&out, 1 << 24);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps : {4, 8, 16}) {
        for (int chains : {1, 2, 4, 8}) {
            auto launch = [&]() {
                if (chains == 1) mma_kernel<1><<<148, warps * 32>>>(out, iters);
                else if (chains == 2) mma_kernel<2><<<148, warps * 32>>>(out, iters);
                else if (chains == 4) mma_kernel<4><<<148, warps * 32>>>(out, iters);
                else mma_kernel<8><<<148, warps * 32>>>(out, iters);
            };
            launch();
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double mmas = 148.0 * warps * chains * iters;
            printf("mma tf32 m16n8k8: warps/SM %2d chains %d : %.3f ms  %.1f TFLOP/s  (%.2f clk/MMA/SM at 1.9 GHz)\n", warps, chains, ms,
                   mmas * 2048 / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.9e9 / (warps * chains * (double)iters));
        }
    }
    // TMA bulk stream: 2 MB (L2 resident) read by every CTA, 16 passes
    float* src;
    const size_t total = 2u << 20;
    cudaMalloc(&src, total);
    cudaMemset(src, 0, total);
    for (int ctas : {1, 32, 128, 148}) {
        for (int stage_kb : {4, 16, 32}) {
            for (int stages : {2, 4, 6}) {
                for (int rotate : {0, 1}) {
                    const int sb = stage_kb * 1024;
                    const size_t smem = 256 + (size_t)stages * sb;
                    if (smem > 200 * 1024) continue;
                    cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    tma_kernel<<<ctas, 64, smem>>>(src, total, sb, stages, out, rotate);
                    cudaEventRecord(e0);
                    for (int r = 0; r < 8; ++r) tma_kernel<<<ctas, 64, smem>>>(src, total, sb, stages, out, rotate);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    ms /= 8;
                    printf("tma bulk: ctas %3d stage %2d KB x %d rotate %d : %.3f ms  %.1f GB/s per CTA  %.2f TB/s total\n", ctas, stage_kb, stages, rotate, ms,
                           total / (ms * 1e-3) / 1e9, ctas * (double)total / (ms * 1e-3) / 1e12);
                }
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
