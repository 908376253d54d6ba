// Shared device helpers for libspair_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spair_b200.h"

#define SPAIR_LAUNCH_CHECK()                       \
    do {                                           \
        cudaError_t e__ = cudaGetLastError();      \
        return e__ == cudaSuccess ? 0 : (int)e__;  \
    } while (0)

#define SPAIR_REQUIRE(cond)                        \
    do {                                           \
        if (!(cond)) return SPAIR_ERR_INVALID;     \
    } while (0)

namespace spair {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// torch.clamp(x, -10, 10) and the mask its backward applies (inclusive bounds).
__device__ __forceinline__ float clamp10(float x) { return fminf(fmaxf(x, -10.0f), 10.0f); }
__device__ __forceinline__ float clamp10_mask(float x) { return (x >= -10.0f && x <= 10.0f) ? 1.0f : 0.0f; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV values per thread; result valid in thread 0.  `red` is shared scratch of
// at least NV * (blockDim.x / 32) floats.  Deterministic (fixed tree).
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * nwarp + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float s = lane < nwarp ? red[i * nwarp + lane] : 0.0f;
            v[i] = warp_sum(s);
        }
    }
}

// exact floor(i / d) for 0 <= i < 2^15, 1 <= d <= 2^10 without an integer division (inv_d = 1.0f / d):
// (i + 0.5) / d is at least 0.5/d away from an integer, far more than the fp32 rounding of the product
__device__ __forceinline__ int fast_div(int i, float inv_d) { return (int)(((float)i + 0.5f) * inv_d); }

__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }

struct NeighbourList {
    int n;
    int dh[SPAIR_MAX_NEIGHBOURS];
    int dw[SPAIR_MAX_NEIGHBOURS];
};

// Opt-in to more than 48 KB of dynamic shared memory for `func`, remembered PER DEVICE (the attribute is per device and
// per function: a process-wide flag would leave a second GPU of the same process without it).  `cache` is a
// zero-initialised static array owned by the call site, one slot per device ordinal; the attribute only ever grows.
constexpr int kMaxDevices = 64;
template <typename F>
inline cudaError_t ensure_dynamic_smem(F func, size_t bytes, size_t (&cache)[kMaxDevices]) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (bytes > cache[dev]) {
        e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        cache[dev] = bytes;
    }
    return cudaSuccess;
}

inline int grid_for(long long work, int block) {
    long long g = (work + block - 1) / block;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace spair
