// Coordinate arithmetic shared by the glimpse, paste and render kernels.
//
// The reference samples through F.affine_grid + F.grid_sample (modules.py:265-269) with
// align_corners=False.  The functions below restate torch's fp32 op ORDER for that pair —
// verified bit-exact against torch 2.11 CPU by an fp32 emulation (DESIGN.md "coordinate
// parity") — using explicit round-to-nearest intrinsics so that nvcc cannot contract the
// multiplies and adds into FMAs where torch does not:
//   base_j = linspace(-1,1,n)[j] * (n-1) / n        linspace element = fma(step, j, -1) for j < n/2,
//                                                    fma(-step, n-1-j, 1) otherwise, step = 2/(n-1)
//   g      = fl(fl(base * a) + c)                    affine_grid's bmm (a = theta[0,0], c = theta[0,2])
//   ix     = fl(fl(fl(g + 1) * (size/2)) - 0.5)      grid_sampler unnormalize
//   out    = fma(v11, se, fma(v10, sw, fma(v01, ne, v00*nw)))   with nw = (x0+1-ix)*(y0+1-iy), ...
#pragma once
#include "common.cuh"

namespace spair {

__host__ __device__ __forceinline__ float base_coord(int j, int n) {
    if (n <= 1) return 0.0f;
#ifdef __CUDA_ARCH__
    const float step = __fdiv_rn(2.0f, (float)(n - 1));
    float v = (j < n / 2) ? fmaf(step, (float)j, -1.0f) : fmaf(-step, (float)(n - 1 - j), 1.0f);
    return __fdiv_rn(__fmul_rn(v, (float)(n - 1)), (float)n);
#else
    const float step = 2.0f / (float)(n - 1);
    float v = (j < n / 2) ? __builtin_fmaf(step, (float)j, -1.0f) : __builtin_fmaf(-step, (float)(n - 1 - j), 1.0f);
    volatile float m = v * (float)(n - 1);
    return m / (float)n;
#endif
}

// normalised grid coordinate -> source pixel coordinate (align_corners=False)
__device__ __forceinline__ float unnormalize(float g, float half_size) {
    return __fadd_rn(__fmul_rn(__fadd_rn(g, 1.0f), half_size), -0.5f);
}

__device__ __forceinline__ float affine_coord(float base, float a, float c) { return __fadd_rn(__fmul_rn(base, a), c); }

// forward direction (modules.py:243-253): theta = [[xs,0,2xt-1],[0,ys,2yt-1]]
struct FwdAffine {
    float ax, cx, ay, cy;
    __device__ __forceinline__ FwdAffine(float xt, float yt, float xs, float ys) {
        ax = xs;
        ay = ys;
        cx = __fadd_rn(__fmul_rn(xt, 2.0f), -1.0f);
        cy = __fadd_rn(__fmul_rn(yt, 2.0f), -1.0f);
    }
};

// inverse direction (modules.py:256-262): Tensor.inverse() of [[xs,0,xt'],[0,ys,yt'],[0,0,1]] gives
// theta = [[1/xs, 0, -(xt'/xs)], [0, 1/ys, -(yt'/ys)]] (divisions, checked bit-exact against torch CPU)
struct InvAffine {
    float ax, cx, ay, cy;
    __device__ __forceinline__ InvAffine(float xt, float yt, float xs, float ys) {
        const float xtp = __fadd_rn(__fmul_rn(xt, 2.0f), -1.0f), ytp = __fadd_rn(__fmul_rn(yt, 2.0f), -1.0f);
        ax = __fdiv_rn(1.0f, xs);
        ay = __fdiv_rn(1.0f, ys);
        cx = -__fdiv_rn(xtp, xs);
        cy = -__fdiv_rn(ytp, ys);
    }
};

}  // namespace spair
