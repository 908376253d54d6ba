// K: KL terms.  Replaces SPAIR._compute_KL (reference models.py:169-262).
//
// Forward: one CTA per image.  Warp 0 runs the sequential count-prior scan with the running count
// distribution (HW+1 entries) held in registers, striped over the 32 lanes (two butterfly reductions
// per cell instead of the reference's bmm + sum + five host syncs per cell); the remaining warps
// evaluate the z_pres-masked Normal KLs (coalesced over the [HW, D] block of the image).  Per-name sums
// (the torch.sum(z_kl, dim=[1,2,3]) of models.py:553) come out of one deterministic block reduction.
// Backward: elementwise, one warp per (image, cell) so that d_z_pres is a warp reduction.
#include "common.cuh"

namespace spair {

constexpr int kKLThreads = 256;

__device__ __forceinline__ float normal_kl(float mean, float std_, float pm, float ps) {
    // torch.distributions.kl._kl_normal_normal
    const float r = std_ / ps;
    const float var_ratio = r * r;
    const float d = (mean - pm) / ps;
    return 0.5f * (var_ratio + d * d - 1.0f - logf(var_ratio));
}

__device__ __forceinline__ int kl_name(int j, int A) { return j < 4 ? j : (j < 4 + A ? 4 : 5); }

template <int NJ>
__device__ __forceinline__ float pres_scan(const float* __restrict__ pres_b, const float* __restrict__ cd0, int HW,
                                           float* __restrict__ kl_out, int kl_stride, float* __restrict__ pz_out) {
    const int lane = threadIdx.x & 31;
    float cd[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int c = lane + 32 * j;
        cd[j] = c <= HW ? cd0[c] : 0.0f;
    }
    float count = 0.0f, total = 0.0f;
    for (int i = 0; i < HW; ++i) {
        const float R = (float)(HW - i);
        float q[NJ];
        float pz = 0.0f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float c = (float)(lane + 32 * j);
            q[j] = fminf(fmaxf(c - count, 0.0f), R) / R;                   // models.py:206
            pz = fmaf(cd[j], q[j], pz);                                    // models.py:214
        }
        pz = warp_sum(pz);
        const float pi = pres_b[i];
        const float kl = pi * (logf(pi + 1e-9f) - logf(pz + 1e-9f)) +
                         (1.0f - pi) * (logf(1.0f - pi + 1e-9f) - logf(1.0f - pz + 1e-9f));   // models.py:223-226
        if (lane == 0) {
            kl_out[(size_t)i * kl_stride] = kl;
            pz_out[i] = pz;
        }
        total += kl;
        const float s = rintf(pi);                                         // torch.round: half to even, models.py:232
        float norm = 0.0f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            cd[j] = (s * q[j] + (1.0f - s) * (1.0f - q[j])) * cd[j];       // models.py:234-237
            norm += cd[j];
        }
        norm = fmaxf(warp_sum(norm), 1e-6f);                               // models.py:238
#pragma unroll
        for (int j = 0; j < NJ; ++j) cd[j] = cd[j] / norm;
        count += s;
    }
    return total;
}

__global__ void __launch_bounds__(kKLThreads)
kl_fwd_kernel(const float* __restrict__ dmean, const float* __restrict__ dstd, const float* __restrict__ pres,
              const float* __restrict__ prior_mean, const float* __restrict__ prior_std,
              const float* __restrict__ cd0, int HW, int A, float* __restrict__ kl_map, float* __restrict__ p_z,
              float* __restrict__ kl_sums) {
    __shared__ float red[7 * (kKLThreads / 32)];
    const int b = blockIdx.x;
    const int D = 4 + A + 1;
    const int warp = threadIdx.x >> 5;
    float sums[7] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    const float* pres_b = pres + (size_t)b * HW;
    if (warp == 0) {
        float* klo = kl_map + (size_t)b * HW * (D + 1) + D;
        float* pzo = p_z + (size_t)b * HW;
        float t;
        if (HW + 1 <= 32 * 4) t = pres_scan<4>(pres_b, cd0, HW, klo, D + 1, pzo);
        else if (HW + 1 <= 32 * 9) t = pres_scan<9>(pres_b, cd0, HW, klo, D + 1, pzo);
        else t = pres_scan<33>(pres_b, cd0, HW, klo, D + 1, pzo);
        if ((threadIdx.x & 31) == 0) sums[6] = t;
    } else {
        const int total = HW * D;
        const float* m = dmean + (size_t)b * total;
        const float* s = dstd + (size_t)b * total;
        for (int e = threadIdx.x - 32; e < total; e += kKLThreads - 32) {
            const int cell = e / D, j = e - cell * D;
            const float kl = pres_b[cell] * normal_kl(m[e], s[e], prior_mean[j], prior_std[j]);   // models.py:175-177
            kl_map[((size_t)b * HW + cell) * (D + 1) + j] = kl;
            const int nm = kl_name(j, A);
#pragma unroll
            for (int k = 0; k < 6; ++k) sums[k] += (nm == k) ? kl : 0.0f;
        }
    }
    block_sum<7>(sums, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) kl_sums[(size_t)b * 7 + k] = sums[k];
    }
}

__global__ void __launch_bounds__(kKLThreads)
kl_bwd_kernel(const float* __restrict__ dmean, const float* __restrict__ dstd, const float* __restrict__ pres,
              const float* __restrict__ prior_mean, const float* __restrict__ prior_std,
              const float* __restrict__ p_z, const float* __restrict__ d_sums, int B, int HW, int A,
              float* __restrict__ d_dmean, float* __restrict__ d_dstd, float* __restrict__ d_pres) {
    const int D = 4 + A + 1;
    const int lane = threadIdx.x & 31;
    const size_t n_obj = (size_t)B * HW;
    for (size_t o = (size_t)blockIdx.x * (kKLThreads / 32) + (threadIdx.x >> 5); o < n_obj;
         o += (size_t)gridDim.x * (kKLThreads / 32)) {
        const size_t b = o / HW;
        const float pi = pres[o];
        const float* ds = d_sums + b * 7;
        float dp = 0.0f;
        for (int j = lane; j < D; j += 32) {
            const float mean = dmean[o * D + j], std_ = dstd[o * D + j], pm = prior_mean[j], ps = prior_std[j];
            const float g = ds[kl_name(j, A)];
            dp = fmaf(g, normal_kl(mean, std_, pm, ps), dp);
            d_dmean[o * D + j] = g * pi * (mean - pm) / (ps * ps);
            d_dstd[o * D + j] = g * pi * (std_ / (ps * ps) - 1.0f / std_);
        }
        dp = warp_sum(dp);
        if (lane == 0) {
            const float pz = p_z[o];
            const float e = 1e-9f;
            const float dkl = logf(pi + e) - logf(pz + e) + pi / (pi + e) - logf(1.0f - pi + e) + logf(1.0f - pz + e) -
                              (1.0f - pi) / (1.0f - pi + e);
            d_pres[o] = dp + ds[6] * dkl;
        }
    }
}

}  // namespace spair

using namespace spair;

extern "C" int spair_kl_fwd(const float* dmean, const float* dstd, const float* pres, const float* prior_mean,
                            const float* prior_std, const float* count_dist0, int B, int HW, int A, float* kl_map,
                            float* p_z, float* kl_sums, void* stream) {
    SPAIR_REQUIRE(dmean && dstd && pres && prior_mean && prior_std && count_dist0 && kl_map && p_z && kl_sums);
    SPAIR_REQUIRE(B > 0 && HW > 0 && HW + 1 <= 32 * 33 && A > 0);
    kl_fwd_kernel<<<B, kKLThreads, 0, (cudaStream_t)stream>>>(dmean, dstd, pres, prior_mean, prior_std, count_dist0, HW,
                                                              A, kl_map, p_z, kl_sums);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_kl_bwd(const float* dmean, const float* dstd, const float* pres, const float* prior_mean,
                            const float* prior_std, const float* kl_map, const float* p_z, const float* d_sums, int B,
                            int HW, int A, float* d_dmean, float* d_dstd, float* d_pres, void* stream) {
    (void)kl_map;
    SPAIR_REQUIRE(dmean && dstd && pres && prior_mean && prior_std && p_z && d_sums && d_dmean && d_dstd && d_pres);
    SPAIR_REQUIRE(B > 0 && HW > 0 && A > 0);
    int grid = grid_for((long long)B * HW, kKLThreads / 32);
    if (grid > kSMs * 8) grid = kSMs * 8;
    kl_bwd_kernel<<<grid, kKLThreads, 0, (cudaStream_t)stream>>>(dmean, dstd, pres, prior_mean, prior_std, p_z, d_sums,
                                                                 B, HW, A, d_dmean, d_dstd, d_pres);
    SPAIR_LAUNCH_CHECK();
}
