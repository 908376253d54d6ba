// Per-cell latent heads and lateral-context kernels.
// All of these are tiny elementwise kernels over one wavefront of cells: a wavefront holds at
// most ceil(Wc/2) cells x B images, so they are launch/latency bound by construction; the work
// per launch is laid out one thread per output element with coalesced row-major stores.
#include "common.cuh"

namespace spair {

// ------------------------------------------------------------------------------------------
// L0: context gather (reference models.py:292-320 + the cat at models.py:76)
// ------------------------------------------------------------------------------------------
__global__ void context_gather_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ box,
                                          const float* __restrict__ attr, const float* __restrict__ depth,
                                          const float* __restrict__ pres, const float* __restrict__ edge,
                                          const int* __restrict__ cells, int n_cells, NeighbourList nb, int B, int F,
                                          int Hc, int Wc, int A, float* __restrict__ dst0, int ld0,
                                          float* __restrict__ dst1, int ld1, float* __restrict__ dst2, int ld2) {
    const int E = A + 6;
    const int width = F + nb.n * E;
    const long long total = (long long)n_cells * B * width;
    const int HW = Hc * Wc;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(idx % width);
        const int r = (int)(idx / width);
        const int b = r % B;
        const int cell = cells[r / B];
        const int h = cell / Wc, w = cell % Wc;
        float v;
        if (col < F) {
            v = feat[((long long)(b * F + col) * Hc + h) * Wc + w];
        } else {
            const int s = (col - F) / E, j = (col - F) % E;
            const int nh = h + nb.dh[s], nw = w + nb.dw[s];
            if (nh >= 0 && nh < Hc && nw >= 0 && nw < Wc) {
                const long long o = (long long)b * HW + nh * Wc + nw;
                if (j < 4) v = box[o * 4 + j];
                else if (j < 4 + A) v = attr[o * A + (j - 4)];
                else if (j == 4 + A) v = depth[o];
                else v = pres[o];
            } else {
                v = edge[j];
            }
        }
        if (dst0) dst0[(long long)r * ld0 + col] = v;
        if (dst1) dst1[(long long)r * ld1 + col] = v;
        if (dst2) dst2[(long long)r * ld2 + col] = v;
    }
}

__global__ void context_grad_gather_kernel(const float* __restrict__ dx0, int ld0, const float* __restrict__ dx1,
                                           int ld1, const float* __restrict__ dx2, int ld2, int col0,
                                           const int* __restrict__ cells, int n_cells,
                                           const int* __restrict__ wf_pos, NeighbourList nb, int B, int Hc, int Wc,
                                           int A, float* __restrict__ out, int ld_out) {
    const int E = A + 6;
    const long long total = (long long)n_cells * B * E;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % E);
        const int r = (int)(idx / E);
        const int b = r % B;
        const int cell = cells[r / B];
        const int h = cell / Wc, w = cell % Wc;
        float acc = 0.0f;
        for (int s = 0; s < nb.n; ++s) {
            // the consumer whose s-th neighbour is this cell
            const int ch = h - nb.dh[s], cw = w - nb.dw[s];
            if (ch < 0 || ch >= Hc || cw < 0 || cw >= Wc) continue;
            const long long cr = (long long)wf_pos[ch * Wc + cw] * B + b;
            const int c = col0 + s * E + j;
            if (dx0) acc += dx0[cr * ld0 + c];
            if (dx1) acc += dx1[cr * ld1 + c];
            if (dx2) acc += dx2[cr * ld2 + c];
        }
        out[(long long)r * ld_out + j] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// L1+L2: box head (reference models.py:322-381)
// component k of the latent: 0 = cy, 1 = cx, 2 = height, 3 = width (models.py:330-336)
// box = [cell_x, cell_y, width, height]; z_where = [xt, yt, xs, ys] (models.py:361,376)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int box_slot(int k) { return k == 0 ? 1 : (k == 1 ? 0 : (k == 2 ? 3 : 2)); }

__global__ void box_head_fwd_kernel(const float* __restrict__ y, int ld_y, const float* __restrict__ eps,
                                    const int* __restrict__ cells, int n_cells, int B, int HW, int Wc,
                                    spair_box_geom g, float* __restrict__ box, float* __restrict__ z_where,
                                    float* __restrict__ dmean, float* __restrict__ dstd, int ld_dist,
                                    float* __restrict__ xdst0, int ldx0, float* __restrict__ xdst1, int ldx1,
                                    int n_pt, float* __restrict__ pt_dst, int ld_pt) {
    const int per_row = 4 + n_pt;
    const long long total = (long long)n_cells * B * per_row;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % per_row);
        const int r = (int)(idx / per_row);
        const float* yr = y + (long long)r * ld_y;
        if (k >= 4) {  // passthrough features -> next network's input (models.py:88)
            pt_dst[(long long)r * ld_pt + (k - 4)] = yr[8 + (k - 4)];
            continue;
        }
        const int b = r % B;
        const int cell = cells[r / B];
        const long long o = (long long)b * HW + cell;
        const float mean = yr[k];
        const float std_ = sigmoid_f(clamp10(yr[4 + k])) * 2.0f;           // modules.py:175
        const float z = mean + eps[o * 4 + k] * std_;                       // Normal.rsample, models.py:450
        const float s = sigmoid_f(clamp10(z));                              // clamped_sigmoid, modules.py:189
        float bval, zval;
        if (k < 2) {
            bval = g.yx_scale * s + g.yx_min;                               // models.py:346-347
            const float pos = (k == 0) ? (float)(cell / Wc) : (float)(cell % Wc);
            zval = (k == 0 ? g.cell_ratio_y : g.cell_ratio_x) * (bval + pos);  // models.py:373-374
        } else {
            bval = g.hw_scale * s + g.hw_min;                               // models.py:358-359
            zval = bval * g.anchor / (k == 2 ? g.img_h : g.img_w);          // models.py:369-370
        }
        const int slot = box_slot(k);
        box[o * 4 + slot] = bval;
        z_where[o * 4 + slot] = zval;
        dmean[o * ld_dist + k] = mean;
        dstd[o * ld_dist + k] = std_;
        if (xdst0) xdst0[(long long)r * ldx0 + slot] = bval;
        if (xdst1) xdst1[(long long)r * ldx1 + slot] = bval;
    }
}

__global__ void box_head_bwd_kernel(const float* __restrict__ y, int ld_y, const float* __restrict__ eps,
                                    const int* __restrict__ cells, int n_cells, int B, int HW, int Wc,
                                    spair_box_geom g, const float* __restrict__ wheel,
                                    const float* __restrict__ d_box0, int ldb0, const float* __restrict__ d_box1,
                                    int ldb1, const float* __restrict__ d_box2, int ldb2,
                                    const float* __restrict__ d_zw_local, int ld_zwl,
                                    const float* __restrict__ d_zw_img, const float* __restrict__ d_dmean,
                                    const float* __restrict__ d_dstd, int ld_dist, int n_pt,
                                    const float* __restrict__ d_pt_src, int ld_pt, float* __restrict__ d_y) {
    const int per_row = 4 + n_pt;
    const long long total = (long long)n_cells * B * per_row;
    const float keep = 1.0f - wheel[0];                                     // models.py:425
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % per_row);
        const int r = (int)(idx / per_row);
        float* dyr = d_y + (long long)r * ld_y;
        if (k >= 4) {
            dyr[8 + (k - 4)] = d_pt_src ? d_pt_src[(long long)r * ld_pt + (k - 4)] : 0.0f;
            continue;
        }
        const float* yr = y + (long long)r * ld_y;
        const int b = r % B;
        const int cell = cells[r / B];
        const long long o = (long long)b * HW + cell;
        const int slot = box_slot(k);
        float d_b = 0.0f, d_z_where = 0.0f;
        if (d_box0) d_b += d_box0[(long long)r * ldb0 + slot];
        if (d_box1) d_b += d_box1[(long long)r * ldb1 + slot];
        if (d_box2) d_b += d_box2[(long long)r * ldb2 + slot];
        if (d_zw_local) d_z_where += d_zw_local[(long long)r * ld_zwl + slot];
        if (d_zw_img) d_z_where += d_zw_img[o * 4 + slot];
        const float ls = yr[4 + k];
        const float sg = sigmoid_f(clamp10(ls));
        const float std_ = sg * 2.0f;
        const float e = eps[o * 4 + k];
        const float z = yr[k] + e * std_;
        const float s = sigmoid_f(clamp10(z));
        float d_s;
        if (k < 2) {
            d_b += d_z_where * (k == 0 ? g.cell_ratio_y : g.cell_ratio_x);
            d_s = d_b * g.yx_scale;
        } else {
            d_b += (d_z_where / (k == 2 ? g.img_h : g.img_w)) * g.anchor;
            d_s = d_b * g.hw_scale;
        }
        const float d_zl = d_s * s * (1.0f - s) * clamp10_mask(z);
        float d_mean = d_zl, d_std = d_zl * e;
        if (d_dmean) d_mean += d_dmean[o * ld_dist + k];
        if (d_dstd) d_std += d_dstd[o * ld_dist + k];
        const float d_ls = d_std * 2.0f * sg * (1.0f - sg) * clamp10_mask(ls);
        dyr[k] = keep * d_mean;
        dyr[4 + k] = keep * d_ls;
    }
}

// ------------------------------------------------------------------------------------------
// L1/L3: Normal heads — attr (identity) and depth (4 * clamped sigmoid) (models.py:83-85,92-97)
// ------------------------------------------------------------------------------------------
__global__ void normal_head_fwd_kernel(const float* __restrict__ y, int ld_y, int W, const float* __restrict__ eps,
                                       const int* __restrict__ cells, int n_cells, int B, int HW, int squash,
                                       float squash_scale, float* __restrict__ out, float* __restrict__ dmean,
                                       float* __restrict__ dstd, int ld_dist, float* __restrict__ xdst0, int ldx0,
                                       float* __restrict__ xdst1, int ldx1, int n_pt, float* __restrict__ pt_dst,
                                       int ld_pt) {
    const int per_row = W + n_pt;
    const long long total = (long long)n_cells * B * per_row;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % per_row);
        const int r = (int)(idx / per_row);
        const float* yr = y + (long long)r * ld_y;
        if (k >= W) {
            pt_dst[(long long)r * ld_pt + (k - W)] = yr[2 * W + (k - W)];
            continue;
        }
        const int b = r % B;
        const long long o = (long long)b * HW + cells[r / B];
        const float mean = yr[k];
        const float std_ = sigmoid_f(clamp10(yr[W + k])) * 2.0f;
        float z = mean + eps[o * W + k] * std_;
        if (squash) z = squash_scale * sigmoid_f(clamp10(z));              // models.py:96
        out[o * W + k] = z;
        dmean[o * ld_dist + k] = mean;
        dstd[o * ld_dist + k] = std_;
        if (xdst0) xdst0[(long long)r * ldx0 + k] = z;
        if (xdst1) xdst1[(long long)r * ldx1 + k] = z;
    }
}

__global__ void normal_head_bwd_kernel(const float* __restrict__ y, int ld_y, int W, const float* __restrict__ eps,
                                       const int* __restrict__ cells, int n_cells, int B, int HW, int squash,
                                       float squash_scale, const float* __restrict__ wheel,
                                       const float* __restrict__ d_out0, int ldo0, const float* __restrict__ d_out1,
                                       int ldo1, const float* __restrict__ d_out2, int ldo2,
                                       const float* __restrict__ d_out_img, const float* __restrict__ d_dmean,
                                       const float* __restrict__ d_dstd, int ld_dist, int n_pt,
                                       const float* __restrict__ d_pt_src, int ld_pt, float* __restrict__ d_y) {
    const int per_row = W + n_pt;
    const long long total = (long long)n_cells * B * per_row;
    const float keep = wheel ? 1.0f - wheel[0] : 1.0f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % per_row);
        const int r = (int)(idx / per_row);
        float* dyr = d_y + (long long)r * ld_y;
        if (k >= W) {
            dyr[2 * W + (k - W)] = d_pt_src ? d_pt_src[(long long)r * ld_pt + (k - W)] : 0.0f;
            continue;
        }
        const float* yr = y + (long long)r * ld_y;
        const int b = r % B;
        const long long o = (long long)b * HW + cells[r / B];
        float d_o = 0.0f;
        if (d_out0) d_o += d_out0[(long long)r * ldo0 + k];
        if (d_out1) d_o += d_out1[(long long)r * ldo1 + k];
        if (d_out2) d_o += d_out2[(long long)r * ldo2 + k];
        if (d_out_img) d_o += d_out_img[o * W + k];
        const float ls = yr[W + k];
        const float sg = sigmoid_f(clamp10(ls));
        const float e = eps[o * W + k];
        float d_z = d_o;
        if (squash) {
            const float z = yr[k] + e * (sg * 2.0f);
            const float s = sigmoid_f(clamp10(z));
            d_z = d_o * squash_scale * s * (1.0f - s) * clamp10_mask(z);
        }
        float d_mean = d_z, d_std = d_z * e;
        if (d_dmean) d_mean += d_dmean[o * ld_dist + k];
        if (d_dstd) d_std += d_dstd[o * ld_dist + k];
        dyr[k] = keep * d_mean;
        dyr[W + k] = keep * (d_std * 2.0f * sg * (1.0f - sg) * clamp10_mask(ls));
    }
}

// ------------------------------------------------------------------------------------------
// L4: presence head (models.py:393-411)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float logistic_noise(float u) {
    const float eps = 10e-10f;                                             // models.py:401
    return logf(u + eps) - logf(1.0f - u + eps);                           // models.py:404
}

__global__ void pres_head_fwd_kernel(const float* __restrict__ y, int ld_y, const float* __restrict__ u,
                                     const int* __restrict__ cells, int n_cells, int B, int HW,
                                     float* __restrict__ pres) {
    const int total = n_cells * B;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += gridDim.x * blockDim.x) {
        const long long o = (long long)(r % B) * HW + cells[r / B];
        const float lo = clamp10(y[(long long)r * ld_y]);
        pres[o] = sigmoid_f((lo + logistic_noise(u[o])) / 1.0f);
    }
}

__global__ void pres_head_bwd_kernel(const float* __restrict__ y, int ld_y, const float* __restrict__ u,
                                     const int* __restrict__ cells, int n_cells, int B, int HW,
                                     const float* __restrict__ wheel, const float* __restrict__ d_local, int ld_local,
                                     const float* __restrict__ d_img, float* __restrict__ d_y) {
    const int total = n_cells * B;
    const float keep = 1.0f - wheel[0];
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += gridDim.x * blockDim.x) {
        const long long o = (long long)(r % B) * HW + cells[r / B];
        const float logit = y[(long long)r * ld_y];
        const float p = sigmoid_f(clamp10(logit) + logistic_noise(u[o]));
        float d = 0.0f;
        if (d_local) d += d_local[(long long)r * ld_local];
        if (d_img) d += d_img[o];
        d_y[(long long)r * ld_y] = keep * d * p * (1.0f - p) * clamp10_mask(logit);
    }
}

__global__ void relu_bwd_kernel(float* __restrict__ dh, int ld_dh, const float* __restrict__ h, int ld_h, int rows,
                                int cols) {
    const long long total = (long long)rows * cols;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cols);
        const long long r = idx / cols;
        if (!(h[r * ld_h + c] > 0.0f)) dh[r * ld_dh + c] = 0.0f;
    }
}

static bool make_neighbours(const int* nb_offsets, int n_nb, NeighbourList& nb) {
    if (!nb_offsets || n_nb < 1 || n_nb > SPAIR_MAX_NEIGHBOURS) return false;
    nb.n = n_nb;
    for (int i = 0; i < n_nb; ++i) {
        nb.dh[i] = nb_offsets[2 * i];
        nb.dw[i] = nb_offsets[2 * i + 1];
    }
    return true;
}

}  // namespace spair

using namespace spair;

extern "C" int spair_abi_version(void) { return SPAIR_ABI_VERSION; }

extern "C" int spair_context_gather_fwd(const float* feat, const float* box, const float* attr, const float* depth,
                                        const float* pres, const float* edge, const int* cells, int n_cells,
                                        const int* nb_offsets, int n_nb, int B, int F, int Hc, int Wc, int A,
                                        float* dst0, int ld0, float* dst1, int ld1, float* dst2, int ld2,
                                        void* stream) {
    NeighbourList nb;
    SPAIR_REQUIRE(feat && box && attr && depth && pres && edge && cells && n_cells > 0 && B > 0);
    SPAIR_REQUIRE(make_neighbours(nb_offsets, n_nb, nb));
    const long long total = (long long)n_cells * B * (F + n_nb * (A + 6));
    int grid = grid_for(total, 256);
    if (grid > kSMs * 8) grid = kSMs * 8;
    context_gather_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feat, box, attr, depth, pres, edge, cells,
                                                                      n_cells, nb, B, F, Hc, Wc, A, dst0, ld0, dst1,
                                                                      ld1, dst2, ld2);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_context_grad_gather(const float* dx0, int ld0, const float* dx1, int ld1, const float* dx2,
                                         int ld2, int col0, const int* cells, int n_cells, const int* wf_pos,
                                         const int* nb_offsets, int n_nb, int B, int Hc, int Wc, int A, float* out,
                                         int ld_out, void* stream) {
    NeighbourList nb;
    SPAIR_REQUIRE(cells && wf_pos && out && n_cells > 0 && B > 0 && ld_out >= A + 6);
    SPAIR_REQUIRE(make_neighbours(nb_offsets, n_nb, nb));
    const long long total = (long long)n_cells * B * (A + 6);
    int grid = grid_for(total, 256);
    if (grid > kSMs * 8) grid = kSMs * 8;
    context_grad_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dx0, ld0, dx1, ld1, dx2, ld2, col0, cells,
                                                                       n_cells, wf_pos, nb, B, Hc, Wc, A, out, ld_out);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_box_head_fwd(const float* y, int ld_y, const float* eps, const int* cells, int n_cells, int B,
                                  int HW, int Wc, const spair_box_geom* geom, float* box, float* z_where,
                                  float* dmean, float* dstd, int ld_dist, float* xdst0, int ldx0, float* xdst1,
                                  int ldx1, int n_pt, float* pt_dst, int ld_pt, void* stream) {
    SPAIR_REQUIRE(y && eps && cells && geom && box && z_where && dmean && dstd && n_cells > 0 && B > 0);
    SPAIR_REQUIRE(ld_y >= 8 + n_pt && (n_pt == 0 || pt_dst));
    const long long total = (long long)n_cells * B * (4 + n_pt);
    int grid = grid_for(total, 256);
    if (grid > kSMs * 8) grid = kSMs * 8;
    box_head_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, ld_y, eps, cells, n_cells, B, HW, Wc, *geom, box,
                                                                z_where, dmean, dstd, ld_dist, xdst0, ldx0, xdst1,
                                                                ldx1, n_pt, pt_dst, ld_pt);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_box_head_bwd(const float* y, int ld_y, const float* eps, const int* cells, int n_cells, int B,
                                  int HW, int Wc, const spair_box_geom* geom, const float* wheel, const float* d_box0,
                                  int ldb0, const float* d_box1, int ldb1, const float* d_box2, int ldb2,
                                  const float* d_zw_local, int ld_zwl, const float* d_zw_img, const float* d_dmean,
                                  const float* d_dstd, int ld_dist, int n_pt, const float* d_pt_src, int ld_pt,
                                  float* d_y, void* stream) {
    SPAIR_REQUIRE(y && eps && cells && geom && wheel && d_y && n_cells > 0 && B > 0 && ld_y >= 8 + n_pt);
    const long long total = (long long)n_cells * B * (4 + n_pt);
    int grid = grid_for(total, 256);
    if (grid > kSMs * 8) grid = kSMs * 8;
    box_head_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, ld_y, eps, cells, n_cells, B, HW, Wc, *geom, wheel,
                                                                d_box0, ldb0, d_box1, ldb1, d_box2, ldb2, d_zw_local,
                                                                ld_zwl, d_zw_img, d_dmean, d_dstd, ld_dist, n_pt,
                                                                d_pt_src, ld_pt, d_y);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_normal_head_fwd(const float* y, int ld_y, int W, const float* eps, const int* cells, int n_cells,
                                     int B, int HW, int squash, float squash_scale, float* out, float* dmean,
                                     float* dstd, int ld_dist, float* xdst0, int ldx0, float* xdst1, int ldx1,
                                     int n_pt, float* pt_dst, int ld_pt, void* stream) {
    SPAIR_REQUIRE(y && eps && cells && out && dmean && dstd && n_cells > 0 && B > 0 && W > 0);
    SPAIR_REQUIRE(ld_y >= 2 * W + n_pt && (n_pt == 0 || pt_dst));
    const long long total = (long long)n_cells * B * (W + n_pt);
    int grid = grid_for(total, 256);
    if (grid > kSMs * 8) grid = kSMs * 8;
    normal_head_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, ld_y, W, eps, cells, n_cells, B, HW, squash,
                                                                   squash_scale, out, dmean, dstd, ld_dist, xdst0,
                                                                   ldx0, xdst1, ldx1, n_pt, pt_dst, ld_pt);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_normal_head_bwd(const float* y, int ld_y, int W, const float* eps, const int* cells, int n_cells,
                                     int B, int HW, int squash, float squash_scale, const float* wheel,
                                     const float* d_out0, int ldo0, const float* d_out1, int ldo1,
                                     const float* d_out2, int ldo2, const float* d_out_img, const float* d_dmean,
                                     const float* d_dstd, int ld_dist, int n_pt, const float* d_pt_src, int ld_pt,
                                     float* d_y, void* stream) {
    SPAIR_REQUIRE(y && eps && cells && d_y && n_cells > 0 && B > 0 && W > 0 && ld_y >= 2 * W + n_pt);
    const long long total = (long long)n_cells * B * (W + n_pt);
    int grid = grid_for(total, 256);
    if (grid > kSMs * 8) grid = kSMs * 8;
    normal_head_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        y, ld_y, W, eps, cells, n_cells, B, HW, squash, squash_scale, wheel, d_out0, ldo0, d_out1, ldo1, d_out2, ldo2,
        d_out_img, d_dmean, d_dstd, ld_dist, n_pt, d_pt_src, ld_pt, d_y);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_pres_head_fwd(const float* y, int ld_y, const float* u, const int* cells, int n_cells, int B,
                                   int HW, float* pres, void* stream) {
    SPAIR_REQUIRE(y && u && cells && pres && n_cells > 0 && B > 0 && ld_y >= 1);
    pres_head_fwd_kernel<<<grid_for((long long)n_cells * B, 128), 128, 0, (cudaStream_t)stream>>>(y, ld_y, u, cells,
                                                                                                  n_cells, B, HW, pres);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_pres_head_bwd(const float* y, int ld_y, const float* u, const int* cells, int n_cells, int B,
                                   int HW, const float* wheel, const float* d_pres_local, int ld_local,
                                   const float* d_pres_img, float* d_y, void* stream) {
    SPAIR_REQUIRE(y && u && cells && wheel && d_y && n_cells > 0 && B > 0 && ld_y >= 1);
    pres_head_bwd_kernel<<<grid_for((long long)n_cells * B, 128), 128, 0, (cudaStream_t)stream>>>(
        y, ld_y, u, cells, n_cells, B, HW, wheel, d_pres_local, ld_local, d_pres_img, d_y);
    SPAIR_LAUNCH_CHECK();
}

extern "C" int spair_relu_bwd(float* dh, int ld_dh, const float* h, int ld_h, int rows, int cols, void* stream) {
    SPAIR_REQUIRE(dh && h && rows > 0 && cols > 0 && ld_dh >= cols && ld_h >= cols);
    int grid = grid_for((long long)rows * cols, 256);
    if (grid > kSMs * 16) grid = kSMs * 16;
    relu_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dh, ld_dh, h, ld_h, rows, cols);
    SPAIR_LAUNCH_CHECK();
}
