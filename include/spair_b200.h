/*
 * spair_b200.h — C-ABI of libspair_b200.so: the sm_100a kernels behind SPAIR's per-cell
 * object pipeline (sample -> glimpse -> render, forward and backward).
 *
 * The reference (yonkshi/SPAIR_pytorch) has no FFI of its own: the path sits behind the
 * Python API of spair/models.py and spair/modules.py.  Each entry point below replaces a
 * block of that Python, cited as file:line, and is what a binding for this path would
 * call (see INTEGRATION.md for the ctypes stubs).  Conventions:
 *
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer to fp32 unless its
 *     comment says "host"; the library never allocates, never synchronises, keeps no
 *     global state and is re-entrant; `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 on success, a positive cudaError_t from the launch, or
 *     SPAIR_ERR_INVALID (-1) for a rejected argument (nothing is launched then);
 *   - "image-major" buffers are indexed [b][cell][...] with cell = h*Wc + w (the layout
 *     `to_H_W_C(z).view(-1, k)` produces at models.py:468,474);
 *   - "rows" of a wavefront are indexed r = k*B + b for the k-th cell in `cells`
 *     (cells with equal w + (L+1)*h are mutually independent).
 */
#ifndef SPAIR_B200_H
#define SPAIR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SPAIR_ERR_INVALID (-1)
#define SPAIR_MAX_NEIGHBOURS 12
#define SPAIR_ABI_VERSION 5   /* 2: packed sweep weights, sweep backward, stem, broadcast_rows; 3: tcgen05 GEMM; 4: tcgen05 sweep; 5: pre-split GEMM weights */

int spair_abi_version(void);

/* HOST helper (no device work): out[j] = the j-th normalised base coordinate of an n-point axis, i.e.
 * torch's `linspace(-1, 1, n) * (n - 1) / n` as used by F.affine_grid(align_corners=False)
 * (reference modules.py:265), reproduced bit-exactly in fp32.  The kernels evaluate the same
 * expression on the device; this entry point exists so the coordinate parity can be checked
 * without a GPU. */
int spair_base_grid(int n, float* out /* host, n floats */);

/* Box decode constants (host struct).  models.py:339-374, config.py:34,38-41. */
typedef struct spair_box_geom {
    float yx_scale;     /* MAX_YX - MIN_YX */
    float yx_min;       /* MIN_YX */
    float hw_scale;     /* MAX_HW - MIN_HW */
    float hw_min;       /* MIN_HW */
    float anchor;       /* ANCHORBOX_SHAPE[0] (the reference uses [0] for both axes, models.py:366) */
    float img_h, img_w; /* INPUT_IMAGE_SHAPE[1], [2] */
    float cell_ratio_y; /* (float)(pixels_per_cell[0] / image_height), models.py:373 */
    float cell_ratio_x; /* (float)(pixels_per_cell[1] / image_width),  models.py:374 */
} spair_box_geom;

/* ------------------------------------------------------------------------------------
 * L0  lateral context.  Replaces SPAIR._get_sequential_context (models.py:292-320) plus the
 * torch.cat((cell_feat, context)) at models.py:76 for all cells of one wavefront.
 * Writes [feat(F) | ctx(n_nb*(A+6))] for every row into up to three destinations (the input
 * buffers of box_network, z_network and obj_network; models.py:76,88,100).  A neighbour
 * outside the grid contributes `edge` (virtual_edge_element, models.py:273-290,316).
 * ---------------------------------------------------------------------------------- */
int spair_context_gather_fwd(const float* feat,            /* [B,F,Hc,Wc] backbone output */
                             const float* box,             /* [B,HW,4] image-major */
                             const float* attr,            /* [B,HW,A] */
                             const float* depth,           /* [B,HW]   */
                             const float* pres,            /* [B,HW]   */
                             const float* edge,            /* [A+6]    */
                             const int* cells, int n_cells,
                             const int* nb_offsets, int n_nb, /* host: n_nb pairs (dh,dw) */
                             int B, int F, int Hc, int Wc, int A,
                             float* dst0, int ld0, float* dst1, int ld1, float* dst2, int ld2,
                             void* stream);

/* Backward of the context concat, as a gather at the PRODUCER cell (deterministic, no
 * atomics): for every row of this wavefront sums, over the later cells that consumed it,
 * the context slices of up to three MLP-input gradient buffers.
 *   out[r, 0:A+6] = sum_{nb} sum_{s<3} dX_s[row(consumer), col0 + nb*(A+6) + j]           */
int spair_context_grad_gather(const float* dx0, int ld0, const float* dx1, int ld1,
                              const float* dx2, int ld2, /* [HW*B, ld] wavefront-major, any may be NULL */
                              int col0,                  /* column where the context block starts (= F) */
                              const int* cells, int n_cells,
                              const int* wf_pos,         /* [HW] position of each cell in wavefront-major order */
                              const int* nb_offsets, int n_nb, /* host */
                              int B, int Hc, int Wc, int A,
                              float* out, int ld_out,    /* [n_cells*B, >=A+6] */
                              void* stream);

/* ------------------------------------------------------------------------------------
 * L1+L2  box head.  Replaces latent_to_mean_std (modules.py:167-176), _freeze_learning
 * (models.py:413-429), the four _sample_z calls and SPAIR._build_box (models.py:322-381).
 * y[:,0:4] = means (cy,cx,h,w), y[:,4:8] = log-stds; y[:,8:8+n_pt] = passthrough features,
 * copied to pt_dst (torch.cat at models.py:88).
 * ---------------------------------------------------------------------------------- */
int spair_box_head_fwd(const float* y, int ld_y,
                       const float* eps,                  /* [B,HW,4] normal noise (cy,cx,h,w) */
                       const int* cells, int n_cells, int B, int HW, int Wc,
                       const spair_box_geom* geom,        /* host */
                       float* box,                        /* [B,HW,4] (cell_x, cell_y, width, height) */
                       float* z_where,                    /* [B,HW,4] (xt, yt, xs, ys) */
                       float* dmean, float* dstd, int ld_dist, /* [B,HW,ld_dist]; cols 0..3 written */
                       float* xdst0, int ldx0, float* xdst1, int ldx1, /* row-local copies of box (may be NULL) */
                       int n_pt, float* pt_dst, int ld_pt,
                       void* stream);

/* d_y[:,0:8] = (1 - wheel) * chain rule; d_y[:,8:8+n_pt] = d_pt_src (passthrough grad). */
int spair_box_head_bwd(const float* y, int ld_y, const float* eps,
                       const int* cells, int n_cells, int B, int HW, int Wc,
                       const spair_box_geom* geom, const float* wheel, /* device scalar: training wheel f */
                       const float* d_box0, int ldb0, const float* d_box1, int ldb1,
                       const float* d_box2, int ldb2,      /* row-local grads wrt box (may be NULL) */
                       const float* d_zw_local, int ld_zwl, /* row-local grad wrt z_where (glimpse), may be NULL */
                       const float* d_zw_img,              /* [B,HW,4] image-major grad wrt z_where (render), may be NULL */
                       const float* d_dmean, const float* d_dstd, int ld_dist, /* image-major, may be NULL */
                       int n_pt, const float* d_pt_src, int ld_pt,
                       float* d_y,                        /* [rows, ld_y] */
                       void* stream);

/* ------------------------------------------------------------------------------------
 * L1/L3  Normal heads (z_what: width A, identity; z_depth: width 1, 4*sigmoid(clamp)).
 * Replaces latent_to_mean_std + (_freeze_learning) + _sample_z (+ clamped_sigmoid) at
 * models.py:83-85 and models.py:92-97.  y[:,0:W] means, y[:,W:2W] log-stds,
 * y[:,2W:2W+n_pt] passthrough (z_network only).
 * ---------------------------------------------------------------------------------- */
int spair_normal_head_fwd(const float* y, int ld_y, int W,
                          const float* eps,               /* [B,HW,W] */
                          const int* cells, int n_cells, int B, int HW,
                          int squash, float squash_scale,   /* squash: out = scale*sigmoid(clamp(z,-10,10)) */
                          float* out,                     /* [B,HW,W] image-major */
                          float* dmean, float* dstd, int ld_dist, /* pre-offset to this head's first column */
                          float* xdst0, int ldx0, float* xdst1, int ldx1,
                          int n_pt, float* pt_dst, int ld_pt,
                          void* stream);

int spair_normal_head_bwd(const float* y, int ld_y, int W, const float* eps,
                          const int* cells, int n_cells, int B, int HW,
                          int squash, float squash_scale,
                          const float* wheel,             /* device scalar or NULL (= not frozen, attr) */
                          const float* d_out0, int ldo0, const float* d_out1, int ldo1,
                          const float* d_out2, int ldo2,  /* row-local grads wrt out */
                          const float* d_out_img,         /* [B,HW,W] image-major, may be NULL */
                          const float* d_dmean, const float* d_dstd, int ld_dist,
                          int n_pt, const float* d_pt_src, int ld_pt,
                          float* d_y, void* stream);

/* ------------------------------------------------------------------------------------
 * L4  presence head.  Replaces SPAIR._build_obj_pres (models.py:393-411).
 * ---------------------------------------------------------------------------------- */
int spair_pres_head_fwd(const float* y, int ld_y, const float* u /* [B,HW] uniform(0,1) */,
                        const int* cells, int n_cells, int B, int HW,
                        float* pres /* [B,HW] */, void* stream);

int spair_pres_head_bwd(const float* y, int ld_y, const float* u,
                        const int* cells, int n_cells, int B, int HW, const float* wheel,
                        const float* d_pres_local, int ld_local, /* row-local (context), may be NULL */
                        const float* d_pres_img,                 /* [B,HW] image-major, may be NULL */
                        float* d_y, void* stream);

/* ------------------------------------------------------------------------------------
 * G  glimpse extractor = stn(image, z_where, [Gh,Gw], inverse=False) (modules.py:216-273)
 * as called from SPAIR._encode_attr (models.py:383-391): fused affine_grid + bilinear
 * grid_sample, border padding, align_corners=False.  cells == NULL: n_cells*B... is ignored
 * and row r samples image r with z_where row r (the plain stn() call); otherwise row
 * r = k*B + b samples image b with z_where[b, cells[k]].
 * ---------------------------------------------------------------------------------- */
int spair_glimpse_fwd(const float* image,                 /* [B,C,Ih,Iw] */
                      const float* z_where,               /* [B,HW,4] image-major (or [B,4]) */
                      const int* cells, int n_cells, int B, int HW,
                      int C, int Ih, int Iw, int Gh, int Gw,
                      float* out, int ld_out,             /* [rows, C*Gh*Gw] */
                      void* stream);

int spair_glimpse_bwd(const float* image, const float* z_where,
                      const int* cells, int n_cells, int B, int HW,
                      int C, int Ih, int Iw, int Gh, int Gw,
                      const float* d_out, int ld_out,
                      float* d_z_where_local,             /* [rows,4], overwritten */
                      float* d_image,                     /* [B,C,Ih,Iw] accumulated with atomics, or NULL */
                      void* stream);

/* stn(image, z_where, [Oh,Ow], inverse=True) (modules.py:255-269): generic paste of n
 * images [n,C,Gh,Gw] onto [n,C,Oh,Ow] canvases, zeros padding.  API-compat path; the model
 * itself uses the fused renderer below. */
int spair_paste_fwd(const float* image, const float* z_where, int n, int C, int Gh, int Gw,
                    int Oh, int Ow, float* out, void* stream);
int spair_paste_bwd(const float* image, const float* z_where, int n, int C, int Gh, int Gw,
                    int Oh, int Ow, const float* d_out,
                    float* d_image /* zero-filled by caller, accumulated */, float* d_z_where /* [n,4] overwritten */,
                    void* stream);

/* ------------------------------------------------------------------------------------
 * R (+E)  fused renderer.  Replaces everything in SPAIR._render after the decoder MLP
 * (models.py:481-540) including stn(..., inverse=True): logit scale/bias + analytical
 * sigmoid, alpha *= z_pres, importance = max(alpha*z_depth, 0.01), inverse warp of
 * colour / alpha / importance, importance-normalised alpha compositing over all HW objects
 * and the final clamp — without materialising [N, C+2, Ih, Iw].  Optionally fuses the BCE of
 * SPAIR._build_loss (models.py:547): per-tile partial sums of
 * -(t*max(log p,-100) + (1-t)*max(log(1-p),-100)).
 * ---------------------------------------------------------------------------------- */
int spair_render_num_tiles(int B, int Ih, int Iw);  /* length of bce_partial */

int spair_render_fwd(const float* logits,                 /* [B*HW, G, G, C+1] raw decoder output, or texel records */
                     const float* z_where,                /* [B*HW,4] */
                     const float* z_depth, const float* z_pres, /* [B*HW] */
                     int B, int HW, int C, int G, int Ih, int Iw,
                     float obj_scale, float alpha_scale, float alpha_bias,
                     int decoded,                         /* 0: `logits` are raw; 1: texel records written by spair_gemm3x with
                                                             SPAIR_GEMM_EPI_TEXEL (colour = sigma(obj_scale * l), alpha channel =
                                                             1 - sigma(alpha_scale * l + alpha_bias)); the scales are then unused
                                                             in the forward and only scale the gradient in the backward */
                     float* recon,                        /* [B,C,Ih,Iw] */
                     float* denom,                        /* [B,Ih,Iw]: sum_n(importance_n + 1e-9), negated where the clamp at 1 was active */
                     const float* target, float* bce_partial, /* both NULL or both set */
                     void* stream);

/* Gradient wrt recon is d_recon (may be NULL) + bce_scale * dBCE/drecon(recon, target)
 * (target may be NULL).  gs_ws is a [B,C+1,Ih,Iw] workspace.  d_logits is the gradient
 * wrt the RAW logits in both modes (with decoded = 1 the sigmoid derivative comes from the records). */
int spair_render_bwd(const float* logits, const float* z_where, const float* z_depth,
                     const float* z_pres, int B, int HW, int C, int G, int Ih, int Iw,
                     float obj_scale, float alpha_scale, float alpha_bias, int decoded,
                     const float* recon, const float* denom,
                     const float* d_recon, const float* target, const float* bce_scale /* device scalar or NULL (=1) */,
                     float* gs_ws,
                     float* d_logits, float* d_z_where, float* d_z_depth, float* d_z_pres,
                     void* stream);

/* ------------------------------------------------------------------------------------
 * K  KL terms.  Replaces SPAIR._compute_KL (models.py:169-262): z_pres-masked Normal KLs of
 * the D = 4+A+1 latent columns against the priors of config.py:45-52, and the sequential
 * count-prior scan for the presence KL.  count_dist0 is the normalised truncated geometric
 * prior of models.py:184-193 ([HW+1], computed by the caller exactly as the reference does).
 * kl_map[b,cell,0:D] Normal KLs, kl_map[b,cell,D] presence KL; kl_sums[b,0:7] per-name sums
 * in the order cy, cx, height, width, attr, depth, pres (the `torch.sum(z_kl, dim=[1,2,3])`
 * of models.py:553).
 * ---------------------------------------------------------------------------------- */
int spair_kl_fwd(const float* dmean, const float* dstd,   /* [B,HW,D] */
                 const float* pres,                       /* [B,HW] (z_pres == z_pres_prob, models.py:409) */
                 const float* prior_mean, const float* prior_std, /* [D] */
                 const float* count_dist0,                /* [HW+1] */
                 int B, int HW, int A,
                 float* kl_map,                           /* [B,HW,D+1] */
                 float* p_z,                              /* [B,HW] prior presence probability per cell */
                 float* kl_sums,                          /* [B,7] */
                 void* stream);

int spair_kl_bwd(const float* dmean, const float* dstd, const float* pres,
                 const float* prior_mean, const float* prior_std, const float* kl_map,
                 const float* p_z, const float* d_sums /* [B,7] */, int B, int HW, int A,
                 float* d_dmean, float* d_dstd, float* d_pres, void* stream);

/* ------------------------------------------------------------------------------------
 * L0-L4 + G fused: the WHOLE forward cell sweep (models.py:68-117) in one persistent launch.
 * The sweep is partitioned by image (all dependencies of a cell are cells of the same image): a CTA
 * owns `ipc` images and walks every wavefront with block-level barriers only.  It performs, per
 * wavefront, exactly what spair_context_gather_fwd -> box MLP -> spair_box_head_fwd ->
 * spair_glimpse_fwd -> encoder MLP -> spair_normal_head_fwd (attr) -> z MLP ->
 * spair_normal_head_fwd (depth) -> obj MLP -> spair_pres_head_fwd do, and fills the same
 * wavefront-major activation buffers (x, h0, h1, y of each MLP; row = position(cell)*B + b), so the
 * backward pass is unchanged.  The MLPs (reference modules.py:124-165: two ReLU hidden layers + a
 * linear output, multi-head outputs concatenated) are evaluated by the kernel itself in fp32 SIMT
 * FMAs from PACKED weights (spair_sweep_pack_weights below).  Limits: max_cells * ipc <=
 * spair_sweep_max_rows(), layer widths <= 256, G <= 64.
 * ---------------------------------------------------------------------------------- */
typedef struct spair_sweep_dims {
    int B, HW, Hc, Wc;      /* batch, cells */
    int F, A, P;            /* backbone features, attribute dims, passthrough features */
    int C, Ih, Iw, G;       /* image channels / size, glimpse side */
    int ipc;                /* images per CTA */
    int n_wavefronts, max_cells, n_nb;
} spair_sweep_dims;

/* Packs nn.Linear weights w[n][k] (reference modules.py:124-165) for the two sweeps, all layers in one launch:
 *   fwd [ceil(k/4)][n][4] : fwd[(g*n + col)*4 + j] = w[col][4g+j]   (reduction over k, one float4 per column and group)
 *   bwd [ceil(n/4)][k][4] : bwd[(g*k + col)*4 + j] = w[4g+j][col]   (reduction over n)
 * zero padded; either destination may be NULL; both must be 16-byte aligned; n_layers <= 16. */
typedef struct spair_sweep_pack {
    const float* w; int n, k;
    float* fwd; float* bwd;
} spair_sweep_pack;

int spair_sweep_pack_weights(const spair_sweep_pack* layers /* host array, device pointers inside */, int n_layers,
                             void* stream);

typedef struct spair_sweep_mlp {
    const float* wt[3];     /* `fwd`-packed weights of hidden0, hidden1, output */
    const float* b[3];      /* biases [n] */
    int k[3], n[3];
    float* x; int ld_x;     /* [HW*B, ld_x] input rows; the kernel fills them */
    float* h0; float* h1;   /* [HW*B, n[0]], [HW*B, n[1]] post-ReLU activations (written) */
    float* y;               /* [HW*B, n[2]] outputs (written) */
} spair_sweep_mlp;

int spair_sweep_max_rows(void);

int spair_sweep_fwd(const spair_sweep_dims* dims,        /* host */
                    const int* order,                    /* [HW] cells in wavefront-major order (device) */
                    const int* starts,                   /* [n_wavefronts+1] (device) */
                    const int* nb_offsets,               /* host: n_nb pairs (dh,dw) */
                    const float* image, const float* feat, const float* edge,
                    const float* eps_where, const float* eps_attr, const float* eps_depth, const float* u_pres,
                    const spair_box_geom* geom,          /* host */
                    const spair_sweep_mlp* box_mlp, const spair_sweep_mlp* enc_mlp,
                    const spair_sweep_mlp* z_mlp, const spair_sweep_mlp* obj_mlp,   /* host structs, device pointers inside */
                    float* box, float* z_where, float* attr, float* depth, float* pres,
                    float* dmean, float* dstd,           /* image-major outputs */
                    void* stream);

/* Backward of the fused sweep, same image partition, wavefronts in reverse order.  Reads the forward
 * activations (h0, h1, y of each MLP) and the `bwd`-packed weights, writes dy / dh1 / dh0 / dx of every
 * row (wavefront-major; they feed the weight-gradient GEMMs and the feature / edge-element reductions done by
 * the caller).  d_* are the image-major gradients arriving from the renderer, the decoder and the KL terms
 * (any may be NULL).  Replaces spair_context_grad_gather, spair_pres_head_bwd, spair_normal_head_bwd x2,
 * spair_glimpse_bwd, spair_box_head_bwd, spair_relu_bwd and the per-wavefront dX GEMMs. */
typedef struct spair_sweep_mlp_bwd {
    const float* w[3];      /* `bwd`-packed weights of hidden0, hidden1, output */
    int k[3], n[3];
    const float* h0; const float* h1; const float* y;   /* forward activations */
    float* dx; int ld_dx;   /* [HW*B, ld_dx] gradient wrt the input rows (written) */
    float* dh0; float* dh1; float* dy;                  /* written */
} spair_sweep_mlp_bwd;

int spair_sweep_bwd(const spair_sweep_dims* dims, const int* order, const int* starts,
                    const int* wf_pos,                   /* [HW] position of each cell in wavefront-major order (device) */
                    const int* nb_offsets,               /* host */
                    const float* image, const float* z_where,
                    const float* eps_where, const float* eps_attr, const float* eps_depth, const float* u_pres,
                    const float* wheel,                  /* device scalar */
                    const spair_box_geom* geom,
                    const spair_sweep_mlp_bwd* box_mlp, const spair_sweep_mlp_bwd* enc_mlp,
                    const spair_sweep_mlp_bwd* z_mlp, const spair_sweep_mlp_bwd* obj_mlp,
                    const float* d_zw, const float* d_attr, const float* d_depth, const float* d_pres,
                    const float* d_dmean, const float* d_dstd,
                    void* stream);

/* ------------------------------------------------------------------------------------
 * The same two sweeps with their dense layers (reference modules.py:124-165, called inside the loop of
 * models.py:68-117) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 in split precision
 * (x = hi + lo; W_hi*[x_hi | x_lo] as one N = 32 MMA + W_lo*x_hi as one N = 16 MMA per k-step, fp32
 * accumulation in TMEM), weights as the M-side operand so an MMA costs what the <= 16 activation rows of a
 * CTA cost.  spair_sweep_tc_pack splits and swizzles all 12 weight matrices once per step into the exact
 * shared-memory image of the operand tiles, in consumption order, one stream per direction
 * (spair_sweep_tc_stream_floats floats each); the kernels stream them with bulk async copies through an
 * mbarrier ring that runs ahead across layer boundaries.  Arguments, buffers, limits and results are those
 * of spair_sweep_fwd / spair_sweep_bwd (results agree to ~1e-6 relative: different summation order).  The `wt`
 * fields of spair_sweep_mlp hold the UNPACKED nn.Linear weights [n][k] here: a hidden unit whose pre-activation is
 * within the split-precision rounding error of zero is re-evaluated from the fp32 operands with float64
 * accumulation, so the ReLU branch (and with it every gradient through the unit) is the exact one.  The `w`
 * fields of spair_sweep_mlp_bwd are ignored.  n / k: 12 layers in the order box0, box1, box2,
 * enc0 .. obj2 (nn.Linear weight [n][k]).
 * ---------------------------------------------------------------------------------- */
int spair_sweep_tc_stream_floats(const int* n, const int* k, int n_layers /* 12 */, int backward);
int spair_sweep_tc_pack(const float* const* w /* host array of 12 device pointers */, const int* n, const int* k,
                        int n_layers, float* fwd_stream, float* bwd_stream /* either may be NULL */, void* stream);
int spair_sweep_fwd_tc(const spair_sweep_dims* dims, const int* order, const int* starts, const int* nb_offsets,
                       const float* image, const float* feat, const float* edge,
                       const float* eps_where, const float* eps_attr, const float* eps_depth, const float* u_pres,
                       const spair_box_geom* geom,
                       const spair_sweep_mlp* box_mlp, const spair_sweep_mlp* enc_mlp,
                       const spair_sweep_mlp* z_mlp, const spair_sweep_mlp* obj_mlp,
                       float* box, float* z_where, float* attr, float* depth, float* pres,
                       float* dmean, float* dstd,
                       const float* fwd_stream, void* stream);
int spair_sweep_bwd_tc(const spair_sweep_dims* dims, const int* order, const int* starts, const int* wf_pos,
                       const int* nb_offsets, const float* image, const float* z_where,
                       const float* eps_where, const float* eps_attr, const float* eps_depth, const float* u_pres,
                       const float* wheel, const spair_box_geom* geom,
                       const spair_sweep_mlp_bwd* box_mlp, const spair_sweep_mlp_bwd* enc_mlp,
                       const spair_sweep_mlp_bwd* z_mlp, const spair_sweep_mlp_bwd* obj_mlp,
                       const float* d_zw, const float* d_attr, const float* d_depth, const float* d_pres,
                       const float* d_dmean, const float* d_dstd,
                       const float* bwd_stream, void* stream);

/* ------------------------------------------------------------------------------------
 * Backbone stem (caller side of the path): nn.ZeroPad2d + the first Conv2d + bias + ReLU of
 * Backbone (reference modules.py:12-111; layer 0 of config.py DEFAULT_BACKBONE_TOPOLOGY) in one
 * launch each way.  HBM-bound (C*16 taps per output): the forward writes the [B,Cout,Ho,Wo] map once,
 * the backward reads dy and the saved post-ReLU y once and produces d_w [Cout,C,4,4] and d_bias [Cout]
 * (the image has no gradient).  Supported: C in {1,3}, Cout = 128, k = 4; padding is the ZeroPad2d's
 * (top, left) — right / bottom padding is implied by Ho, Wo.  ws: workspace of
 * spair_stem_bwd_ctas() * Cout * (C*16 + 4) floats, 16-byte aligned.  nhwc = 1: y and dy are channels-last
 * [B,Ho,Wo,Cout] (16-byte aligned) — the layout the GEMM tail of the backbone (spair_conv_gemm3x) consumes, so no
 * transpose pass separates the two; same values either way.
 * ---------------------------------------------------------------------------------- */
int spair_stem_bwd_ctas(void);
int spair_stem_conv_fwd(const float* x, const float* w, const float* bias, int B, int C, int Ih, int Iw, int Cout,
                        int k, int stride, int pad_t, int pad_l, int Ho, int Wo, int nhwc, float* y, void* stream);
int spair_stem_conv_bwd(const float* x, const float* y, const float* dy, int B, int C, int Ih, int Iw, int Cout,
                        int k, int stride, int pad_t, int pad_l, int Ho, int Wo, int nhwc, float* ws, float* d_w,
                        float* d_bias, void* stream);

/* out[r][:] = row[:] for r < rows (vectorised when cols % 4 == 0): pre-fills the decoder's [B*HW, G*G*(C+1)] logits
 * (reference models.py:165,477) with the bias so that the output GEMM runs with beta = 1. */
int spair_broadcast_rows(const float* row, int rows, int cols, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * D*  the path's own dense contractions on the tcgen05 tensor cores at fp32 accuracy (3xTF32 split precision, fp32
 * accumulation in TMEM): the decoder MLP `object_decoder` (reference models.py:165,474-481, modules.py:124-165) and the
 * weight gradients of the per-object MLPs.  C[M,N] = A . B (+ bias) with the reduction over K:
 *   a_kmajor = 1: A is stored [M][lda] (K contiguous);  0: stored [K][lda] (M contiguous)
 *   b_kmajor = 1: B is stored [N][ldb] (K contiguous);  0: stored [K][ldb] (N contiguous)
 * so y = x W^T is (1,1), dx = dy W is (1,0) and dW = dy^T x is (0,0) on the tensors as torch stores them.
 * lda, ldb must be multiples of 4 floats and A, B 16-byte aligned (TMA).  epilogue (splits == 1 only):
 *   SPAIR_GEMM_EPI_NONE   C = acc + bias
 *   SPAIR_GEMM_EPI_RELU   C = max(acc + bias, 0)                                  (build_MLP hidden layers)
 *   SPAIR_GEMM_EPI_TEXEL  C = sigma(s * (acc + bias) + b) per channel of a `period`-channel texel: channels
 *                         0..period-2 use (s_colour, 0), the last one (s_alpha, b_alpha), sigma(x) = 1/(exp(-x)+1)
 *                         — the per-texel transcendental part of SPAIR._render (models.py:485-493), so the renderer
 *                         reads decoded texel records and the logits never reach HBM.
 * splits > 1 (long reductions with a small output): partial products go to `workspace` ([splits][M][N] floats) and are
 * summed in a fixed order by a second launch; spair_gemm_splits() returns the split count the library would pick.
 * ---------------------------------------------------------------------------------- */
#define SPAIR_GEMM_EPI_NONE 0
#define SPAIR_GEMM_EPI_RELU 1
#define SPAIR_GEMM_EPI_TEXEL 2
int spair_gemm_block_n(int N, int b_kmajor);
int spair_gemm_splits(int M, int N, int K);
int spair_gemm3x(const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor, float* C, int ldc, int M,
                 int N, int K, const float* bias /* [N] or NULL */, int epilogue, int period, float s_colour, float s_alpha,
                 float b_alpha, float* workspace, int splits,
                 unsigned* kink_ws, int kink_cap /* SPAIR_GEMM_EPI_RELU with both operands K-major: outputs whose
                     pre-activation is within the product's rounding error of zero are listed in kink_ws ([0] = counter,
                     then up to kink_cap entries) and re-evaluated with float64 accumulation by a second launch, so that
                     the ReLU takes the exact branch (the reference's gradient below the layer depends on it); NULL / 0: off */,
                 const float* B_hi, const float* B_lo /* both NULL, or the TF32 hi / lo planes of B (a weight matrix; same shape
                     and pitch; spair_split_tf32).  The kernel then loads the planes and only splits the A tiles — half of the
                     splitter work, identical results; B itself is still read by the exact-ReLU pass */,
                 void* stream);

/* hi[i] = src[i] rounded to TF32 (nearest, ties away; stored as fp32), lo[i] = TF32(src[i] - hi[i]): the operand split the
 * GEMM kernels otherwise perform on every tile, done once per step for the weight matrices (nn.Linear / nn.Conv2d
 * weights of reference modules.py:44-66,124-165, models.py:165). */
int spair_split_tf32(const float* src, float* hi, float* lo, int n, void* stream);

/* Implicit-GEMM convolution of the backbone tail on a channels-last input x [B,H,W,C] (k x k window, stride s, no
 * padding; reference modules.py:44-66), the patch matrix read tile by tile with TMA im2col loads instead of being
 * materialised:  mode 1 (forward)  out[B*Ho*Wo, Cout] = act(patches(x) . other^T + bias), other = weight [Cout][(kh,kw,c)];
 *                mode 2 (weight gradient)  out[Cout][(kh,kw,c)] = other^T . patches(x), other = dy [B*Ho*Wo, Cout].
 * C must be a multiple of 32.  epilogue / workspace / splits / kink_ws as in spair_gemm3x. */
int spair_conv_gemm3x(const float* x, int B, int H, int W, int C, int k, int stride, int mode, const float* other,
                      int ld_other, float* out, int ldc, int Cout, const float* bias, int epilogue, float* workspace,
                      int splits, unsigned* kink_ws, int kink_cap,
                      const float* w_hi, const float* w_lo /* mode 1 only: NULL, or the TF32 planes of the weights `other` */,
                      void* stream);

/* Input gradient of the same convolution (k % stride == 0, Cout % 32 == 0) without a d_col matrix: stride*stride stride-1
 * sub-convolutions over dy (one GEMM per output parity class; the epilogue scatters the rows to their pixels of dx).
 * dy [B,Ho,Wo,Cout] channels-last; wc: the stride*stride packed class weights [Cin][T*T*Cout], T = k / stride,
 * wc[py*stride+px][c][(kh*T + kw)*Cout + co] = w[co][c][py + stride*(T-1-kh)][px + stride*(T-1-kw)];
 * dx [B,H,W,Cin] is fully overwritten. */
int spair_conv_dgrad3x(const float* dy, int B, int H, int W, int Cin, int k, int stride, int Cout, const float* wc,
                       float* dx, const float* wc_hi, const float* wc_lo /* NULL, or the TF32 planes of the class weights wc */,
                       void* stream);

/* Patch gather / transposed gather around spair_gemm3x for the k x k / stride s convolutions of the backbone tail
 * (reference modules.py:44-66), channels-last activations.  col[m][(kh*k + kw)*C + c] = x[b][s*oy+kh][s*ox+kw][c] with
 * m = (b*Ho + oy)*Wo + ox, Ho = (H-k)/s + 1 (no padding: the Backbone pads once, before the stem);
 * spair_col2im_nhwc is its adjoint (dx[b][y][x][c] = sum of the d_col entries that read x[b][y][x][c], fixed order).
 * C must be a multiple of 4; x / col / dx 16-byte aligned. */
int spair_im2col_nhwc(const float* x /* [B,H,W,C] */, int B, int H, int W, int C, int k, int stride,
                      float* col /* [B*Ho*Wo, k*k*C] */, void* stream);
int spair_col2im_nhwc(const float* dcol /* [B*Ho*Wo, k*k*C] */, int B, int H, int W, int C, int k, int stride,
                      float* dx /* [B,H,W,C] */, void* stream);

/* out[b][c][r] = in[b][r][c] (NCHW <-> channels-last of a [B, C, H*W] feature map), coalesced both ways. */
int spair_transpose_batched(const float* in, int B, int R, int C, float* out, void* stream);

/* Bias gradient of a dense / 1x1-conv layer fused with the ReLU mask of its output: if y != NULL, g[r][c] *= (y[r][c] > 0)
 * in place; out[c] = sum_r g[r][c] (two launches, fixed order).  cols, ld_g, ld_y multiples of 4, 16-byte aligned;
 * ws: spair_colsum_chunks(rows, cols) * cols floats. */
int spair_colsum_chunks(int rows, int cols);
int spair_relu_bwd_colsum(float* g, int ld_g, const float* y /* or NULL */, int ld_y, int rows, int cols, float* ws,
                          float* out /* [cols] */, void* stream);

/* Elementwise helper of the manual MLP backward: dh *= (h > 0), row-strided. */
int spair_relu_bwd(float* dh, int ld_dh, const float* h, int ld_h, int rows, int cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPAIR_B200_H */
