/*
 * b200nerf.h — C-ABI of libb200nerf.so, the B200 (sm_100a) volumetric-rendering core.
 *
 * This is the drop-in boundary of SURVEY.md §8(b).  The reference (nerfstudio v1.1.5) has no FFI of its
 * own: its hot path calls three un-vendored CUDA packages (tiny-cuda-nn, nerfacc, gsplat) or falls back to
 * PyTorch ops.  Every entry point below replaces one of those call sites; the comment above each group
 * cites the reference file:line (relative to nerfstudio/) whose arithmetic it reproduces.
 *
 * Conventions
 *   - every function is `extern "C" int f(..., void* stream)`; `stream` is a cudaStream_t (NULL = default).
 *   - return 0 on success; negative = argument error (B2N_E_*); positive = cudaError_t of the launch.
 *     b2n_last_error() returns a static string describing the last failure on the calling thread.
 *   - all pointers are DEVICE pointers unless the name ends in `_host`; tensors are dense row-major fp32
 *     unless stated; int64 for index tensors that the reference exposes as torch.long.
 *   - the library allocates nothing persistent; outputs and workspaces are caller-allocated.
 *   - there is NO CPU fallback anywhere behind this header.
 */
#ifndef B200NERF_H
#define B200NERF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2N_OK 0
#define B2N_E_ARG (-1)      /* bad argument (null pointer, size out of range) */
#define B2N_E_UNSUPPORTED (-2) /* configuration outside what the kernels are built for */

#define B2N_MAX_LEVELS 32
#define B2N_MAX_MLP_LAYERS 10

/* grid addressing mode */
#define B2N_GRID_TORCH 0 /* nerfstudio torch path: field_components/encodings.py:398-458 */
#define B2N_GRID_TCNN 1  /* tiny-cuda-nn HashGrid semantics; parity unpinned */

/* activations (field_components/mlp.py:33-58 names) */
#define B2N_ACT_NONE 0
#define B2N_ACT_RELU 1
#define B2N_ACT_SIGMOID 2
#define B2N_ACT_SOFTPLUS 3
#define B2N_ACT_TANH 4

/* spacing functions of SpacedSampler subclasses (model_components/ray_samplers.py:131-248) */
#define B2N_SPACING_UNIFORM 0
#define B2N_SPACING_PIECEWISE 1 /* UniformLinDispPiecewiseSampler */
#define B2N_SPACING_LINDISP 2
#define B2N_SPACING_SQRT 3
#define B2N_SPACING_LOG 4

/* background handling of RGBRenderer (model_components/renderers.py:71-119) */
#define B2N_BG_NONE 0        /* "random": nothing blended */
#define B2N_BG_LAST_SAMPLE 1
#define B2N_BG_CONSTANT 2    /* white / black / explicit colour: bg[3] */

/* Multiresolution grid description (host struct, passed by pointer, copied into kernel params). */
typedef struct B2nGrid {
  int32_t n_levels;
  int32_t n_features; /* per level: 1, 2, 4 or 8 */
  int32_t log2_hashmap_size;
  int32_t mode; /* B2N_GRID_* */
  float scale[B2N_MAX_LEVELS];     /* torch: encoder.scalings[l]; tcnn: exp2(l*log2(g))*base-1 */
  uint32_t resolution[B2N_MAX_LEVELS]; /* tcnn: ceil(scale)+1 (dense stride); unused in torch mode */
  uint32_t offset[B2N_MAX_LEVELS];     /* first table row of the level (torch: l*T) */
  uint32_t size[B2N_MAX_LEVELS];       /* rows in the level (torch: T) */
  uint32_t hashed[B2N_MAX_LEVELS];     /* 1 = XOR-prime hash, 0 = dense index */
} B2nGrid;

/* Tiny MLP description (host struct).  Layer i maps in_i -> out_i with nn.Linear layout W[out_i][in_i]
 * (+ bias[out_i]; b[i] == NULL means no bias, as tcnn networks have none).  w/b are DEVICE pointers — the
 * reference keeps one nn.Parameter per layer (`layers.i.weight/bias`), so the boundary takes them as they are.
 * skip[i] != 0 means the layer input is cat([network_input, previous_hidden]) (mlp.py:172-173). */
typedef struct B2nMlp {
  int32_t n_layers;
  int32_t in_dim;
  int32_t hidden_act; /* B2N_ACT_* applied after every layer but the last */
  int32_t out_act;    /* B2N_ACT_* applied after the last layer */
  int32_t out_dims[B2N_MAX_MLP_LAYERS];
  int32_t skip[B2N_MAX_MLP_LAYERS];
  const float* w[B2N_MAX_MLP_LAYERS];
  const float* b[B2N_MAX_MLP_LAYERS];
} B2nMlp;
/* gradient destinations (device pointers, same shapes as w/b); entries may be NULL to skip; ACCUMULATED into */
typedef struct B2nMlpGrad {
  float* dw[B2N_MAX_MLP_LAYERS];
  float* db[B2N_MAX_MLP_LAYERS];
} B2nMlpGrad;

const char* b2n_version(void);
const char* b2n_last_error(void);
/* device properties the host side sizes grids with: fills sm_count, cc_major, cc_minor, max_smem_optin */
int b2n_device_info(int32_t* out4_host);
/* launch-geometry knobs (e.g. "hash_levels_per_block_fwd"); returns 1 if the key exists */
int b2n_tune(const char* key, int value);

/* ---- K1/K2: multiresolution hash grid --------------------------------------------------------------
 * replaces HashEncoding.pytorch_fwd / tcnn.Encoding("HashGrid")   field_components/encodings.py:362-365,417-463
 * x [N,3] in [0,1];  table [rows,F] fp32;  y [N, L*F] (row-major, level-major features).
 * idx_out (optional, may be NULL): int64 [N,L,8] corner rows in the reference's hashed_0..7 order (for parity). */
int b2n_hashgrid_fwd(const B2nGrid* grid_host, const float* x, const float* table, int64_t n, float* y,
                     int64_t* idx_out, void* stream);
/* dtable [rows,F] is ACCUMULATED into (caller zeroes);  dx [N,3] optional (NULL = skip) is overwritten. */
int b2n_hashgrid_bwd(const B2nGrid* grid_host, const float* x, const float* table, const float* dy, int64_t n,
                     float* dtable, float* dx, void* stream);

/* position gradient only: dx [N,3] = d<dy, y>/dx (overwritten), no table scatter — for callers that get dtable from the
 * run-length scatter of b2n_hashgrid_bwd(dx = NULL) and need d x as well (camera optimiser: positions_bwd -> poses). */
int b2n_hashgrid_dx(const B2nGrid* grid_host, const float* x, const float* table, const float* dy, int64_t n, float* dx,
                    void* stream);

/* ---- K3: tiny fused MLP (fp32 SIMT) -----------------------------------------------------------------
 * replaces MLP.pytorch_fwd / tcnn.Network   field_components/mlp.py:110-114,160-184
 * x [N,in] row-major, y [N,out_last].  `hidden` (optional workspace, required for bwd) receives the
 * post-activation outputs of layers 0..n_layers-2, feature-major: layer i at hidden + hid_off_i,
 * laid out [out_i][N]; hid_off_i = N * sum_{j<i} out_j.  The whole network must fit in shared memory
 * (B2N_E_UNSUPPORTED otherwise; all nerfacto / instant-ngp networks do). */
int b2n_mlp_fwd(const B2nMlp* mlp_host, const float* x, int64_t n, float* y, float* hidden, void* stream);
/* dx [N,in] optional (NULL = skip), overwritten. */
int b2n_mlp_bwd(const B2nMlp* mlp_host, const B2nMlpGrad* grad_host, const float* x, const float* y,
                const float* hidden, const float* dy, int64_t n, float* dx, void* stream);

/* Tensor-core variant of the same two calls: tcgen05.mma kind::tf32 with the 3xTF32 split (fp32-level accuracy,
 * ~1e-6 relative), accumulators in TMEM.  Eligible networks: <= 4 layers, every width <= 64, hidden widths a
 * multiple of 4, ReLU (or no) hidden activation, none/sigmoid/relu output, no skip connections — i.e. all
 * nerfacto / instant-ngp networks; otherwise B2N_E_UNSUPPORTED (use the SIMT call).  x / dx rows may be padded
 * (row strides in floats).  `hidden` here is ROW-major per layer: layer i at hidden + N * sum_{j<i} out_j,
 * laid out [N][out_i] (fwd and bwd of this variant agree; not interchangeable with the SIMT layout). */
int b2n_mlp_tc_fwd(const B2nMlp* mlp_host, const float* x, int64_t x_stride, int64_t n, float* y, float* hidden,
                   void* stream);
int b2n_mlp_tc_bwd(const B2nMlp* mlp_host, const B2nMlpGrad* grad_host, const float* x, int64_t x_stride,
                   const float* y, const float* hidden, const float* dy, int64_t n, float* dx, int64_t dx_stride,
                   void* stream);

/* TMA staging of the weights: b2n_mlp_tc_pack writes (once per optimisation step) the exact shared-memory operand image
 * of both directions — hi/lo tf32 split, UMMA core-matrix layout, hi and lo halves stacked, biases — into a caller
 * workspace of b2n_mlp_tc_workspace_bytes() bytes (128-byte aligned); the *_ws calls then stage it into every persistent
 * CTA with cp.async.bulk.tensor (tensor map built per call, mbarrier complete_tx) instead of per-CTA scalar loads.
 * workspace == NULL: same as the calls above. */
int64_t b2n_mlp_tc_workspace_bytes(const B2nMlp* mlp_host);
int b2n_mlp_tc_pack(const B2nMlp* mlp_host, void* workspace, void* stream);
int b2n_mlp_tc_fwd_ws(const B2nMlp* mlp_host, const float* x, int64_t x_stride, int64_t n, float* y, float* hidden,
                      const void* workspace, void* stream);
int b2n_mlp_tc_bwd_ws(const B2nMlp* mlp_host, const B2nMlpGrad* grad_host, const float* x, int64_t x_stride,
                      const float* y, const float* hidden, const float* dy, int64_t n, float* dx, int64_t dx_stride,
                      const void* workspace, void* stream);

/* ---- wide dense layers (vanilla-nerf 8 x 256 with skip: fields/vanilla_nerf_field.py:84-107, mlp.py:160-179) --------------
 * One layer per call on a register-tiled fp32 GEMM with fused bias + activation (networks too wide for the fused kernels).
 *  fwd: y [n,out] = act(x [n,in] (row stride x_stride) w[out,in]^T + b).
 *  bwd: dz_scratch [n,out] receives dy * act'(y); dx [n,in] (row stride dx_stride, overwritten; NULL = skip);
 *       dw [out,in] and db [out] are ACCUMULATED into (NULL = skip). */
int b2n_linear_fwd(const float* x, int64_t n, int32_t in_dim, int64_t x_stride, const float* w, const float* b,
                   int32_t out_dim, int32_t act, float* y, void* stream);
int b2n_linear_bwd(const float* x, int64_t n, int32_t in_dim, int64_t x_stride, const float* w, int32_t out_dim, int32_t act,
                   const float* y, const float* dy, float* dz_scratch, float* dx, int64_t dx_stride, float* dw, float* db,
                   void* stream);

/* ---- K5/K6: direction / frequency encodings ---------------------------------------------------------
 * SHEncoding (encodings.py:752-799, utils/spherical_harmonics.py:24-81): levels in 1..5, out [N,levels^2].
 * remap01 != 0 applies d <- (d+1)/2 first (fields/base_field.py:136-142 fused in). */
int b2n_sh_fwd(const float* dirs, int64_t n, int32_t levels, int32_t remap01, float* out, void* stream);
/* NeRFEncoding (encodings.py:148-186): out [N, D*F*2 (+D if include_input)], freqs_host[F] = 2^linspace. */
int b2n_freq_fwd(const float* x, int64_t n, int32_t d, const float* freqs_host, int32_t n_freq,
                 int32_t include_input, float* out, void* stream);
int b2n_freq_bwd(const float* x, const float* dout, int64_t n, int32_t d, const float* freqs_host, int32_t n_freq,
                 int32_t include_input, float* dx, void* stream);

/* ---- a8-a10: sample positions -> unit cube -----------------------------------------------------------
 * Frustums.get_positions + SceneContraction(inf) + (x+2)/4 | aabb normalise + selector
 *   cameras/rays.py:50-59; field_components/spatial_distortions.py:66-69; fields/nerfacto_field.py:205-213
 * Ray form: origins/directions [R,3], starts/ends [R,S] with row stride `bin_stride` (so starts=bins,
 * ends=bins+1 of an [R,S+1] edge array works).  Point form: pass positions [N,3] with R=N,S=1 and
 * directions=NULL.  x_out [R*S,3] (already multiplied by the selector), sel_out uint8 [R*S]. */
int b2n_positions_fwd(const float* origins, const float* directions, const float* starts, const float* ends,
                      int64_t bin_stride, int64_t n_rays, int32_t n_samples, int32_t contraction,
                      const float* aabb_host6, float* x_out, uint8_t* sel_out, void* stream);

/* backward of the ray form: dx [R*S,3] (gradient w.r.t. x_out) -> d_origins [R,3], d_directions [R,3] (overwritten, or
 * added to when accumulate != 0 — 2 = with atomics, for callers that run another producer of the same gradients
 * concurrently; either may be NULL), through the selector, the normalisation and the L-inf contraction's Jacobian.  Carries the photometric
 * gradient to CameraOptimizer's pose corrections (cameras/camera_optimizers.py:148-153). */
int b2n_positions_bwd(const float* origins, const float* directions, const float* starts, const float* ends,
                      int64_t bin_stride, int64_t n_rays, int32_t n_samples, int32_t contraction,
                      const float* aabb_host6, const float* dx, int32_t accumulate, float* d_origins, float* d_directions,
                      void* stream);

/* density = avg_init * exp(h) * sel, bwd: dh = g * avg_init * exp(clamp(h,-15,15)) * sel
 *   field_components/activations.py:28-41; fields/nerfacto_field.py:226-232.  h has row stride h_stride. */
int b2n_density_act_fwd(const float* h, int64_t h_stride, const uint8_t* sel, int64_t n, float avg_init,
                        float* density, void* stream);
int b2n_density_act_bwd(const float* h, int64_t h_stride, const uint8_t* sel, const float* g, int64_t n,
                        float avg_init, float* dh, int64_t dh_stride, void* stream);

/* ---- a5/a6: samplers --------------------------------------------------------------------------------
 * SpacedSampler.generate_ray_samples (ray_samplers.py:78-128).  lin [S+1] = linspace(0,1,S+1) computed by the
 * host with torch (bit-identical bins).  jitter: NULL (eval) | [R] (jitter_stride=0... single) | [R,S+1].
 * Outputs spacing bins sbins [R,S+1] and euclidean bins ebins [R,S+1]. */
int b2n_spaced_sample(const float* nears, const float* fars, const float* lin, const float* jitter,
                      int32_t jitter_per_bin, int64_t n_rays, int32_t n_samples, int32_t spacing,
                      float* sbins, float* ebins, void* stream);
/* PDFSampler.generate_ray_samples (ray_samplers.py:276-372), include_original=False.
 * bins [R,S+1] spacing-domain edges, weights [R,S] (annealing w**anneal fused in: pass anneal=1 for none),
 * u_base [nb] = linspace(0,1-1/nb,nb) from the host, jitter NULL | [R] | [R,nb].  anneal_dev (optional) points
 * to the exponent in device memory and overrides `anneal` (lets a captured CUDA graph see a changing schedule).
 * Outputs new_sbins [R,nb], new_ebins [R,nb]; optional cdf_out [R,S+1], inds_out int64 [R,nb]. */
int b2n_pdf_sample(const float* bins, const float* weights, const float* u_base, const float* jitter,
                   int32_t jitter_per_bin, const float* nears, const float* fars, int64_t n_rays, int32_t n_in,
                   int32_t n_out, float anneal, const float* anneal_dev, float histogram_padding, float eps,
                   int32_t spacing,
                   float* new_sbins, float* new_ebins, float* cdf_out, int64_t* inds_out, void* stream);
/* RaySamples.get_weights of the level being resampled (cameras/rays.py:129-152; euclidean edges `ebins`, density [R,n_in]
 * -> weights [R,n_in], written) followed by b2n_pdf_sample on those weights (histogram over the spacing-domain edges
 * `sbins`), one launch: the inner loop of ProposalNetworkSampler.generate_ray_samples
 * (model_components/ray_samplers.py:568-604).  Same arithmetic as b2n_weights_fwd + b2n_pdf_sample. */
int b2n_weights_pdf_sample(const float* sbins, const float* ebins, const float* density, const float* u_base,
                           const float* jitter, int32_t jitter_per_bin, const float* nears, const float* fars,
                           int64_t n_rays, int32_t n_in, int32_t n_out, float anneal, const float* anneal_dev,
                           float histogram_padding, float eps, int32_t spacing, float* weights, float* new_sbins,
                           float* new_ebins, void* stream);

/* parity pin of the PDF sampler's normaliser: out[r] = sum(x[r, 0..n_cols)) accumulated in the order of torch's CPU
 * `sum` kernel (8-lane vectors x 4 interleaved rows x 4 cascade levels; model_components/ray_samplers.py:306 calls it),
 * so that cdf and searchsorted indices reproduce the reference's bits. */
int b2n_torch_row_sum(const float* x, int64_t n_rows, int32_t n_cols, float* out, void* stream);

/* ---- a21-a23: transmittance weights and compositing ---------------------------------------------------
 * Sample intervals are given as starts/ends [R,S] with row stride `bin_stride`: for an [R,S+1] edge array
 * pass starts=edges, ends=edges+1, bin_stride=S+1 (what the reference's samplers produce).
 * RaySamples.get_weights (cameras/rays.py:129-152): density [R,S] -> w [R,S]. */
int b2n_weights_fwd(const float* starts, const float* ends, int64_t bin_stride, const float* density, int64_t n_rays,
                    int32_t n_samples, float* weights, void* stream);
int b2n_weights_bwd(const float* starts, const float* ends, int64_t bin_stride, const float* density,
                    const float* dweights, int64_t n_rays, int32_t n_samples, float* ddensity, void* stream);
/* RGBRenderer + AccumulationRenderer + DepthRenderer(expected|median) (renderers.py:71-119,292-385).
 * rgb [R,S,3]; outputs rgb_out [R,3], acc [R], depth_exp [R] (sum(w t)/(sum(w)+1e-10); the global min/max clip of
 * renderers.py:381 is applied by the host), depth_med [R], med_idx int64 [R].
 * Any output pointer may be NULL.  eval_mode: nan_to_num(rgb) before, clamp01 after (renderers.py:225-231). */
int b2n_composite_fwd(const float* rgb, const float* weights, const float* starts, const float* ends,
                      int64_t bin_stride, int64_t n_rays, int32_t n_samples, int32_t bg_mode, const float* bg_host3,
                      int32_t eval_mode, float* rgb_out, float* acc, float* depth_exp, float* depth_med,
                      int64_t* med_idx, void* stream);
/* grads wrt rgb samples and weights given d_rgb_out [R,3], d_acc [R] (NULL=0), d_depth_exp [R] (NULL=0). */
int b2n_composite_bwd(const float* rgb, const float* weights, const float* starts, const float* ends,
                      int64_t bin_stride, const float* d_rgb_out, const float* d_acc, const float* d_depth_exp,
                      int64_t n_rays, int32_t n_samples, int32_t bg_mode, const float* bg_host3, float* d_rgb,
                      float* d_weights, void* stream);

/* ---- a24: proposal losses (model_components/losses.py:53-155) -------------------------------------------
 * interlevel term of ONE proposal level: c [R,Sc+1], w [R,Sc] (final level, constants), cp [R,Sp+1], wp [R,Sp].
 * loss_rows [R] = sum_i clip(w-w_outer,0)^2/(w+1e-7)  (host divides by R*Sc for the mean);
 * d_wp [R,Sp] = d(sum over the row)/d wp, scaled by `gscale` (NULL = skip). */
int b2n_interlevel_fwd_bwd(const float* c, const float* w, const float* cp, const float* wp, int64_t n_rays,
                           int32_t sc, int32_t sp, float gscale, float* loss_rows, float* d_wp, void* stream);
/* distortion: t [R,S+1], w [R,S] -> loss_rows [R]; d_w [R,S] scaled by gscale (NULL = skip). */
int b2n_distortion_fwd_bwd(const float* t, const float* w, int64_t n_rays, int32_t s, float gscale, float* loss_rows,
                           float* d_w, void* stream);

/* ---- a1/a2: ray generation (cameras/cameras.py:599-929; camera_utils.py:318-330,375-478) ---------------
 * perspective cameras with optional OpenCV distortion (k1,k2,k3,k4,p1,p2), pixel-centre coordinates.
 * c2w [C,3,4], intr [C,4]=(fx,fy,cx,cy), dist [C,6] or NULL, ray_indices int64 [R,3]=(cam,row,col). */
int b2n_raygen(const float* c2w, const float* intr, const float* dist, const int64_t* ray_indices, int64_t n_rays,
               float* origins, float* directions, float* pixel_area, float* directions_norm, int64_t* camera_indices,
               void* stream);
/* Cameras.generate_rays(camera_indices, coords, camera_opt_to_camera, distortion_params_delta, keep_shape)
 * (cameras/cameras.py:321-503) for perspective cameras.  coords != NULL: ray i = camera cam_idx[i], pixel-centre
 * coords[i] = (y, x) float.  coords == NULL: the whole images of the n_cams listed cameras, rays laid out
 * [height, width, n_cams] (camera fastest — the reference's (h, w, num_rays) shape), n_rays = n_cams*height*width.
 * cam_opt [n,3,4] / dist_delta [n,6] optional, n = n_rays (coords given) or n_cams (whole images). */
int b2n_raygen_coords(const float* c2w, const float* intr, const float* dist, const int64_t* cam_idx,
                      const float* coords, int64_t n_rays, int32_t n_cams, int32_t height, int32_t width,
                      const float* cam_opt, const float* dist_delta, float* origins, float* directions,
                      float* pixel_area, float* directions_norm, int64_t* camera_indices, void* stream);
/* Device-side training-ray pipeline: PixelSampler.sample_method + collate_image_dataset_batch
 * (data/pixel_samplers.py:137-174,265-318) + RayGenerator (model_components/ray_generators.py:41-56) over a uint8 image
 * cache resident in HBM: images uint8 [n_images, height, width, channels >= 3]; image_idx int64 [n_images] (absolute
 * camera index of each cached image; NULL = identity); u [R,3] = the sampler's U(0,1) draws, indices = trunc(u * [n_images,
 * height, width]).  Outputs (each optional): ray_indices int64 [R,3] (camera,row,col), the RayBundle fields as b2n_raygen,
 * rgb [R,3] = pixel / 255 (get_image_float32). */
int b2n_pixel_sample_raygen(const float* c2w, const float* intr, const float* dist, const uint8_t* images,
                            int32_t n_images, int32_t height, int32_t width, int32_t channels, const int64_t* image_idx,
                            const float* u, int64_t n_rays, int64_t* ray_indices, float* origins, float* directions,
                            float* pixel_area, float* directions_norm, int64_t* camera_indices, float* rgb, void* stream);
/* AABBBoxCollider (model_components/scene_colliders.py:47-108). */
int b2n_aabb_collide(const float* origins, const float* directions, const float* aabb_host6, float near_plane,
                     int64_t n_rays, float* nears, float* fars, void* stream);

/* CameraOptimizer.apply_to_raybundle, mode SO3xR3 (cameras/camera_optimizers.py:148-153; exponential map
 * cameras/lie_groups.py:25-58): out_origins = origins + t[cam], out_directions = R(w[cam]) directions with
 * pose_adjustment [C,6] = (t | w).  frozen: optional uint8 [C], 1 = non-trainable camera (identity, no gradient;
 * camera_optimizers.py:127-131).  bwd ACCUMULATES d_pose_adjustment [C,6]; either upstream gradient may be NULL. */
int b2n_pose_apply_fwd(const float* pose_adjustment, const int64_t* camera_indices, const uint8_t* frozen,
                       const float* origins, const float* directions, int64_t n_rays, float* out_origins,
                       float* out_directions, void* stream);
int b2n_pose_apply_bwd(const float* pose_adjustment, const int64_t* camera_indices, const uint8_t* frozen,
                       const float* directions, const float* d_out_origins, const float* d_out_directions,
                       int64_t n_rays, float* d_pose_adjustment, void* stream);

/* CameraOptimizer.get_loss_dict (cameras/camera_optimizers.py:155-162): loss_out[0] += mean_c |t_c| trans_pen + mean_c |w_c|
 * rot_pen; d_pose_adjustment [C,6] += gscale * its gradient (either output may be NULL). */
int b2n_pose_regularizer(const float* pose_adjustment, int32_t n_cams, float trans_l2_penalty, float rot_l2_penalty,
                         float gscale, float* loss_out, float* d_pose_adjustment, void* stream);

/* ---- K7-K12: packed (instant-ngp) path; nerfacc 0.5.2 call sites models/instant_ngp.py:120-198 ----------
 * pack_info: ray_indices int64 [M] sorted -> packed_info int64 [R,2] (start,count). */
int b2n_pack_info(const int64_t* ray_indices, int64_t m, int64_t n_rays, int64_t* packed_info, void* stream);
/* render_weight_from_density on packed samples: per-ray exclusive scan of sigma*dt. */
int b2n_packed_weights_fwd(const float* t_starts, const float* t_ends, const float* sigmas,
                           const int64_t* packed_info, int64_t n_rays, float* weights, float* trans, float* alphas,
                           void* stream);
int b2n_packed_weights_bwd(const float* t_starts, const float* t_ends, const float* sigmas,
                           const int64_t* packed_info, const float* dweights, int64_t n_rays, float* dsigmas,
                           void* stream);
/* accumulate_along_rays: out[r,:] = sum_{i in ray r} w[i]*values[i,:] (values NULL -> D=1, v=1). */
int b2n_packed_accumulate_fwd(const float* weights, const float* values, int32_t d, const int64_t* packed_info,
                              int64_t n_rays, float* out, void* stream);
int b2n_packed_accumulate_bwd(const float* weights, const float* values, int32_t d, const int64_t* packed_info,
                              const float* dout, int64_t n_rays, float* dweights, float* dvalues, void* stream);
/* occupancy-grid ray marching, two passes.  binaries uint8 [levels,res,res,res]; roi aabb (xmin..zmax) host.
 * pass 1 (counts): samples per ray -> counts int32 [R].  Host (or caller) turns counts into offsets [R]
 * (exclusive scan) and total M; pass 2 writes ray_indices int64 [M], t_starts/t_ends [M] at those offsets. */
int b2n_occgrid_count(const float* origins, const float* directions, const float* t_min, const float* t_max,
                      const uint8_t* binaries, int32_t levels, int32_t res, const float* roi_host6, float step,
                      float cone_angle, float near_plane, float far_plane, const float* jitter, int64_t n_rays,
                      int32_t* counts, void* stream);
int b2n_occgrid_fill(const float* origins, const float* directions, const float* t_min, const float* t_max,
                     const uint8_t* binaries, int32_t levels, int32_t res, const float* roi_host6, float step,
                     float cone_angle, float near_plane, float far_plane, const float* jitter, int64_t n_rays,
                     const int64_t* offsets, int64_t* ray_indices, float* t_starts, float* t_ends, void* stream);

/* exclusive scan of per-ray counts -> pack offsets int64 [n] and the total (device int64) — the step between the count and
 * the fill pass of the march / the pruning (replaces a framework cumsum). */
int b2n_scan_counts(const int32_t* counts, int64_t n, int64_t* offsets, int64_t* total, void* stream);
/* same for large n (e.g. 1 M Gaussians): multi-CTA, scratch_sums int32 [ceil(n/1024)], scratch_offsets int64 [ceil(n/1024)] */
int b2n_scan_counts_ws(const int32_t* counts, int64_t n, int64_t* offsets, int64_t* total, int32_t* scratch_sums,
                       int64_t* scratch_offsets, void* stream);
/* packed sample midpoints x[i] = o[ray_i] + d[ray_i] * (ts_i + te_i)/2: what VolumetricSampler.get_sigma_fn evaluates the
 * density at (model_components/ray_samplers.py:406-430). */
int b2n_packed_positions(const float* origins, const float* directions, const int64_t* ray_indices,
                         const float* t_starts, const float* t_ends, int64_t m, float* x, void* stream);
/* K8: visibility pruning of marched samples (nerfacc render_visibility_from_density inside OccGridEstimator.sampling,
 * called from ray_samplers.py:481-493): keep sample i iff trans[i] >= early_stop_eps && alphas[i] >= alpha_thre.
 * alpha_cap_dev (optional device scalar): the threshold used is min(alpha_thre, *alpha_cap_dev) — nerfacc caps alpha_thre
 * with occs.mean(), which b2n_occgrid_binarize leaves in device memory.
 * Two passes like the march: counts int32 [R] -> b2n_scan_counts -> fill (order within a ray preserved). */
int b2n_packed_prune_count(const float* trans, const float* alphas, const int64_t* packed_info, int64_t n_rays,
                           float early_stop_eps, float alpha_thre, const float* alpha_cap_dev, int32_t* counts,
                           void* stream);
int b2n_packed_prune_fill(const float* trans, const float* alphas, const int64_t* packed_info, int64_t n_rays,
                          float early_stop_eps, float alpha_thre, const float* alpha_cap_dev, const int64_t* offsets,
                          const float* t_starts,
                          const float* t_ends, int64_t* out_ray_indices, float* out_t_starts, float* out_t_ends,
                          void* stream);
/* K9: occupancy-grid update (nerfacc OccGridEstimator.update_every_n_steps as called by models/instant_ngp.py:149-164).
 *  points:   x [n,3] = lo + ((coord(cell) + jitter)/res) * (hi - lo) for one level; cell_ids int64 [n] or NULL (= 0..n-1,
 *            the warm-up "all cells" case); coord = (id / res^2, id / res % res, id % res); level_aabb_host6 = (lo | hi).
 *  ema:      occs[level_offset + cell] = max(occs[..] * ema_decay, occ_new) — candidates are formed from the OLD values;
 *            a cell listed several times gets the largest candidate.  scratch_n: float [n].
 *  binarize: binaries[i] = occs[i] > min(mean(occs), occ_thre) over all levels (n cells); the mean is a deterministic
 *            fp64 reduction.  scratch_parts: double [512]; stats_out: float [2] = (threshold used, mean(occs));
 *            binaries may be NULL (statistics only). */
int b2n_occgrid_points(const int64_t* cell_ids, const float* jitter, int64_t n, int32_t res,
                       const float* level_aabb_host6, float* x, void* stream);
int b2n_occgrid_ema(float* occs, const int64_t* cell_ids, const float* occ_new, int64_t n, int64_t level_offset,
                    float ema_decay, float* scratch_n, void* stream);
int b2n_occgrid_binarize(const float* occs, int64_t n, float occ_thre, double* scratch_parts, float* stats_out,
                         uint8_t* binaries, void* stream);
/* nerfacc.ray_aabb_intersect (models/instant_ngp.py does not call it per step; VolumetricSampler users and the reference's
 * tests/utils/test_aabb_intersection.py do): aabbs [K,6]; t_mins/t_maxs [n,K], hits uint8 [n,K]; misses = miss_value. */
int b2n_ray_aabb_intersect(const float* origins, const float* directions, const float* aabbs, int64_t n_rays,
                           int32_t n_boxes, float near_plane, float far_plane, float miss_value, float* t_mins,
                           float* t_maxs, uint8_t* hits, void* stream);

/* ---- 3D Gaussian splatting (gsplat.rendering.rasterization as models/splatfacto.py:555-581 calls it) -----
 * gsplat 1.4.0's sources are unavailable: the stages restate its published algorithm (oracle/splat_oracle.py) — PARITY
 * UNPINNED.  One camera per call: viewmat_host16 = world->camera 4x4 (row-major, host), k_host9 = intrinsics 3x3 (host).
 *  project_fwd  means [N,3], quats [N,4] (w,x,y,z; normalised inside), scales [N,3] -> means2d [N,2], depths [N], conics
 *               [N,3] (inverse of the eps2d-blurred 2-D covariance: a, b, c of [[a,b],[b,c]]), radii int32 [N] (3 sigma,
 *               pixels; 0 = culled), tiles_touched int32 [N] (16x16 tiles in the radius box); sh [N,K,3] != NULL also
 *               writes colors [N,3] = max(SH_degree(dir) + 0.5, 0).
 *  emit         per (Gaussian, tile) key = tile << 32 | depth bits at offsets = exclusive scan of tiles_touched
 *               (b2n_scan_counts); the caller sorts the keys (any stable 64-bit sort) and permutes gaussian_ids with them.
 *  tile_ranges  tile_lo / tile_hi int32 [tiles] (caller zero-fills): range of each tile in the sorted list.
 *  rasterize    front-to-back blending per pixel centre: alpha = min(0.999, opacity exp(-sigma)), skip alpha < 1/255,
 *               stop before T <= 1e-4.  out [H,W,3] (or [H,W,4] with `extra` [N], e.g. depth), out_alpha [H,W],
 *               last_idx int32 [H,W] (for the backward).  The backward ACCUMULATES v_means2d / v_conics / v_opacities /
 *               v_colors (/ v_extra): caller zero-fills.
 *  project_bwd  (v_means2d, v_depths | NULL, v_conics, v_colors | NULL) -> v_means, v_quats, v_scales, v_sh (overwritten). */
int b2n_gs_project_fwd(const float* means, const float* quats, const float* scales, const float* sh, int32_t sh_k,
                       int32_t sh_degree, int64_t n, const float* viewmat_host16, const float* k_host9, int32_t width,
                       int32_t height, float near_plane, float far_plane, float eps2d, float radius_clip, float* means2d,
                       float* depths, float* conics, int32_t* radii, int32_t* tiles_touched, float* colors, void* stream);
int b2n_gs_project_bwd(const float* means, const float* quats, const float* scales, const float* sh, int32_t sh_k,
                       int32_t sh_degree, int64_t n, const float* viewmat_host16, const float* k_host9, int32_t width,
                       int32_t height, float near_plane, float far_plane, float eps2d, float radius_clip,
                       const int32_t* radii, const float* conics, const float* colors, const float* v_means2d,
                       const float* v_depths, const float* v_conics, const float* v_colors, float* v_means, float* v_quats,
                       float* v_scales, float* v_sh, void* stream);
int b2n_gs_emit(const float* means2d, const int32_t* radii, const float* depths, const int64_t* offsets, int64_t n,
                int32_t width, int32_t height, int64_t* keys, int32_t* gaussian_ids, void* stream);
int b2n_gs_tile_ranges(const int64_t* sorted_keys, int64_t m, int32_t* tile_lo, int32_t* tile_hi, void* stream);
int b2n_gs_rasterize_fwd(int32_t width, int32_t height, const int32_t* tile_lo, const int32_t* tile_hi,
                         const int32_t* sorted_ids, const float* means2d, const float* conics, const float* opacities,
                         const float* colors, const float* extra, float* out, float* out_alpha, int32_t* last_idx,
                         void* stream);
int b2n_gs_rasterize_bwd(int32_t width, int32_t height, const int32_t* tile_lo, const int32_t* tile_hi,
                         const int32_t* sorted_ids, const float* means2d, const float* conics, const float* opacities,
                         const float* colors, const float* extra, const float* out_alpha, const int32_t* last_idx,
                         const float* v_out, const float* v_out_alpha, float* v_means2d, float* v_conics,
                         float* v_opacities, float* v_colors, float* v_extra, void* stream);

/* ---- optimiser step either side of the path: torch.optim.Adam semantics ---------------
 * p,g,m,v flat fp32 [n]; step is the 1-based step count; grads are multiplied by grad_scale first
 * (1/world_size after a sum-allreduce, or 1/loss_scale).  Hyper-parameters are doubles: 1-beta and the bias
 * corrections are formed in double, as torch does with its Python floats. */
int b2n_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int32_t step, double lr, double beta1,
                  double beta2, double eps, float grad_scale, void* stream);

/* same update with the per-step scalars in device memory: hyper3 = {lr/(1-beta1^t), 1/sqrt(1-beta2^t), grad_scale}
 * (CUDA-graph friendly: the host refreshes 12 bytes before each replay). */
int b2n_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper3, double beta1,
                      double beta2, double eps, void* stream);

/* ---- glue of the captured nerfacto step (fields/nerfacto_field.py:234-310; models/nerfacto.py:363-391) -----
 * head_in [R*S, out_stride >= n_sh+geo+n_emb] = [ sh[ray] | base_out[n, 1:1+geo] | emb row ]; emb_mode 0 = zeros,
 * 1 = emb[cam[ray]] (training), 2 = emb[0] (a pre-averaged row, eval).  When out is 16-byte aligned and
 * out_stride % 4 == 0 the rows are written in whole 4-column groups: the padding columns
 * [n_sh+geo+n_emb, round_up_4(n_sh+geo+n_emb)) (all < out_stride) are zero-filled. */
int b2n_head_input_fwd(const float* sh, int32_t n_sh, const float* base_out, int32_t base_w, int32_t geo,
                       const float* emb, const int64_t* cam, int32_t n_emb, int32_t emb_mode, int64_t n_rays,
                       int32_t n_samples, float* out, int32_t out_stride, void* stream);
/* d_base_out [R*S, base_w] (col 0 = d_dens_pre, cols 1..geo from d_in) is overwritten; d_emb rows are ACCUMULATED. */
int b2n_head_input_bwd(const float* d_in, int32_t in_stride, int32_t n_sh, int32_t geo, int32_t n_emb, const float* d_dens_pre,
                       const int64_t* cam, int64_t n_rays, int32_t n_samples, float* d_base_out, int32_t base_w,
                       float* d_emb, void* stream);
/* The same two operators restricted to a part of the columns, so that a caller can run the part that does not depend on
 * the field off its critical path: forward part 1 = the 4-column groups that do not read base_out (SH + embedding columns:
 * per-ray constants), 2 = the groups that do (geo features), 0 = all; needs 16-byte aligned rows.  Backward part 1 =
 * d_base_out only, 2 = the embedding-row accumulation only (d_base_out may be NULL), 0 = both. */
int b2n_head_input_fwd_part(const float* sh, int32_t n_sh, const float* base_out, int32_t base_w, int32_t geo,
                            const float* emb, const int64_t* cam, int32_t n_emb, int32_t emb_mode, int64_t n_rays,
                            int32_t n_samples, float* out, int32_t out_stride, int32_t part, void* stream);
int b2n_head_input_bwd_part(const float* d_in, int32_t in_stride, int32_t n_sh, int32_t geo, int32_t n_emb,
                            const float* d_dens_pre, const int64_t* cam, int64_t n_rays, int32_t n_samples,
                            float* d_base_out, int32_t base_w, float* d_emb, int32_t part, void* stream);
/* loss_out[0] += mean((pred-gt)^2); d_pred = gscale * 2 (pred-gt)/n  (either output may be NULL) */
int b2n_mse_fwd_bwd(const float* pred, const float* gt, int64_t n, float gscale, float* loss_out, float* d_pred,
                    void* stream);
/* out[0] += scale * sum(rows[0..n)) */
int b2n_sum_rows(const float* rows, int64_t n, float scale, float* out, void* stream);

/* ---- step glue: keeps the captured training step free of framework launches --------------------------------------
 * b2n_step_begin: the step's stratified draws and accumulator zeroing in one launch.  uniforms[n] in [0,1) from
 * Philox-4x32-10: element 4q+j = word j of philox(counter = (q, draw), key = seed) >> 8, times 2^-24;
 * rng_state3 (device, 3 x uint64) = {seed, draw, 0}: `draw` is advanced by the kernel, so every replay of a captured
 * graph gets a fresh stream (replaces torch.rand in model_components/ray_samplers.py:99-105, :335-340).
 * zero0[n_zero0], zero1[n_zero1] are set to 0 (either may be NULL with n = 0). */
int b2n_step_begin(float* uniforms, int64_t n_uniforms, uint64_t* rng_state3, float* zero0, int64_t n_zero0, float* zero1,
                   int64_t n_zero1, void* stream);
/* dst[i] += src[i] */
int b2n_add_inplace(float* dst, const float* src, int64_t n, void* stream);
/* out[0] = ((terms[0] + terms[1]) + ...) (+ extra[0] if not NULL): the Trainer's sum over the loss dict
 * (engine/trainer.py:511). */
int b2n_loss_total(const float* terms, int32_t n_terms, const float* extra, float* out, void* stream);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream` (a memset node when captured) */
int b2n_zero_async(void* ptr, int64_t bytes, void* stream);

/* ---- the per-ray middle of the nerfacto training step in one launch (SURVEY 8 a21-a24 composed) ---------------------
 * One warp per ray: density activation (nerfacto_field.py:226-232) + RaySamples.get_weights (cameras/rays.py:129-152),
 * RGB / accumulation / expected + median depth renderers (model_components/renderers.py:71-119,292-385), the MSE gradient
 * (models/nerfacto.py:372), the interlevel loss against both proposal levels and the distortion loss
 * (model_components/losses.py:53-155), then the backward of compositing, get_weights (all three levels) and the density
 * activation.  Same arithmetic as the separate operators above (bit-identical gradients).
 * Inputs: sbins/ebins_main [R, S+1] (spacing / euclidean edges), base_out [R*S, base_stride] (column 0 = density
 * pre-activation), selector [R*S], rgb [R*S,3], gt [R,3]; per proposal level l (arrays of 2 HOST pointers to device
 * buffers): sbins [R,Sl+1], ebins [R,Sl+1], weights [R,Sl], density [R,Sl].  d_prop_weights2 == NULL: proposals frozen
 * (no proposal gradients are produced).  gscales: mse 1, interlevel mult/(R*S), distortion mult/R.
 * Outputs: density, weights [R,S]; rgb_out [R,3]; accumulation, depth_expected, depth_median [R]; d_rgb [R*S,3];
 * d_weights, d_weights_distortion [R,S]; d_density_pre [R*S]; d_prop_weights2[l], d_prop_density2[l] [R,Sl];
 * loss_rows4 = {interlevel 0, interlevel 1, distortion, squared rgb error} per ray [R] each. */
int b2n_nerfacto_ray_tail(int64_t n_rays, int32_t s_main, int32_t s_prop0, int32_t s_prop1, const float* sbins_main,
                          const float* ebins_main, const float* base_out, int32_t base_stride, const uint8_t* selector,
                          float avg_init, const float* rgb, const float* gt, int32_t bg_mode, const float* bg_host3,
                          float mse_gscale, float interlevel_gscale, float distortion_gscale,
                          const float* const* prop_sbins2, const float* const* prop_ebins2,
                          const float* const* prop_weights2, const float* const* prop_density2,
                          float* const* d_prop_weights2, float* const* d_prop_density2, float* density, float* weights,
                          float* rgb_out, float* accumulation, float* depth_expected, float* depth_median, float* d_rgb,
                          float* d_weights, float* d_weights_distortion, float* d_density_pre, float* const* loss_rows4,
                          void* stream);
/* losses5[0] = rgb_scale * sum(rows4[3]), [1] = interlevel_scale * (sum(rows4[0]) + sum(rows4[1])),
 * [2] = distortion_scale * sum(rows4[2]), [3] = ((l0 + l1) + l2) + losses5[4] (engine/trainer.py:511); fixed reduction
 * order (deterministic).  losses5[4] (the camera-optimiser regulariser) is read, not written. */
int b2n_loss_finalize(const float* const* loss_rows4, int64_t n_rays, float interlevel_scale, float distortion_scale,
                      float rgb_scale, float* losses5, void* stream);

/* ---- fused proposal density field (fields/density_fields.py:94-117) ---------------------------
 * ray sample -> unit cube -> hash grid (F=2, <= 8 levels) -> MLP in->16->1 (ReLU) -> avg_init * trunc_exp * selector,
 * ONE launch; the network is read from the device pointers in mlp_host (w[0] [16][in], b[0], w[1] [1][16], b[1]).
 * Ray form: origins/directions [R,3], starts/ends [R,S] (row stride bin_stride); point form: directions NULL and
 * origins = positions [R,3].  density [R*S].  Other shapes: B2N_E_UNSUPPORTED (use the unfused calls). */
int b2n_density_field_fwd(const B2nGrid* grid_host, const B2nMlp* mlp_host, const float* table, const float* origins,
                          const float* directions, const float* starts, const float* ends, int64_t bin_stride,
                          int64_t n_rays, int32_t n_samples, int32_t contraction, const float* aabb_host6,
                          float avg_init, float* density, void* stream);
/* backward from d_density [R*S]: recomputes the forward, ACCUMULATES into dtable and grad_host->dw/db[0..1]. */
int b2n_density_field_bwd(const B2nGrid* grid_host, const B2nMlp* mlp_host, const B2nMlpGrad* grad_host,
                          const float* table, const float* origins, const float* directions, const float* starts,
                          const float* ends, int64_t bin_stride, int64_t n_rays, int32_t n_samples,
                          int32_t contraction, const float* aabb_host6, float avg_init, const float* d_density,
                          float* dtable, void* stream);

/* same, visiting only the samples with a non-zero d_density: live_ws (int32 [R*S + 1], scratch) receives their count and
 * indices from a compaction pass.  Every gradient of the proposal field is proportional to d_density, which the interlevel
 * loss leaves exactly zero wherever the proposal histogram already bounds the final weights. */
int b2n_density_field_bwd_ws(const B2nGrid* grid_host, const B2nMlp* mlp_host, const B2nMlpGrad* grad_host,
                             const float* table, const float* origins, const float* directions, const float* starts,
                             const float* ends, int64_t bin_stride, int64_t n_rays, int32_t n_samples,
                             int32_t contraction, const float* aabb_host6, float avg_init, const float* d_density,
                             float* dtable, int32_t* live_ws, void* stream);

/* same with the position gradient: every live sample adds d loss / d (sample position), through the unit-cube map's Jacobian,
 * to its ray's d_origins [R,3] and (times the sample's mid-point distance) d_directions [R,3] — ACCUMULATED with atomics
 * (either may be NULL; ray form only).  This is how the interlevel loss reaches CameraOptimizer's pose corrections. */
int b2n_density_field_bwd_rays(const B2nGrid* grid_host, const B2nMlp* mlp_host, const B2nMlpGrad* grad_host,
                               const float* table, const float* origins, const float* directions, const float* starts,
                               const float* ends, int64_t bin_stride, int64_t n_rays, int32_t n_samples,
                               int32_t contraction, const float* aabb_host6, float avg_init, const float* d_density,
                               float* dtable, int32_t* live_ws, float* d_origins, float* d_directions, void* stream);

/* ---- tcgen05 self-test (diagnostic; pins the tensor-core operand/TMEM semantics the MLP kernels rely on) -------
 * One 128-row tile.  mode 0: D = A[128][k] * B[n][k]^T (K-major x K-major);  mode 1: D = A[128][k] * W[k][n]
 * (K-major x MN-major);  mode 2: D = A[128][m]^T * B[128][n] (MN-major x MN-major, reduction over the 128 rows).
 * three_pass: 3xTF32 (hi*hi + lo*hi + hi*lo).  out receives the raw TMEM image [128 lanes][n columns]. */
int b2n_tc_selftest(int32_t mode, int32_t three_pass, const float* a, int32_t a_rows, int32_t a_cols, const float* b,
                    int32_t b_rows, int32_t b_cols, int32_t m, int32_t n, int32_t k, float* out128xn, void* stream);

/* diagnostic: average cycles {issue, issue->completion} of n_mma back-to-back 128 x n x 8 tf32 MMAs (one CTA) */
int b2n_tc_timing(int32_t n_mma, int32_t n, int32_t reps, long long* out2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200NERF_H */
