/*
 * nerfca.h -- C ABI of libnerfca_b200.so: the B200 (sm_100a) implementation of the NeRF-CA
 * training / rendering inner loop (kirstenmaas/NeRF-CA).
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; its boundary for this path is
 * the Python module surface (model/CPPN.py, model/Temporal.py, train/model_helpers.py,
 * train/proj_helpers.py).  Each entry point below names the reference function(s) it replaces.
 * The Python host side in nerf-ca_b200/ binds these with ctypes.
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer owned by the caller; the library never
 *     allocates, frees or retains caller memory and never synchronises the stream;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value 0 = ok, negative = error (NERFCA_E_*); nerfca_last_error() returns a
 *     thread-local message for the last failing call;
 *   - gradients are ACCUMULATED (+=) into the caller's buffers: zero-fill them first;
 *   - sample index p = ray * n_depth + s (ray-major), the order of
 *     `query_points.reshape((-1, 3))` in train/model_helpers.py:120.
 */
#ifndef NERFCA_H_
#define NERFCA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NERFCA_API __attribute__((visibility("default")))
#else
#define NERFCA_API
#endif

#define NERFCA_ABI_VERSION 2
#define NERFCA_MAX_LAYERS 8 /* first layer + hidden H->H layers + output layer */

enum { NERFCA_OK = 0, NERFCA_E_ARG = -1, NERFCA_E_UNSUPPORTED = -2, NERFCA_E_CUDA = -3, NERFCA_E_WORKSPACE = -4 };

enum { NERFCA_F32 = 0, NERFCA_F64 = 1 };

/* positional-encoding family, model/CPPN.py:112-135 */
enum {
  NERFCA_ENC_NONE = 0,   /* pos_enc == 'none' or pos_enc_basis == 0: identity                         */
  NERFCA_ENC_BANDS = 1,  /* [x, per band l: sin(2^l x) x3, sin(2^l x + pi/2) x3] * band_weight[l];      */
                         /* band_weight = freq_mask_alpha ('free_windowed'), the eased window          */
                         /* ('nerfies_windowed') or NULL (any other mode string: unwindowed)           */
  NERFCA_ENC_FOURIER = 2 /* [sin(v) | cos(v)], v = 2 pi * tile(x, L) * fourier_coeff (CPPN.py:115-118) */
};

/* output activation, train/model_helpers.py:63-70 */
enum { NERFCA_ACT_SIGMOID = 0, NERFCA_ACT_SOFTPLUS = 1, NERFCA_ACT_CLAMP = 2 };

/* arithmetic of the MLP chain */
enum {
  NERFCA_PREC_FP32 = 0, /* fp32 SIMT kernels (reference-exact arithmetic up to summation order)      */
  NERFCA_PREC_BF16 = 1  /* tcgen05 tensor cores: bf16 operands, fp32 accumulation in TMEM           */
};

/* One coordinate-MLP field: model/CPPN.py (n_latent == 0) or model/Temporal.py (n_latent > 0),
 * num_late_layers == 0 (the only functional branch of Temporal.query_time, Temporal.py:128-136).
 * weight[k] / bias[k] are the reference's nn.Linear tensors, row-major [out, in]:
 *   k = 0            early_pts_layers.0      [hidden, in_dim]
 *   k = 1..n_hidden  early_pts_layers.{2k}   [hidden, hidden]
 *   k = n_hidden+1   output_linear.0         [1, hidden]
 * bias[k] may be NULL (use_bias = False).  in_dim = enc_dim + n_latent with
 * enc_dim = 3 (NONE), 3 + 6 n_freq (BANDS), 6 n_freq (FOURIER).                                        */
typedef struct nerfca_field_t {
  int32_t enc_mode;
  int32_t n_freq;
  int32_t n_latent;            /* num_time_dim; 0 for the static field                                  */
  int32_t n_phases;            /* rows of time_latents (10 in the reference, Temporal.py:25-26)         */
  int32_t hidden;              /* num_filters                                                           */
  int32_t n_hidden;            /* num_early_layers                                                      */
  const float* band_weight;    /* device [n_freq] or NULL                                               */
  const float* fourier_coeff;  /* device [3 n_freq] (already multiplied by fourier_sigma) or NULL       */
  const float* latents;        /* device [n_phases, n_latent] or NULL                                   */
  const float* weight[NERFCA_MAX_LAYERS]; /* device                                                     */
  const float* bias[NERFCA_MAX_LAYERS];   /* device or NULL                                             */
} nerfca_field_t;

typedef struct nerfca_field_grads_t {
  float* latents;                     /* device [n_phases, n_latent] or NULL                            */
  float* weight[NERFCA_MAX_LAYERS];   /* device, same shapes as nerfca_field_t.weight                   */
  float* bias[NERFCA_MAX_LAYERS];     /* device or NULL                                                 */
} nerfca_field_grads_t;

/* Where the sample positions come from.  Either explicit points (module-level calls
 * CPPN.forward(x) / Temporal.forward_composite(x, ts)) or rays + one shared depth vector
 * (train/model_helpers.py:101,118: pts = o + d * z), in which case points are formed inside the
 * kernels and never stored.  ray_dtype selects the reference's rounding: F64 -> fl32(fl64(o + d*z))
 * (training path, rays come from the float64 ray table), F32 -> o + fl32(d*z) (eval path,
 * run_composite.py:351).                                                                                */
typedef struct nerfca_samples_t {
  int64_t n_points;             /* P (= n_rays * n_depth when ray-generated)                            */
  const float* points;          /* device [P,3] or NULL                                                 */
  int32_t n_rays;
  int32_t n_depth;
  const void* origins;          /* device, element (r, c) at origins[r * ray_stride + c]                */
  const void* dirs;             /* device, same addressing                                              */
  int32_t ray_dtype;            /* NERFCA_F32 / NERFCA_F64                                              */
  int32_t ray_stride;           /* elements between rays: 3 if packed, 12 for rows of rays_train[B,4,3] */
  const float* depth;           /* device [n_depth]                                                     */
  const int32_t* phase_point;   /* device [P] cardiac phase per sample, or NULL                         */
  const int32_t* phase_ray;     /* device [n_rays] cardiac phase per ray, or NULL                       */
} nerfca_samples_t;

/* Loss hyper-parameters of one training step: train/run_composite.py:276-292 and
 * train/model_helpers.py:250-262.  *_weight are the already-scheduled weights (linear_param_decay). */
typedef struct nerfca_loss_cfg_t {
  double favor_s_weight;        /* blend-ratio entropy                                                  */
  double dyn_entropy_weight;    /* dynamic ray entropy                                                  */
  double occl_weight;           /* dynamic occlusion                                                    */
  double l1_weight;             /* multiplies both static L1 and L2 (run_composite.py:292)              */
  double entro_mask_thre;
  double entro_weighted_thresh;
  int32_t entro_use_weighting;
  int32_t n_rays_global;        /* B of the whole job: the 1/B of every mean (== n_rays on one GPU)     */
} nerfca_loss_cfg_t;

#define NERFCA_N_LOSS_TERMS 16
/* indices into the float64 loss-term vector written by nerfca_composite_loss (raw SUMS over this
 * call's rays; the host divides by B / B*N and all-reduces across ranks)                              */
enum {
  NERFCA_T_PIXEL_SUM = 0,     /* sum_r w_r (pix_r - gt_r)^2                                             */
  NERFCA_T_BLENDW_SUM = 1,    /* sum_{r,s} sigma_d / (sigma_s + sigma_d + 1e-10)                        */
  NERFCA_T_SIGMA_S_MAX = 2,
  NERFCA_T_SIGMA_D_MAX = 3,
  NERFCA_T_FAVOR_SUM = 4,     /* sum_{r,s} binary entropy of the blend ratio                            */
  NERFCA_T_S_ENT_SUM = 5,     /* sum_r masked static ray entropy                                        */
  NERFCA_T_S_SUM_SUM = 6,     /* sum_r sum_s sigma_s delta                                              */
  NERFCA_T_D_ENT_SUM = 7,
  NERFCA_T_D_SUM_SUM = 8,
  NERFCA_T_OCCL_SUM = 9,      /* sum_r sum_s sigma_d delta                                              */
  NERFCA_T_L1_SUM = 10,
  NERFCA_T_L2_SUM = 11
};

NERFCA_API const char* nerfca_last_error(void);
NERFCA_API int nerfca_abi_version(void);

/* A2  train/proj_helpers.py:65-90 get_ray_values_tigre.  pose = float32 row-major 4x4 (host memory),
 * i.e. source_matrix_tigre(...) rounded to fp32 (:68).  Writes origins/dirs [W,H,3] float32 (device),
 * bit-identical to the reference's fp32 arithmetic.                                                    */
NERFCA_API int nerfca_gen_rays(const float* pose_host, int32_t width, int32_t height, float du, float dv, float off_u,
                    float off_v, float dsd, float* origins, float* dirs, void* stream);

/* A3  train/model_helpers.py:3-12 randomize_depth with the uniform draw supplied by the caller
 * (t_rand device [n]); out device [n].  Bit-identical fp32.                                           */
NERFCA_API int nerfca_jitter_depth(const float* z, const float* t_rand, int32_t n, float* out, void* stream);

/* N1  train/run_composite.py:250-273: assemble a training batch ON THE DEVICE from the device-resident ray table
 * rays_train [R,4,3] float64 / phases_train [R] int64 (train/data_helpers.py:157-163) and the step's ray ids
 * (int64 [B], drawn by the host RNG exactly as upstream): rays_out [B,4,3] float64 = rays_train[ids] (bit copy),
 * phases_out [B] int32 = phases_train[ids] (run_composite.py:265 `.int()`).  An id outside [0, R) is an argument
 * error reported through err_flag (device int32, set to 1; may be null).  Replaces the host-side fancy index of a
 * multi-GB float64 array plus a 96 B/ray H2D copy by an 8 B/ray copy of the ids.                          */
NERFCA_API int nerfca_gather_batch(const double* rays_table, const int64_t* phases_table, int64_t n_table,
                        const int64_t* ids, int32_t n_batch, double* rays_out, int32_t* phases_out,
                        int32_t* err_flag, void* stream);

/* A4  train/model_helpers.py:118-121 / run_composite.py:351-352: materialise the sample points
 * [P,3] float32 (device) of a ray-generated sample set.  Bit-identical.                               */
NERFCA_API int nerfca_sample_points(const nerfca_samples_t* samples, float* points_out, void* stream);

/* A5  CPPN.pos_enc / Temporal.pos_enc (+ the latent concat of Temporal.query_time:124):
 * writes the first-layer input [P, in_dim] float32 (device).  Parity / debug entry.                  */
NERFCA_API int nerfca_encode(const nerfca_field_t* field, const nerfca_samples_t* samples, float* enc_out, void* stream);

/* Scratch sizes (bytes) for the calls below; stash = activations kept for the backward pass.        */
NERFCA_API size_t nerfca_field_stash_bytes(const nerfca_field_t* field, int64_t n_points, int32_t precision);
NERFCA_API size_t nerfca_field_workspace_bytes(const nerfca_field_t* field, int64_t n_points, int32_t precision, int32_t backward);

/* A5-A7  CPPN.forward (model/CPPN.py:88-110) / Temporal.forward_composite (model/Temporal.py:138-151):
 * raw_out[P] = field(sample p).  stash may be NULL (inference); workspace must hold
 * nerfca_field_workspace_bytes(field, P, precision, 0).                                              */
NERFCA_API int nerfca_field_forward(const nerfca_field_t* field, const nerfca_samples_t* samples, int32_t precision,
                         float* raw_out, void* stash, void* workspace, void* stream);

/* A11  autograd of A5-A7: given d_raw[P] accumulates every parameter gradient of the field
 * (12 Linear tensors + time_latents).  No gradient flows to the sample positions (the reference
 * never differentiates w.r.t. rays).                                                                  */
NERFCA_API int nerfca_field_backward(const nerfca_field_t* field, const nerfca_samples_t* samples, int32_t precision,
                          const float* d_raw, const void* stash, void* workspace,
                          const nerfca_field_grads_t* grads, void* stream);

/* A9  train/model_helpers.py:72-97 render_volume_density[_composite].
 * raw_d == NULL selects the single-field form (sigma returned UNscaled, :91-92).  acc_dtype is the
 * dtype the reference would compute delta / the ray sum / pix in (ray_directions.dtype, :73): F64 in
 * training, F32 in eval; pix_out and dists_out have that dtype.  sigma_*_out [n_rays, n_depth] f32.   */
NERFCA_API int nerfca_integrate(const float* raw_s, const float* raw_d, const float* depth, const float* i0, int32_t n_rays,
                     int32_t n_depth, int32_t activation, int32_t acc_dtype, void* pix_out, float* sigma_s_out,
                     float* sigma_d_out, void* dists_out, void* stream);

/* autograd of nerfca_integrate: d_pix [n_rays] (acc_dtype), d_sigma_* [n_rays,n_depth] f32 (NULL = 0)
 * -> d_raw_s, d_raw_d [P] f32 (overwritten).                                                          */
NERFCA_API int nerfca_integrate_backward(const float* raw_s, const float* raw_d, const float* depth, int32_t n_rays,
                              int32_t n_depth, int32_t activation, int32_t acc_dtype, const void* d_pix,
                              const float* d_sigma_s, const float* d_sigma_d, float* d_raw_s, float* d_raw_d,
                              void* stream);

/* A9 + A10 fused, training form (float64 ray sums as in the reference): line integral, weighted MSE,
 * the regularisers of compute_losses (train/model_helpers.py:189-262) and the closed-form
 * dL/d_raw of the total loss of run_composite.py:292 (SURVEY 8(a')).  gt / wpix are float64 [n_rays]
 * (columns of the ray table; element r at gt[r * gw_stride]).  raw_d == NULL: the static run of
 * run_nerf.py:227-230 (loss = wMSE + occl_weight * occlusion, sigma unscaled).
 * terms_out: float64 [NERFCA_N_LOSS_TERMS] device, accumulated (+=, max for the two maxima).        */
NERFCA_API int nerfca_composite_loss(const float* raw_s, const float* raw_d, const float* depth, const float* i0,
                          const double* gt, const double* wpix, int32_t gw_stride, int32_t n_rays, int32_t n_depth,
                          int32_t activation, const nerfca_loss_cfg_t* cfg, double* pix_out, double* terms_out,
                          float* d_raw_s, float* d_raw_d, void* stream);

/* One whole training step of train/run_composite.py:283-305 (minus the optimizer) as ONE call: both fields forward over the
 * same ray-generated sample set (a single launch on the tcgen05 path), nerfca_composite_loss, both fields backward.
 * dynamic_field == NULL selects the static run of train/run_nerf.py:205-233.  Scratch is caller-allocated:
 * raw_* / d_raw_* float32 [P]; stash / workspace of nerfca_step_stash_bytes / nerfca_step_workspace_bytes bytes.
 * Parameter gradients are accumulated (+=); pix_out float64 [n_rays]; terms_out float64 [NERFCA_N_LOSS_TERMS] (+=).   */
enum {
  NERFCA_STEP_PACKED = 1,       /* the packed bf16 parameter blocks at the head of `workspace` are current (kept so by
                                   nerfca_adam_step's repack): skip the pack launch                                      */
  NERFCA_STEP_ZERO_TERMS = 2    /* clear terms_out before accumulating (done inside the forward launch)                  */
};
typedef struct nerfca_step_t {
  const nerfca_field_t* static_field;
  const nerfca_field_t* dynamic_field;
  const nerfca_field_grads_t* static_grads;
  const nerfca_field_grads_t* dynamic_grads;
  const nerfca_samples_t* samples;
  int32_t precision;
  int32_t activation;
  const float* i0;              /* device [n_rays]                                                       */
  const double* gt;             /* device, element r at gt[r * gw_stride]                                 */
  const double* wpix;
  int32_t gw_stride;
  int32_t flags;                /* NERFCA_STEP_* bits                                                      */
  const nerfca_loss_cfg_t* loss;
  float* raw_s; float* raw_d; float* d_raw_s; float* d_raw_d;
  void* stash; void* workspace;
  double* pix_out; double* terms_out;
} nerfca_step_t;
NERFCA_API size_t nerfca_step_stash_bytes(const nerfca_step_t* step);
NERFCA_API size_t nerfca_step_workspace_bytes(const nerfca_step_t* step);
NERFCA_API int nerfca_train_step(const nerfca_step_t* step, void* stream);

/* No-grad evaluation of both fields over one sample set in a single launch (the render path of
 * train/run_composite.py:346-361): raw_s / raw_d float32 [P]; dynamic_field / raw_d may be NULL.
 * workspace: nerfca_step_workspace_bytes of a step descriptor with the same fields / samples / precision.          */
NERFCA_API int nerfca_fields_forward(const nerfca_field_t* static_field, const nerfca_field_t* dynamic_field,
                          const nerfca_samples_t* samples, int32_t precision, float* raw_s, float* raw_d, void* workspace,
                          void* stream);

/* Render (train/run_composite.py:346-361,407-413; north_star (c)): no-grad rays -> pixels with the X-ray line integral FUSED into the
 * output layer's epilogue of the tcgen05 forward: every sample's sigma * delta is reduced per ray by a segmented warp-shuffle scan
 * and one atomic per warp and ray, so neither the per-sample field outputs nor the sigma arrays ever reach HBM.  Eval arithmetic
 * (float32 rays: o + fl32(d z), float32 sums).  pix = i0 - sum_s (sigma_s + sigma_d) delta;  pix_static / pix_dynamic (both or
 * neither; need dynamic_field): the single-field images of render_volume_density (:407-413).  precision must be NERFCA_PREC_BF16
 * (the fp32 path is nerfca_fields_forward + nerfca_integrate).  workspace: nerfca_render_workspace_bytes.                     */
NERFCA_API size_t nerfca_render_workspace_bytes(const nerfca_field_t* static_field, const nerfca_field_t* dynamic_field,
                                     const nerfca_samples_t* samples, int32_t precision);
NERFCA_API int nerfca_render_rays(const nerfca_field_t* static_field, const nerfca_field_t* dynamic_field, const nerfca_samples_t* samples,
                       int32_t precision, const float* i0, int32_t activation, float* pix, float* pix_static, float* pix_dynamic,
                       void* workspace, void* stream);

/* N4  display normalisation of an eval image, train/run_composite.py:394-413: out[i] = (img[i] - min) / (max - min) on the device
 * (float32 [n]; out may alias img).  minmax_out: optional device float[2] receiving (min, max); scratch8: 8 bytes of device scratch. */
NERFCA_API int nerfca_normalize_image(const float* img, int64_t n, float* out, float* minmax_out, void* scratch8, void* stream);

/* N2  torch.optim.Adam(foreach) + LinearLR of train/run_composite.py:209-215,305-308 over ONE flat fp32 parameter buffer
 * (params / grads / exp_avg / exp_avg_sq device [n]), bit-for-bit for fp32 parameters (tests/test_gpu_parity.py::
 * test_adam_matches_torch_bit_for_bit).  The per-update scalars are computed by the HOST in python double arithmetic exactly as
 * torch does and passed by value (they change every step; the graph replay of nerfca_graph_* updates them in place):
 *   lr                      the LinearLR value of this update (torch's recursive form, lr_scheduler.py::LinearLR.get_lr)
 *   bias_correction1        1 - beta1 ** t            bias_correction2_sqrt   (1 - beta2 ** t) ** 0.5       (t = 1, 2, ...)
 *   m = lerp(m, g, 1 - beta1);  v = beta2 v + (1 - beta2) (g g);  p += -(lr / bc1) * (m / (sqrt(v) / bc2_sqrt + eps))
 * grads are multiplied by grad_scale first (1 on one GPU) and, if zero_grads != 0, cleared afterwards so the next step's
 * kernels can accumulate without a separate memset.  repack != NULL: every updated parameter of the two fields is also written,
 * converted to bf16, to its position in the packed operand blocks at the head of `workspace` (the tcgen05 kernels' copy of the
 * weights), so the next nerfca_train_step can run with NERFCA_STEP_PACKED and no pack launch.                            */
typedef struct nerfca_adam_step_t {
  double lr, beta1, beta2, eps;
  double bias_correction1;
  double bias_correction2_sqrt;
} nerfca_adam_step_t;
typedef struct nerfca_repack_t {
  const nerfca_field_t* static_field;   /* weight / bias pointers must lie inside `params`                             */
  const nerfca_field_t* dynamic_field;  /* or NULL                                                                     */
  void* workspace;                      /* the step workspace (nerfca_step_workspace_bytes)                            */
} nerfca_repack_t;
NERFCA_API int nerfca_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                     const nerfca_adam_step_t* cfg, float grad_scale, int32_t zero_grads, const nerfca_repack_t* repack, void* stream);

/* Launch accounting and per-kernel device timing (used by bench.py for `gpu_launches` and the roofline line).
 * nerfca_launch_count: kernels launched by this library since load.  nerfca_profile_enable(1) starts recording a
 * CUDA-event pair around every kernel family launch on its own stream (reset first); nerfca_profile_enable(0) stops.
 * nerfca_profile_read synchronises the recorded events and returns total milliseconds and launch count of a family. */
enum { NERFCA_K_RAYS = 0, NERFCA_K_PACK = 1, NERFCA_K_FIELD_FWD = 2, NERFCA_K_LOSS = 3, NERFCA_K_FIELD_BWD = 4,
       NERFCA_K_ADAM = 5, NERFCA_K_COUNT = 6 };
/* (e) multi-GPU (nothing upstream): the step's only exchange -- the sum of the ranks' flat gradient buffers --
 * fused with the optimizer step of nerfca_adam_step, over NVLink / NVSwitch peer memory instead of a separate collective.
 * Every rank's gradient buffer [n] and a signal pad (>= 1 KB of uint32, zero before the first step) must be peer-mapped;
 * `grads` / `signals` are DEVICE arrays of world_size peer pointers in rank order, `own_signals` this rank's pad.
 * epoch = 1, 2, 3, ... must advance by one per call on every rank.  After the call (stream order) the parameters of all
 * ranks are bit-identical and this rank's gradient buffer is zero.  The grid of the update kernel must be co-resident with
 * its peers' (it is: n / 1024 blocks), all waits are bounded.  terms_out != NULL: the step's NERFCA_N_LOSS_TERMS float64 loss
 * sums of every rank, stored in the peer-mapped buffer at float offset terms_offset (>= n, even) behind the gradients, are
 * summed over the ranks in the same exchange (entries 2, 3: maxima) into terms_out (device, local) -- no second collective.   */
typedef struct nerfca_peers_t {
  int32_t rank, world_size;
  const float* const* grads;
  uint32_t* const* signals;
  uint32_t* own_signals;
} nerfca_peers_t;
NERFCA_API int nerfca_allreduce_adam_step(const nerfca_peers_t* peers, uint32_t epoch, float* params, float* grads, float* exp_avg,
                               float* exp_avg_sq, int64_t n, const nerfca_adam_step_t* cfg, const nerfca_repack_t* repack,
                               int64_t terms_offset, double* terms_out, void* stream);

/* CUDA-graph replay of a training step (nothing upstream; removes the per-kernel launch gaps of run_composite.py:283-308's
 * eager sequence).  Bracket the step's calls with begin / end_launch on a NON-default stream: the launches are captured, the
 * previous step's executable graph is updated in place (new pointers / per-step scalars) and launched once.  abort leaves
 * capture mode after a failed call inside the bracket.                                                                      */
typedef struct nerfca_graph nerfca_graph_t;
NERFCA_API int nerfca_graph_create(nerfca_graph_t** out);
NERFCA_API int nerfca_graph_begin(nerfca_graph_t* g, void* stream);
NERFCA_API int nerfca_graph_end_launch(nerfca_graph_t* g, void* stream);
NERFCA_API int nerfca_graph_abort(nerfca_graph_t* g, void* stream);
NERFCA_API int nerfca_graph_stats(const nerfca_graph_t* g, int64_t* launches, int64_t* updates, int64_t* instantiations);
NERFCA_API int nerfca_graph_destroy(nerfca_graph_t* g);

/* Parity / debug entry: the bf16 first-layer input tile X0 exactly as the tcgen05 kernels build it in registers (A5 on the hot
 * path: range-reduced MUFU sin/cos + double-angle steps).  x0_out: device uint16 [roundup128(P), kpad0] bf16 bit patterns
 * (may be NULL to query kpad0 only); columns: in_dim features, then the constant-1 bias column, then (onehot != 0 and they fit)
 * one-hot phase columns, then zero padding.                                                                                */
NERFCA_API int nerfca_debug_x0(const nerfca_field_t* field, const nerfca_samples_t* samples, int32_t onehot, uint16_t* x0_out,
                    int32_t* kpad0_out, void* stream);

NERFCA_API int64_t nerfca_launch_count(void);
NERFCA_API int nerfca_profile_enable(int32_t on);
NERFCA_API int nerfca_profile_read(int32_t kind, double* ms_total, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* NERFCA_H_ */
