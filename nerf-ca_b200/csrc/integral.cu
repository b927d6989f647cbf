// X-ray line integral (A9), its autograd, and the fused training loss + closed-form dL/d_raw (A9 + A10).
// One warp per ray (one CTA per ray in the fused loss kernel); ray sums by warp shuffle; HBM traffic = the raw field outputs in,
// sigma / gradients out.
#include <stdlib.h>

#include "common.cuh"

namespace nerfca {

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// delta_s = z[s+1] - z[s] (fp32 subtraction), last = 1e-10 in the accumulation dtype (model_helpers.py:73-74)
template <typename ACC>
__device__ __forceinline__ ACC delta_at(const float* __restrict__ z, int s, int n) {
  return (s == n - 1) ? (ACC)1e-10 : (ACC)__fsub_rn(__ldg(z + s + 1), __ldg(z + s));
}

template <typename ACC>
__global__ void integrate_kernel(const float* __restrict__ raw_s, const float* __restrict__ raw_d,
                                 const float* __restrict__ z, const float* __restrict__ i0, int n_rays, int n, int act,
                                 ACC* __restrict__ pix, float* __restrict__ sig_s, float* __restrict__ sig_d,
                                 ACC* __restrict__ dists) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_rays) return;
  const size_t off = (size_t)warp * n;
  ACC acc = 0;
  for (int s = lane; s < n; s += 32) {
    const ACC d = delta_at<ACC>(z, s, n);
    if (raw_d) {  // composite: sigma = act(raw) * 1e-2 (fp32), weights = (ss + sd) * delta
      const float ss = __fmul_rn(act_fwd(act, raw_s[off + s]), 0.01f);
      const float sd = __fmul_rn(act_fwd(act, raw_d[off + s]), 0.01f);
      sig_s[off + s] = ss;
      sig_d[off + s] = sd;
      acc += (ACC)__fadd_rn(ss, sd) * d;
    } else {      // single field: sigma unscaled, weights = sigma * delta * 1e-2
      const float sg = act_fwd(act, raw_s[off + s]);
      sig_s[off + s] = sg;
      acc += ((ACC)sg * d) * (ACC)1e-2;
    }
    if (warp == 0 && dists) dists[s] = d;
  }
  acc = warp_sum(acc);
  if (lane == 0) pix[warp] = (ACC)__ldg(i0 + warp) - acc;
}

template <typename ACC>
__global__ void integrate_bwd_kernel(const float* __restrict__ raw_s, const float* __restrict__ raw_d,
                                     const float* __restrict__ z, int n_rays, int n, int act, const ACC* __restrict__ d_pix,
                                     const float* __restrict__ d_sig_s, const float* __restrict__ d_sig_d,
                                     float* __restrict__ d_raw_s, float* __restrict__ d_raw_d) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_rays * n) return;
  const int ray = (int)(idx / n), s = (int)(idx - (long long)ray * n);
  const ACC d = delta_at<ACC>(z, s, n);
  const ACC gp = d_pix ? d_pix[ray] : (ACC)0;
  if (raw_d) {
    const ACC common = -gp * d;
    const ACC gs = common + (d_sig_s ? (ACC)d_sig_s[idx] : (ACC)0);
    const ACC gd = common + (d_sig_d ? (ACC)d_sig_d[idx] : (ACC)0);
    d_raw_s[idx] = (float)(gs * (ACC)0.01f) * act_bwd(act, raw_s[idx]);
    d_raw_d[idx] = (float)(gd * (ACC)0.01f) * act_bwd(act, raw_d[idx]);
  } else {
    const ACC gs = -gp * d * (ACC)1e-2 + (d_sig_s ? (ACC)d_sig_s[idx] : (ACC)0);
    d_raw_s[idx] = (float)gs * act_bwd(act, raw_s[idx]);
  }
}

struct LossCfg {
  double c_f, c_e, c_o, c_l, mask_thre, w_thresh;
  int use_weighting, b_global;
};

// Fused A9 + A10 + closed-form backward (SURVEY 8(a')), training dtypes (float64 ray sums).
// One CTA of LOSS_THREADS threads per ray (a warp per ray leaves the SMs at ~10 % occupancy for 1024 rays, and the fp64 logarithms
// make the kernel latency bound); sigma_s / sigma_d of the ray live in shared memory between the three sweeps, the ray sums go
// through warp shuffles and a small shared-memory exchange (fixed order: the result does not depend on scheduling).
// sums `v[0..N)` over the CTA; every thread returns with the totals.  scratch: N * (LOSS_THREADS / 32) doubles
template <int N, int LOSS_THREADS>
__device__ __forceinline__ void block_sum(double (&v)[N], double* scratch) {
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();                       // the scratch area of the previous exchange has been read
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) scratch[k * (LOSS_THREADS / 32) + wib] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < LOSS_THREADS / 32; ++w) t += scratch[k * (LOSS_THREADS / 32) + w];
    v[k] = t;
  }
}
template <int LOSS_THREADS>
__global__ void __launch_bounds__(LOSS_THREADS) composite_loss_kernel(const float* __restrict__ raw_s, const float* __restrict__ raw_d,
                                      const float* __restrict__ z, const float* __restrict__ i0,
                                      const double* __restrict__ gt, const double* __restrict__ wpix, int gw_stride,
                                      int n_rays, int n, int act, LossCfg c, double* __restrict__ pix_out,
                                      double* __restrict__ terms, float* __restrict__ d_raw_s, float* __restrict__ d_raw_d) {
  extern __shared__ double sm_d[];
  double* scratch = sm_d;                                   // 8 * (LOSS_THREADS / 32) doubles
  float* ss_c = reinterpret_cast<float*>(sm_d + 8 * (256 / 32));
  float* sd_c = ss_c + n;
  const int ray = blockIdx.x, lane = threadIdx.x;           // `lane`: index of the thread within the ray's CTA
  if (ray >= n_rays) return;
  const size_t off = (size_t)ray * n;
  const double B = (double)c.b_global, BN = B * (double)n;

  // sweep 1: sigma, ray sums, blend-ratio terms
  double W = 0, Ss = 0, Sd = 0, l2 = 0, bw_sum = 0, fav_sum = 0;
  float mx_s = 0.f, mx_d = 0.f;
  for (int s = lane; s < n; s += LOSS_THREADS) {
    const double d = delta_at<double>(z, s, n);
    const float ss = __fmul_rn(act_fwd(act, raw_s[off + s]), 0.01f);
    const float sd = __fmul_rn(act_fwd(act, raw_d[off + s]), 0.01f);
    ss_c[s] = ss; sd_c[s] = sd;
    W += (double)__fadd_rn(ss, sd) * d;
    const double as = (double)ss * d, ad = (double)sd * d;
    Ss += as; Sd += ad; l2 += as * as;
    mx_s = fmaxf(mx_s, ss); mx_d = fmaxf(mx_d, sd);
    // compute_ratio / compute_blendw_loss, fp32 like the reference (model_helpers.py:189-204)
    const float bw = sd / (__fadd_rn(__fadd_rn(ss, sd), 1e-10f));
    const float b = fminf(fmaxf(bw, 1e-19f), 1.0f);  // fl32(1 - 1e-19) == 1
    const float r = fmaxf(1.f - b, 1e-19f);
    bw_sum += (double)bw;
    fav_sum += (double)(-(b * logf(b) + r * logf(r)));
  }
  {
    // the six ray sums in one exchange, the two maxima in a second one
    double v[6] = {W, Ss, Sd, l2, bw_sum, fav_sum};
    block_sum<6, LOSS_THREADS>(v, scratch);
    W = v[0]; Ss = v[1]; Sd = v[2]; l2 = v[3]; bw_sum = v[4]; fav_sum = v[5];
    mx_s = warp_max(mx_s); mx_d = warp_max(mx_d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { scratch[2 * (threadIdx.x >> 5)] = (double)mx_s; scratch[2 * (threadIdx.x >> 5) + 1] = (double)mx_d; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < LOSS_THREADS / 32; ++w) { mx_s = fmaxf(mx_s, (float)scratch[2 * w]); mx_d = fmaxf(mx_d, (float)scratch[2 * w + 1]); }
  }

  const double w = wpix[(size_t)ray * gw_stride];
  const double pix = (double)__ldg(i0 + ray) - W;
  const double res = pix - gt[(size_t)ray * gw_stride];
  const double Ss_hat = fmax(Ss, 1e-19), Sd_hat = fmax(Sd, 1e-19);
  const bool mask_s = !(Ss < c.mask_thre);
  const bool mask_d = !(Sd < c.mask_thre) || (c.use_weighting && w > 1.0 + c.w_thresh);

  // sweep 2: ray entropies and the sum_j h_j p_j normaliser of the dynamic-entropy gradient
  double ent_s = 0, ent_d = 0, hp_d = 0;
  for (int s = lane; s < n; s += LOSS_THREADS) {
    const double d = delta_at<double>(z, s, n);
    const double ps = (double)ss_c[s] * d / Ss_hat, pd = (double)sd_c[s] * d / Sd_hat;
    ent_s -= ps * log(ps + 1e-10);
    const double lg = log(pd + 1e-10);
    ent_d -= pd * lg;
    hp_d += -(lg + pd / (pd + 1e-10)) * pd;
  }
  {
    double v[3] = {ent_s, ent_d, hp_d};
    block_sum<3, LOSS_THREADS>(v, scratch);
    ent_s = v[0]; ent_d = v[1]; hp_d = v[2];
  }

  if (lane == 0) {
    pix_out[ray] = pix;
    atomicAdd(terms + NERFCA_T_PIXEL_SUM, w * res * res);
    atomicAdd(terms + NERFCA_T_BLENDW_SUM, bw_sum);
    atomicAdd(terms + NERFCA_T_FAVOR_SUM, fav_sum);
    atomicAdd(terms + NERFCA_T_S_ENT_SUM, mask_s ? ent_s : 0.0);
    atomicAdd(terms + NERFCA_T_S_SUM_SUM, Ss);
    atomicAdd(terms + NERFCA_T_D_ENT_SUM, mask_d ? ent_d : 0.0);
    atomicAdd(terms + NERFCA_T_D_SUM_SUM, Sd);
    atomicAdd(terms + NERFCA_T_OCCL_SUM, Sd);
    atomicAdd(terms + NERFCA_T_L1_SUM, Ss);
    atomicAdd(terms + NERFCA_T_L2_SUM, l2);
    // sigma >= 0, so the maxima can be taken on the bit patterns of non-negative doubles
    atomicMax((unsigned long long*)(terms + NERFCA_T_SIGMA_S_MAX), (unsigned long long)__double_as_longlong((double)mx_s));
    atomicMax((unsigned long long*)(terms + NERFCA_T_SIGMA_D_MAX), (unsigned long long)__double_as_longlong((double)mx_d));
  }
  if (!d_raw_s) return;

  // sweep 3: dL/d_raw
  const double g_px_c = -(2.0 / B) * w * res;
  const double live_d = (Sd >= 1e-19) ? 1.0 : 0.0;
  for (int s = lane; s < n; s += LOSS_THREADS) {
    const double d = delta_at<double>(z, s, n);
    const float ssf = ss_c[s], sdf = sd_c[s];
    const double ss = ssf, sd = sdf;
    const double g_px = g_px_c * d;
    // blend entropy
    const double tot = (double)__fadd_rn(__fadd_rn(ssf, sdf), 1e-10f);
    const float bwf = sdf / (float)tot;
    double e1 = 0.0;
    if (bwf >= 1e-19f && bwf <= 1.0f) {
      // (single-precision logarithms in this sweep: these regulariser gradients carry weights of 1e-12 .. 1e-4 against the pixel term,
      // so a 1e-7 relative error in them is far below the fp32 rounding of d_raw itself; a double log costs ~10x as much here)
      const double b = fmin(fmax((double)bwf, 1e-19), 1.0);
      const double omb = 1.0 - b;
      e1 = -((double)logf((float)b) + 1.0);
      if (omb >= 1e-19) e1 += (double)logf((float)omb) + 1.0;
    }
    const double kf = c.c_f / BN * e1 / (tot * tot);
    const double g_fd = kf * (ss + 1e-10), g_fs = kf * (-sd);
    // dynamic ray entropy
    const double pd = sd * d / Sd_hat;
    const double h = -((double)logf((float)(pd + 1e-10)) + pd / (pd + 1e-10));
    const double g_ed = mask_d ? c.c_e / B * (h - live_d * hp_d) / Sd_hat * d : 0.0;
    const double g_od = c.c_o / B * d;
    const double g_ls = c.c_l * (d + 2.0 * ss * d * d);
    const double gs = g_px + g_fs + g_ls, gd = g_px + g_fd + g_ed + g_od;
    d_raw_s[off + s] = (float)(gs * 0.01) * act_bwd(act, raw_s[off + s]);
    d_raw_d[off + s] = (float)(gd * 0.01) * act_bwd(act, raw_d[off + s]);
  }
}

// static run of run_nerf.py:227-230: loss = mean(w (pix-gt)^2) + occl_weight * mean_r sum_s sigma delta, sigma unscaled
__global__ void static_loss_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ i0,
                                   const double* __restrict__ gt, const double* __restrict__ wpix, int gw_stride, int n_rays,
                                   int n, int act, LossCfg c, double* __restrict__ pix_out, double* __restrict__ terms,
                                   float* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n_rays) return;
  const size_t off = (size_t)ray * n;
  double S = 0;
  float mx = 0.f;
  for (int s = lane; s < n; s += 32) {
    const float sg = act_fwd(act, raw[off + s]);
    S += (double)sg * delta_at<double>(z, s, n);
    mx = fmaxf(mx, sg);
  }
  S = warp_sum(S); mx = warp_max(mx);
  const double w = wpix[(size_t)ray * gw_stride];
  const double pix = (double)__ldg(i0 + ray) - S * 1e-2;
  const double res = pix - gt[(size_t)ray * gw_stride];
  if (lane == 0) {
    pix_out[ray] = pix;
    atomicAdd(terms + NERFCA_T_PIXEL_SUM, w * res * res);
    atomicAdd(terms + NERFCA_T_OCCL_SUM, S);
    atomicMax((unsigned long long*)(terms + NERFCA_T_SIGMA_S_MAX), (unsigned long long)__double_as_longlong((double)mx));
  }
  if (!d_raw) return;
  const double B = (double)c.b_global;
  const double k = -(2.0 / B) * w * res * 1e-2 + c.c_o / B;
  for (int s = lane; s < n; s += 32)
    d_raw[off + s] = (float)(k * delta_at<double>(z, s, n)) * act_bwd(act, raw[off + s]);
}

// pixels from the per-ray attenuation sums the render forward accumulated (A9 eval form, fp32): composite, static-only, dynamic-only
__global__ void render_finalize_kernel(const float* __restrict__ sum_s, const float* __restrict__ sum_d, const float* __restrict__ i0,
                                       int n_rays, float* __restrict__ pix, float* __restrict__ pix_s, float* __restrict__ pix_d) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const float base = __ldg(i0 + r), a = sum_s[r], b = sum_d ? sum_d[r] : 0.f;
  pix[r] = base - (a + b);
  if (pix_s) pix_s[r] = base - a;
  if (pix_d) pix_d[r] = base - b;
}
int launch_render_finalize(const float* sum_s, const float* sum_d, const float* i0, int n_rays, float* pix, float* pix_s, float* pix_d,
                           cudaStream_t st) {
  render_finalize_kernel<<<div_up(n_rays, 256), 256, 0, st>>>(sum_s, sum_d, i0, n_rays, pix, pix_s, pix_d);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

// ---- N4: display normalisation of an image, run_composite.py:394-413: (img - min) / (max - min) ------------------------------------
__device__ __forceinline__ unsigned f2ord(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
__global__ void minmax_kernel(const float* __restrict__ img, long long n, unsigned* __restrict__ mm) {
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = img[i];
    lo = fminf(lo, v); hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0) { atomicMin(mm, f2ord(lo)); atomicMax(mm + 1, f2ord(hi)); }
}
__global__ void normalize_kernel(const float* __restrict__ img, long long n, const unsigned* __restrict__ mm, float* __restrict__ out,
                                 float* __restrict__ minmax_out) {
  const float lo = ord2f(mm[0]), hi = ord2f(mm[1]);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && minmax_out) { minmax_out[0] = lo; minmax_out[1] = hi; }
  if (i < n) out[i] = __fdiv_rn(__fsub_rn(img[i], lo), __fsub_rn(hi, lo));
}

}  // namespace nerfca

using namespace nerfca;

extern "C" int nerfca_normalize_image(const float* img, int64_t n, float* out, float* minmax_out, void* scratch8, void* stream) {
  NERFCA_REQUIRE(img && out && scratch8, NERFCA_E_ARG, "null pointer");
  if (n <= 0) return NERFCA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* mm = (unsigned*)scratch8;
  const unsigned init[2] = {0xFFFFFFFFu, 0u};
  NERFCA_CUDA_OK(cudaMemsetAsync(mm, 0xFF, 4, st));
  NERFCA_CUDA_OK(cudaMemsetAsync(mm + 1, 0, 4, st));
  (void)init;
  const unsigned blocks = div_up(n, 256) < 1184u ? div_up(n, 256) : 1184u;     // 8 CTAs per SM at most
  minmax_kernel<<<blocks, 256, 0, st>>>(img, (long long)n, mm);
  NERFCA_LAUNCH_OK();
  normalize_kernel<<<div_up(n, 256), 256, 0, st>>>(img, (long long)n, mm, out, minmax_out);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

extern "C" int nerfca_integrate(const float* raw_s, const float* raw_d, const float* depth, const float* i0, int32_t n_rays,
                                int32_t n_depth, int32_t activation, int32_t acc_dtype, void* pix_out, float* sigma_s_out,
                                float* sigma_d_out, void* dists_out, void* stream) {
  NERFCA_REQUIRE(raw_s && depth && i0 && pix_out && sigma_s_out, NERFCA_E_ARG, "null pointer");
  NERFCA_REQUIRE(!raw_d || sigma_d_out, NERFCA_E_ARG, "sigma_d_out required with raw_d");
  NERFCA_REQUIRE(n_depth > 0, NERFCA_E_ARG, "n_depth must be positive");
  if (n_rays <= 0) return NERFCA_OK;
  const unsigned blocks = div_up((long long)n_rays * 32, 256);
  if (acc_dtype == NERFCA_F64)
    integrate_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(raw_s, raw_d, depth, i0, n_rays, n_depth, activation,
                                                                      (double*)pix_out, sigma_s_out, sigma_d_out,
                                                                      (double*)dists_out);
  else
    integrate_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(raw_s, raw_d, depth, i0, n_rays, n_depth, activation,
                                                                     (float*)pix_out, sigma_s_out, sigma_d_out,
                                                                     (float*)dists_out);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

extern "C" int nerfca_integrate_backward(const float* raw_s, const float* raw_d, const float* depth, int32_t n_rays,
                                         int32_t n_depth, int32_t activation, int32_t acc_dtype, const void* d_pix,
                                         const float* d_sigma_s, const float* d_sigma_d, float* d_raw_s, float* d_raw_d,
                                         void* stream) {
  NERFCA_REQUIRE(raw_s && depth && d_raw_s, NERFCA_E_ARG, "null pointer");
  NERFCA_REQUIRE(!raw_d || d_raw_d, NERFCA_E_ARG, "d_raw_d required with raw_d");
  const long long total = (long long)n_rays * n_depth;
  if (total <= 0) return NERFCA_OK;
  if (acc_dtype == NERFCA_F64)
    integrate_bwd_kernel<double><<<div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        raw_s, raw_d, depth, n_rays, n_depth, activation, (const double*)d_pix, d_sigma_s, d_sigma_d, d_raw_s, d_raw_d);
  else
    integrate_bwd_kernel<float><<<div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(
        raw_s, raw_d, depth, n_rays, n_depth, activation, (const float*)d_pix, d_sigma_s, d_sigma_d, d_raw_s, d_raw_d);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

extern "C" int nerfca_composite_loss(const float* raw_s, const float* raw_d, const float* depth, const float* i0,
                                     const double* gt, const double* wpix, int32_t gw_stride, int32_t n_rays, int32_t n_depth,
                                     int32_t activation, const nerfca_loss_cfg_t* cfg, double* pix_out, double* terms_out,
                                     float* d_raw_s, float* d_raw_d, void* stream) {
  NERFCA_REQUIRE(raw_s && depth && i0 && gt && wpix && cfg && pix_out && terms_out, NERFCA_E_ARG, "null pointer");
  NERFCA_REQUIRE(n_depth > 0 && gw_stride > 0 && cfg->n_rays_global > 0, NERFCA_E_ARG, "bad sizes");
  NERFCA_REQUIRE(!raw_d || !d_raw_s || d_raw_d, NERFCA_E_ARG, "d_raw_d required with raw_d");
  if (n_rays <= 0) return NERFCA_OK;
  LossCfg c;
  c.c_f = cfg->favor_s_weight; c.c_e = cfg->dyn_entropy_weight; c.c_o = cfg->occl_weight; c.c_l = cfg->l1_weight;
  c.mask_thre = cfg->entro_mask_thre; c.w_thresh = cfg->entro_weighted_thresh; c.use_weighting = cfg->entro_use_weighting;
  c.b_global = cfg->n_rays_global;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(NERFCA_K_LOSS, st);
  if (!raw_d) {
    static_loss_kernel<<<div_up((long long)n_rays * 32, 128), 128, 0, st>>>(raw_s, depth, i0, gt, wpix, gw_stride, n_rays,
                                                                            n_depth, activation, c, pix_out, terms_out, d_raw_s);
    NERFCA_LAUNCH_OK();
    return NERFCA_OK;
  }
  // one CTA of 128 threads per ray (256, NERFCA_LOSS_THREADS=256, measured no faster: 27.0 vs 25.8 us, r3k): the kernel is bound by its three
  // dependent sweeps with fp64 logarithms, not by throughput
  static const int loss_threads = (getenv("NERFCA_LOSS_THREADS") && atoi(getenv("NERFCA_LOSS_THREADS")) == 256) ? 256 : 128;
  const size_t smem = 8 * (256 / 32) * sizeof(double) + (size_t)2 * n_depth * sizeof(float);
  NERFCA_REQUIRE(smem <= 200 * 1024, NERFCA_E_UNSUPPORTED, "n_depth too large for the fused loss kernel");
  if (loss_threads == 128) {
    if (smem > 48 * 1024)
      NERFCA_CUDA_OK(cudaFuncSetAttribute(composite_loss_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    composite_loss_kernel<128><<<n_rays, 128, smem, st>>>(raw_s, raw_d, depth, i0, gt, wpix, gw_stride, n_rays, n_depth, activation, c,
                                                        pix_out, terms_out, d_raw_s, d_raw_d);
  } else {
    if (smem > 48 * 1024)
      NERFCA_CUDA_OK(cudaFuncSetAttribute(composite_loss_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    composite_loss_kernel<256><<<n_rays, 256, smem, st>>>(raw_s, raw_d, depth, i0, gt, wpix, gw_stride, n_rays, n_depth, activation, c,
                                                        pix_out, terms_out, d_raw_s, d_raw_d);
  }
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}
