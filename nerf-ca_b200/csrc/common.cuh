// Shared device helpers for the NeRF-CA B200 kernels: error plumbing, sample-point formation (A4),
// positional-encoding features (A5) and the output activations (A9).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/nerfca.h"

namespace nerfca {

void set_error(const std::string& msg);

#define NERFCA_REQUIRE(cond, code, msg)                                        \
  do {                                                                         \
    if (!(cond)) {                                                             \
      ::nerfca::set_error(std::string(__func__) + ": " + (msg));               \
      return (code);                                                           \
    }                                                                          \
  } while (0)

#define NERFCA_CUDA_OK(expr)                                                                   \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      ::nerfca::set_error(std::string(__func__) + ": " #expr " -> " + cudaGetErrorString(e__)); \
      return NERFCA_E_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

#define NERFCA_LAUNCH_OK()                                                                       \
  do {                                                                                           \
    ::nerfca::note_launch();                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess) {                                                                    \
      ::nerfca::set_error(std::string(__func__) + ": kernel launch -> " + cudaGetErrorString(e__)); \
      return NERFCA_E_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

// ---- launch accounting + optional per-kernel device timing (api.cu) -----------------------------------------
void note_launch();
// Brackets the launches issued inside its lifetime with two CUDA events on the launching stream when profiling is
// enabled (nerfca_profile_enable); a no-op otherwise.  kind = NERFCA_K_* of include/nerfca.h.
struct ProfScope {
  ProfScope(int kind, cudaStream_t st);
  ~ProfScope();
  int kind_;
  cudaStream_t st_;
  void* rec_;
};

inline int enc_dim_of(const nerfca_field_t& f) {
  if (f.enc_mode == NERFCA_ENC_NONE || f.n_freq <= 0) return 3;
  if (f.enc_mode == NERFCA_ENC_FOURIER) return 6 * f.n_freq;
  return 3 + 6 * f.n_freq;
}
inline int in_dim_of(const nerfca_field_t& f) { return enc_dim_of(f) + f.n_latent; }
int validate_field(const nerfca_field_t* f);
int validate_samples(const nerfca_samples_t* s, bool need_phase);

// ---- plain-old-data views passed to kernels by value --------------------------------------------------
struct SampleSrc {
  const float* points;
  const void* origins;
  const void* dirs;
  const float* depth;
  const int32_t* phase_point;
  const int32_t* phase_ray;
  long long n_points;  // samples visible to a launch: local index q in [0, n_points) maps to absolute sample base + q
  long long base;
  int n_rays, n_depth, ray_f64, ray_stride;
};
inline SampleSrc make_src(const nerfca_samples_t& s, long long base = 0, long long count = -1) {
  SampleSrc r;
  r.points = s.points; r.origins = s.origins; r.dirs = s.dirs; r.depth = s.depth;
  r.phase_point = s.phase_point; r.phase_ray = s.phase_ray; r.n_points = (count < 0) ? s.n_points : count; r.base = base;
  r.n_rays = s.n_rays; r.n_depth = s.n_depth; r.ray_f64 = (s.ray_dtype == NERFCA_F64); r.ray_stride = s.ray_stride;
  return r;
}

struct EncDesc {
  const float* band_weight;
  const float* fourier_coeff;
  const float* latents;
  int mode, n_freq, n_latent, n_phases, enc_dim, in_dim;
};
inline EncDesc make_enc(const nerfca_field_t& f) {
  EncDesc e;
  e.band_weight = f.band_weight; e.fourier_coeff = f.fourier_coeff; e.latents = f.latents;
  e.mode = (f.n_freq <= 0) ? NERFCA_ENC_NONE : f.enc_mode; e.n_freq = f.n_freq; e.n_latent = f.n_latent;
  e.n_phases = f.n_phases; e.enc_dim = enc_dim_of(f); e.in_dim = in_dim_of(f);
  return e;
}

// ---- A4: sample position p -> (x,y,z) float32, bit-identical to the reference -------------------------
// training (float64 rays):  fl32( o + d * fl64(z) ) with separately rounded f64 multiply and add
//                           (train/model_helpers.py:118-120: f64 tensor ops, then .float())
// eval (float32 rays):      o + fl32(d * z)        (train/run_composite.py:351)
__device__ __forceinline__ void load_point(const SampleSrc& s, long long q, float& x, float& y, float& z) {
  const long long p = q + s.base;
  if (s.points) {
    x = __ldg(s.points + 3 * p); y = __ldg(s.points + 3 * p + 1); z = __ldg(s.points + 3 * p + 2);
    return;
  }
  // 32-bit index math whenever the sample index fits (a 64-bit division costs ~80 instructions)
  const int ray = (p < 0x7fffffffLL) ? (int)((unsigned)p / (unsigned)s.n_depth) : (int)(p / s.n_depth);
  const int k = (int)(p - (long long)ray * s.n_depth);
  const float t = __ldg(s.depth + k);
  if (s.ray_f64) {
    const double* o = (const double*)s.origins + (size_t)ray * s.ray_stride;
    const double* d = (const double*)s.dirs + (size_t)ray * s.ray_stride;
    const double td = (double)t;
    x = __double2float_rn(__dadd_rn(__ldg(o + 0), __dmul_rn(__ldg(d + 0), td)));
    y = __double2float_rn(__dadd_rn(__ldg(o + 1), __dmul_rn(__ldg(d + 1), td)));
    z = __double2float_rn(__dadd_rn(__ldg(o + 2), __dmul_rn(__ldg(d + 2), td)));
  } else {
    const float* o = (const float*)s.origins + (size_t)ray * s.ray_stride;
    const float* d = (const float*)s.dirs + (size_t)ray * s.ray_stride;
    x = __fadd_rn(__ldg(o + 0), __fmul_rn(__ldg(d + 0), t));
    y = __fadd_rn(__ldg(o + 1), __fmul_rn(__ldg(d + 1), t));
    z = __fadd_rn(__ldg(o + 2), __fmul_rn(__ldg(d + 2), t));
  }
}

__device__ __forceinline__ int load_phase(const SampleSrc& s, long long q) {
  const long long p = q + s.base;
  if (s.phase_point) return __ldg(s.phase_point + p);
  if (s.phase_ray) return __ldg(s.phase_ray + ((p < 0x7fffffffLL) ? (int)((unsigned)p / (unsigned)s.n_depth) : (int)(p / s.n_depth)));
  return 0;
}

// ---- A5: one input feature of the first layer -----------------------------------------------------------
// Feature order of the BANDS family (model/CPPN.py:120-133): [x, y, z] then for band l:
// sin(2^l x), sin(2^l y), sin(2^l z), sin(2^l x + pi/2), ... ; every band feature times band_weight[l].
// The "+ pi/2" is an fp32 add of fl32(pi/2) to the fp32 product, exactly as torch evaluates it.
__device__ __forceinline__ float enc_feature(const EncDesc& e, int f, float x, float y, float z, int phase) {
  if (f >= e.enc_dim) {  // latent concat, model/Temporal.py:124,144-147
    const int t = f - e.enc_dim;
    if (phase < 0 || phase >= e.n_phases) phase = 0;       // same convention as the tensor-core kernels (emit_x0_row): never read outside the table
    return (t < e.n_latent) ? __ldg(e.latents + (size_t)phase * e.n_latent + t) : 0.f;
  }
  if (e.mode == NERFCA_ENC_NONE) return f == 0 ? x : (f == 1 ? y : z);
  if (e.mode == NERFCA_ENC_FOURIER) {
    const int half = 3 * e.n_freq;
    const int g = (f < half) ? f : f - half;
    const int c = g % 3;
    const float v = c == 0 ? x : (c == 1 ? y : z);
    // 2 * np.pi * basis_values * coeff: the python double 2*pi multiplies the fp32 tensor (scalar cast to fp32)
    const float arg = __fmul_rn(__fmul_rn(6.28318530717958647692f, v), __ldg(e.fourier_coeff + g));
    return (f < half) ? sinf(arg) : cosf(arg);
  }
  if (f < 3) return f == 0 ? x : (f == 1 ? y : z);
  const int g = f - 3;
  const int band = g / 6;
  const int r = g - band * 6;
  const int c = r % 3;
  const float v = c == 0 ? x : (c == 1 ? y : z);
  float arg = __fmul_rn(v, __int_as_float((127 + band) << 23));  // x * 2^band, exact
  if (r >= 3) arg = __fadd_rn(arg, 1.57079637050628662109375f);    // + fl32(0.5 * pi)
  float val = sinf(arg);
  if (e.band_weight) val = __fmul_rn(__ldg(e.band_weight + band), val);
  return val;
}

// ---- A9: output activations (train/model_helpers.py:63-70) ----------------------------------------------
__device__ __forceinline__ float softplus_f(float v) { return v > 20.f ? v : log1pf(expf(v)); }  // beta=1, threshold=20
__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }
__device__ __forceinline__ float act_fwd(int act, float v) {
  if (act == NERFCA_ACT_SOFTPLUS) return softplus_f(v);
  if (act == NERFCA_ACT_CLAMP) return fminf(fmaxf(softplus_f(v), 0.f), 1.f);
  return sigmoid_f(v);
}
// d act / d raw
__device__ __forceinline__ float act_bwd(int act, float v) {
  if (act == NERFCA_ACT_SOFTPLUS) return v > 20.f ? 1.f : sigmoid_f(v);
  if (act == NERFCA_ACT_CLAMP) {
    const float sp = softplus_f(v);
    return (sp > 0.f && sp < 1.f) ? (v > 20.f ? 1.f : sigmoid_f(v)) : 0.f;
  }
  const float s = sigmoid_f(v);
  return s * (1.f - s);
}

inline unsigned div_up(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

}  // namespace nerfca
