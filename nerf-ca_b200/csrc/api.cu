// C-ABI plumbing: error state, argument validation and precision dispatch of the field entry points.
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace nerfca {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

// ---- launch accounting / profiling ----------------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { int kind; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;            // recorded pairs
static std::vector<ProfRec> g_prof_pool;       // reusable event pairs

ProfScope::ProfScope(int kind, cudaStream_t st) : kind_(kind), st_(st), rec_(nullptr) {
  if (!g_prof_on) return;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return;   // no event timing inside a graph capture
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec* r = new ProfRec;
  if (!g_prof_pool.empty()) { *r = g_prof_pool.back(); g_prof_pool.pop_back(); }
  else { cudaEventCreate(&r->e0); cudaEventCreate(&r->e1); }
  r->kind = kind;
  cudaEventRecord(r->e0, st);
  rec_ = r;
}
ProfScope::~ProfScope() {
  if (!rec_) return;
  ProfRec* r = static_cast<ProfRec*>(rec_);
  cudaEventRecord(r->e1, st_);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(*r);
  delete r;
}

int validate_field(const nerfca_field_t* f) {
  NERFCA_REQUIRE(f != nullptr, NERFCA_E_ARG, "field is null");
  NERFCA_REQUIRE(f->enc_mode >= NERFCA_ENC_NONE && f->enc_mode <= NERFCA_ENC_FOURIER, NERFCA_E_ARG, "bad enc_mode");
  NERFCA_REQUIRE(f->n_freq >= 0 && f->n_freq <= 32, NERFCA_E_ARG, "n_freq out of range [0,32]");
  NERFCA_REQUIRE(f->hidden > 0 && f->hidden <= 1024, NERFCA_E_ARG, "hidden out of range");
  NERFCA_REQUIRE(f->n_hidden >= 0 && f->n_hidden + 2 <= NERFCA_MAX_LAYERS, NERFCA_E_ARG, "too many layers");
  NERFCA_REQUIRE(f->n_latent >= 0 && f->n_latent <= 64, NERFCA_E_ARG, "n_latent out of range [0,64]");
  if (f->n_latent > 0) NERFCA_REQUIRE(f->latents && f->n_phases > 0, NERFCA_E_ARG, "latents missing");
  if (f->enc_mode == NERFCA_ENC_FOURIER && f->n_freq > 0)
    NERFCA_REQUIRE(f->fourier_coeff != nullptr, NERFCA_E_ARG, "fourier_coeff missing");
  for (int l = 0; l < f->n_hidden + 2; ++l) NERFCA_REQUIRE(f->weight[l] != nullptr, NERFCA_E_ARG, "weight pointer missing");
  return NERFCA_OK;
}

int validate_samples(const nerfca_samples_t* s, bool need_phase) {
  NERFCA_REQUIRE(s != nullptr, NERFCA_E_ARG, "samples is null");
  NERFCA_REQUIRE(s->n_points >= 0, NERFCA_E_ARG, "negative n_points");
  if (!s->points && s->n_points > 0) {
    NERFCA_REQUIRE(s->origins && s->dirs && s->depth, NERFCA_E_ARG, "neither points nor rays given");
    NERFCA_REQUIRE(s->n_rays > 0 && s->n_depth > 0 && (long long)s->n_rays * s->n_depth == s->n_points, NERFCA_E_ARG,
                   "n_points != n_rays * n_depth");
    NERFCA_REQUIRE(s->ray_stride >= 3, NERFCA_E_ARG, "ray_stride < 3");
    NERFCA_REQUIRE(s->ray_dtype == NERFCA_F32 || s->ray_dtype == NERFCA_F64, NERFCA_E_ARG, "bad ray_dtype");
  }
  if (need_phase && s->n_points > 0) {
    NERFCA_REQUIRE(s->phase_point || s->phase_ray, NERFCA_E_ARG, "phases missing for a field with latents");
    NERFCA_REQUIRE(s->phase_point || s->n_depth > 0, NERFCA_E_ARG, "phase_ray needs n_depth");
  }
  return NERFCA_OK;
}

// fp32 SIMT path (mlp_simt.cu)
size_t simt_stash_bytes(const nerfca_field_t& f, long long P);
size_t simt_workspace_bytes(const nerfca_field_t& f, long long P, int backward);
int simt_field_forward(const nerfca_field_t& f, const nerfca_samples_t& s, float* raw_out, void* stash, void* workspace,
                       cudaStream_t st);
int simt_field_backward(const nerfca_field_t& f, const nerfca_samples_t& s, const float* d_raw, const void* stash,
                        void* workspace, const nerfca_field_grads_t& gr, cudaStream_t st);
// bf16 layer-wise tcgen05 path for shapes outside the fused kernels (mlp_wide.cu)
int wide_supported(const nerfca_field_t& f);
size_t wide_stash_bytes(const nerfca_field_t& f, long long P);
size_t wide_workspace_bytes(const nerfca_field_t& f, long long P, int backward);
int wide_field_forward(const nerfca_field_t& f, const nerfca_samples_t& s, float* raw_out, void* stash, void* workspace, cudaStream_t st);
int wide_field_backward(const nerfca_field_t& f, const nerfca_samples_t& s, const float* d_raw, const void* stash, void* workspace,
                        const nerfca_field_grads_t& gr, cudaStream_t st);
// bf16 fused tcgen05 path (mlp_tc.cu)
bool tc_shape_ok(const nerfca_field_t& f, bool training);
int tc_supported(const nerfca_field_t& f);
size_t tc_stash_bytes(const nerfca_field_t& f, long long P);
size_t tc_workspace_bytes(const nerfca_field_t& f, long long P, int backward);
int tc_field_forward(const nerfca_field_t& f, const nerfca_samples_t& s, float* raw_out, void* stash, void* workspace,
                     cudaStream_t st);
int tc_field_backward(const nerfca_field_t& f, const nerfca_samples_t& s, const float* d_raw, const void* stash,
                      void* workspace, const nerfca_field_grads_t& gr, cudaStream_t st);
size_t tc_stash_bytes_n(int n_nets, long long P);
size_t tc_workspace_bytes_n(const nerfca_field_t* const* f, int n_nets, long long P, int backward);
int tc_fields_forward(const nerfca_field_t* const* f, int n_nets, const nerfca_samples_t& s, float* const* raw_out, void* stash,
                      void* workspace, int pack, double* zero_terms, cudaStream_t st, float* const* ray_sum = nullptr, int act = 0);
int launch_render_finalize(const float* sum_s, const float* sum_d, const float* i0, int n_rays, float* pix, float* pix_s, float* pix_d,
                           cudaStream_t st);
int tc_fields_backward(const nerfca_field_t* const* f, int n_nets, const nerfca_samples_t& s, const float* const* d_raw,
                       const void* stash, void* workspace, int pack, const nerfca_field_grads_t* const* gr, cudaStream_t st);

int tc_debug_x0(const nerfca_field_t& f, const nerfca_samples_t& s, int onehot, uint16_t* out, int* kpad0_out, cudaStream_t st);

static size_t up256(size_t n) { return (n + 255) & ~(size_t)255; }

// Which kernels serve a bf16 request: the fused chain kernels when every field has their shape, else the layer-wise GEMM path for ALL
// fields of the call (one stash / workspace format per call).  `training`: the call writes or reads an activation stash.
static bool bf16_wide(const nerfca_field_t* a, const nerfca_field_t* b, bool training) {
  return (a && !tc_shape_ok(*a, training)) || (b && !tc_shape_ok(*b, training));
}

}  // namespace nerfca

using namespace nerfca;

extern "C" const char* nerfca_last_error(void) { return g_last_error.c_str(); }
extern "C" int nerfca_abi_version(void) { return NERFCA_ABI_VERSION; }

extern "C" int64_t nerfca_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int nerfca_profile_enable(int32_t on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (on) {
    for (auto& r : g_prof) g_prof_pool.push_back(r);
    g_prof.clear();
  }
  g_prof_on = on != 0;
  return NERFCA_OK;
}

extern "C" int nerfca_profile_read(int32_t kind, double* ms_total, int64_t* launches) {
  NERFCA_REQUIRE(kind >= 0 && kind < NERFCA_K_COUNT && ms_total && launches, NERFCA_E_ARG, "bad kind / null pointer");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms = 0;
  long long n = 0;
  for (auto& r : g_prof) {
    if (r.kind != kind) continue;
    NERFCA_CUDA_OK(cudaEventSynchronize(r.e1));
    float t = 0.f;
    NERFCA_CUDA_OK(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t;
    ++n;
  }
  *ms_total = ms;
  *launches = n;
  return NERFCA_OK;
}

extern "C" size_t nerfca_field_stash_bytes(const nerfca_field_t* field, int64_t n_points, int32_t precision) {
  if (!field || n_points <= 0) return 0;
  if (precision == NERFCA_PREC_BF16) return bf16_wide(field, nullptr, true) ? wide_stash_bytes(*field, n_points) : tc_stash_bytes(*field, n_points);
  return simt_stash_bytes(*field, n_points);
}

extern "C" size_t nerfca_field_workspace_bytes(const nerfca_field_t* field, int64_t n_points, int32_t precision,
                                               int32_t backward) {
  if (!field || n_points <= 0) return 0;
  if (precision == NERFCA_PREC_BF16) {
    // (a forward's size covers both the stash-writing and the stash-less call: a field of the fused kernels' shape whose latent table
    // is too large for their backward trains on the layer-wise path but is still evaluated by the fused forward)
    size_t n = 0;
    if (bf16_wide(field, nullptr, true)) n = wide_workspace_bytes(*field, n_points, backward);
    if (!bf16_wide(field, nullptr, backward != 0)) {
      const size_t m = tc_workspace_bytes(*field, n_points, backward);
      n = m > n ? m : n;
    }
    return n;
  }
  return simt_workspace_bytes(*field, n_points, backward);
}

extern "C" int nerfca_field_forward(const nerfca_field_t* field, const nerfca_samples_t* samples, int32_t precision,
                                    float* raw_out, void* stash, void* workspace, void* stream) {
  int rc = validate_field(field);
  if (rc) return rc;
  rc = validate_samples(samples, field->n_latent > 0);
  if (rc) return rc;
  NERFCA_REQUIRE(precision == NERFCA_PREC_FP32 || precision == NERFCA_PREC_BF16, NERFCA_E_ARG, "bad precision");
  if (samples->n_points == 0) return NERFCA_OK;
  NERFCA_REQUIRE(raw_out != nullptr, NERFCA_E_ARG, "raw_out is null");
  NERFCA_REQUIRE(workspace || nerfca_field_workspace_bytes(field, samples->n_points, precision, 0) == 0 ||
                     (stash && precision == NERFCA_PREC_FP32),
                 NERFCA_E_WORKSPACE, "workspace is null");
  if (precision == NERFCA_PREC_BF16) {
    if (bf16_wide(field, nullptr, stash != nullptr)) {
      rc = wide_supported(*field);
      if (rc) return rc;
      ProfScope prof(NERFCA_K_FIELD_FWD, (cudaStream_t)stream);
      return wide_field_forward(*field, *samples, raw_out, stash, workspace, (cudaStream_t)stream);
    }
    return tc_field_forward(*field, *samples, raw_out, stash, workspace, (cudaStream_t)stream);
  }
  ProfScope prof(NERFCA_K_FIELD_FWD, (cudaStream_t)stream);
  return simt_field_forward(*field, *samples, raw_out, stash, workspace, (cudaStream_t)stream);
}

extern "C" int nerfca_field_backward(const nerfca_field_t* field, const nerfca_samples_t* samples, int32_t precision,
                                     const float* d_raw, const void* stash, void* workspace,
                                     const nerfca_field_grads_t* grads, void* stream) {
  int rc = validate_field(field);
  if (rc) return rc;
  rc = validate_samples(samples, field->n_latent > 0);
  if (rc) return rc;
  NERFCA_REQUIRE(precision == NERFCA_PREC_FP32 || precision == NERFCA_PREC_BF16, NERFCA_E_ARG, "bad precision");
  if (samples->n_points == 0) return NERFCA_OK;
  NERFCA_REQUIRE(d_raw && stash && workspace && grads, NERFCA_E_ARG, "null pointer");
  for (int l = 0; l < field->n_hidden + 2; ++l) {
    NERFCA_REQUIRE(grads->weight[l] != nullptr, NERFCA_E_ARG, "weight gradient pointer missing");
    NERFCA_REQUIRE(!field->bias[l] || grads->bias[l], NERFCA_E_ARG, "bias gradient pointer missing");
  }
  if (precision == NERFCA_PREC_BF16) {
    if (bf16_wide(field, nullptr, true)) {
      rc = wide_supported(*field);
      if (rc) return rc;
      ProfScope prof(NERFCA_K_FIELD_BWD, (cudaStream_t)stream);
      return wide_field_backward(*field, *samples, d_raw, stash, workspace, *grads, (cudaStream_t)stream);
    }
    return tc_field_backward(*field, *samples, d_raw, stash, workspace, *grads, (cudaStream_t)stream);
  }
  ProfScope prof(NERFCA_K_FIELD_BWD, (cudaStream_t)stream);
  return simt_field_backward(*field, *samples, d_raw, stash, workspace, *grads, (cudaStream_t)stream);
}

extern "C" int nerfca_debug_x0(const nerfca_field_t* field, const nerfca_samples_t* samples, int32_t onehot, uint16_t* x0_out,
                               int32_t* kpad0_out, void* stream) {
  int rc = validate_field(field);
  if (rc) return rc;
  rc = validate_samples(samples, field->n_latent > 0);
  if (rc) return rc;
  rc = tc_supported(*field);
  if (rc) return rc;
  int kp = 0;
  rc = tc_debug_x0(*field, *samples, onehot, x0_out, &kp, (cudaStream_t)stream);
  if (kpad0_out) *kpad0_out = kp;
  return rc;
}

// ---- whole-step entry points ------------------------------------------------------------------------------------------
static int validate_step(const nerfca_step_t* s, bool training) {
  NERFCA_REQUIRE(s != nullptr, NERFCA_E_ARG, "step is null");
  int rc = validate_field(s->static_field);
  if (rc) return rc;
  if (s->dynamic_field) {
    rc = validate_field(s->dynamic_field);
    if (rc) return rc;
  }
  rc = validate_samples(s->samples, s->dynamic_field && s->dynamic_field->n_latent > 0);
  if (rc) return rc;
  NERFCA_REQUIRE(s->precision == NERFCA_PREC_FP32 || s->precision == NERFCA_PREC_BF16, NERFCA_E_ARG, "bad precision");
  if (training) {
    NERFCA_REQUIRE(!s->samples->points && s->samples->n_rays > 0, NERFCA_E_ARG, "a training step needs ray-generated samples");
    NERFCA_REQUIRE(s->static_grads && (!s->dynamic_field || s->dynamic_grads), NERFCA_E_ARG, "gradient descriptors missing");
  }
  if (s->precision == NERFCA_PREC_BF16 && bf16_wide(s->static_field, s->dynamic_field, training)) {
    rc = wide_supported(*s->static_field);
    if (rc) return rc;
    if (s->dynamic_field) rc = wide_supported(*s->dynamic_field);
  }
  return rc;
}

extern "C" size_t nerfca_step_stash_bytes(const nerfca_step_t* s) {
  if (!s || !s->static_field || !s->samples || s->samples->n_points <= 0) return 0;
  const long long P = s->samples->n_points;
  if (s->precision == NERFCA_PREC_BF16) {
    if (!bf16_wide(s->static_field, s->dynamic_field, true)) return tc_stash_bytes_n(s->dynamic_field ? 2 : 1, P);
    size_t n = up256(wide_stash_bytes(*s->static_field, P));
    if (s->dynamic_field) n += up256(wide_stash_bytes(*s->dynamic_field, P));
    return n;
  }
  size_t n = up256(simt_stash_bytes(*s->static_field, P));
  if (s->dynamic_field) n += up256(simt_stash_bytes(*s->dynamic_field, P));
  return n;
}

extern "C" size_t nerfca_step_workspace_bytes(const nerfca_step_t* s) {
  if (!s || !s->static_field || !s->samples || s->samples->n_points <= 0) return 0;
  const long long P = s->samples->n_points;
  if (s->precision == NERFCA_PREC_BF16 && !bf16_wide(s->static_field, s->dynamic_field, true)) {
    const nerfca_field_t* f[2] = {s->static_field, s->dynamic_field};
    return tc_workspace_bytes_n(f, s->dynamic_field ? 2 : 1, P, 1);
  }
  const bool wide = s->precision == NERFCA_PREC_BF16;
  // one buffer serves every fp32 call made with this descriptor: the stash-less forward of nerfca_fields_forward lays out
  // enc | ping | pong (in_dim + 2 hidden floats per sample of a chunk), the backward two hidden-wide gradient tiles
  size_t n = 0;
  const nerfca_field_t* fs[2] = {s->static_field, s->dynamic_field};
  for (int i = 0; i < 2; ++i) {
    if (!fs[i]) continue;
    for (int backward = 0; backward < 2; ++backward) {
      const size_t m = wide ? wide_workspace_bytes(*fs[i], P, backward) : simt_workspace_bytes(*fs[i], P, backward);
      n = m > n ? m : n;
    }
  }
  return n;
}

extern "C" int nerfca_fields_forward(const nerfca_field_t* fs, const nerfca_field_t* fd, const nerfca_samples_t* samples,
                                     int32_t precision, float* raw_s, float* raw_d, void* workspace, void* stream) {
  nerfca_step_t st = {};
  st.static_field = fs; st.dynamic_field = fd; st.samples = samples; st.precision = precision;
  int rc = validate_step(&st, false);
  if (rc) return rc;
  if (samples->n_points == 0) return NERFCA_OK;
  NERFCA_REQUIRE(raw_s && (!fd || raw_d), NERFCA_E_ARG, "output pointer is null");
  NERFCA_REQUIRE(workspace || nerfca_step_workspace_bytes(&st) == 0, NERFCA_E_WORKSPACE, "workspace is null");
  cudaStream_t cs = (cudaStream_t)stream;
  if (precision == NERFCA_PREC_BF16 && !bf16_wide(fs, fd, true)) {
    const nerfca_field_t* f[2] = {fs, fd};
    float* outs[2] = {raw_s, raw_d};
    return tc_fields_forward(f, fd ? 2 : 1, *samples, outs, nullptr, workspace, 1, nullptr, cs);
  }
  ProfScope prof(NERFCA_K_FIELD_FWD, cs);
  if (precision == NERFCA_PREC_BF16) {      // (the step descriptor sizes its workspace with the training criterion: same here)
    rc = wide_field_forward(*fs, *samples, raw_s, nullptr, workspace, cs);
    if (rc || !fd) return rc;
    return wide_field_forward(*fd, *samples, raw_d, nullptr, workspace, cs);
  }
  rc = simt_field_forward(*fs, *samples, raw_s, nullptr, workspace, cs);
  if (rc || !fd) return rc;
  return simt_field_forward(*fd, *samples, raw_d, nullptr, workspace, cs);
}

// ---- render: no-grad rays -> pixels with the line integral fused into the output-layer epilogue (tcgen05 path) --------------------
extern "C" size_t nerfca_render_workspace_bytes(const nerfca_field_t* fs, const nerfca_field_t* fd, const nerfca_samples_t* samples,
                                                int32_t precision) {
  if (!fs || !samples || samples->n_points <= 0 || precision != NERFCA_PREC_BF16 || bf16_wide(fs, fd, true)) return 0;
  const nerfca_field_t* f[2] = {fs, fd};
  return up256(tc_workspace_bytes_n(f, fd ? 2 : 1, samples->n_points, 0)) + 2 * up256((size_t)samples->n_rays * sizeof(float));
}

extern "C" int nerfca_render_rays(const nerfca_field_t* fs, const nerfca_field_t* fd, const nerfca_samples_t* samples, int32_t precision,
                                  const float* i0, int32_t activation, float* pix, float* pix_static, float* pix_dynamic, void* workspace,
                                  void* stream) {
  nerfca_step_t st = {};
  st.static_field = fs; st.dynamic_field = fd; st.samples = samples; st.precision = precision;
  int rc = validate_step(&st, false);
  if (rc) return rc;
  NERFCA_REQUIRE(precision == NERFCA_PREC_BF16 && !bf16_wide(fs, fd, true), NERFCA_E_UNSUPPORTED,
                 "nerfca_render_rays is the fused tcgen05 path (other precisions / shapes: nerfca_fields_forward + nerfca_integrate)");
  NERFCA_REQUIRE(!samples->points && samples->n_rays > 0, NERFCA_E_ARG, "rendering needs ray-generated samples");
  NERFCA_REQUIRE(i0 && pix && workspace, NERFCA_E_ARG, "null pointer");
  NERFCA_REQUIRE(!fd || (pix_static && pix_dynamic) || (!pix_static && !pix_dynamic), NERFCA_E_ARG, "give both component images or neither");
  cudaStream_t cs = (cudaStream_t)stream;
  const nerfca_field_t* f[2] = {fs, fd};
  const int n = fd ? 2 : 1;
  uint8_t* ws = (uint8_t*)workspace;
  const size_t pack_bytes = up256(tc_workspace_bytes_n(f, n, samples->n_points, 0));
  const size_t sum_bytes = up256((size_t)samples->n_rays * sizeof(float));
  float* sums[2] = {(float*)(ws + pack_bytes), (float*)(ws + pack_bytes + sum_bytes)};
  NERFCA_CUDA_OK(cudaMemsetAsync(sums[0], 0, 2 * sum_bytes, cs));
  rc = tc_fields_forward(f, n, *samples, nullptr, nullptr, workspace, 1, nullptr, cs, sums, activation);
  if (rc) return rc;
  return launch_render_finalize(sums[0], fd ? sums[1] : nullptr, i0, samples->n_rays, pix, pix_static, pix_dynamic, cs);
}

extern "C" int nerfca_train_step(const nerfca_step_t* s, void* stream) {
  int rc = validate_step(s, true);
  if (rc) return rc;
  const nerfca_samples_t& smp = *s->samples;
  if (smp.n_points == 0) return NERFCA_OK;
  const bool dyn = s->dynamic_field != nullptr;
  NERFCA_REQUIRE(s->raw_s && s->d_raw_s && (!dyn || (s->raw_d && s->d_raw_d)), NERFCA_E_ARG, "scratch pointer is null");
  NERFCA_REQUIRE(s->stash && s->workspace, NERFCA_E_WORKSPACE, "stash / workspace is null");
  NERFCA_REQUIRE(s->i0 && s->gt && s->wpix && s->loss && s->pix_out && s->terms_out, NERFCA_E_ARG, "null pointer");
  cudaStream_t cs = (cudaStream_t)stream;
  const nerfca_field_t* f[2] = {s->static_field, s->dynamic_field};
  const nerfca_field_grads_t* g[2] = {s->static_grads, s->dynamic_grads};
  float* raws[2] = {s->raw_s, s->raw_d};
  const float* draws[2] = {s->d_raw_s, s->d_raw_d};
  const int n = dyn ? 2 : 1;
  const long long P = smp.n_points;
  const bool fused = s->precision == NERFCA_PREC_BF16 && !bf16_wide(s->static_field, s->dynamic_field, true);
  const bool wide = s->precision == NERFCA_PREC_BF16 && !fused;
  uint8_t* stash_d = (uint8_t*)s->stash + (fused ? 0 : up256(wide ? wide_stash_bytes(*s->static_field, P) : simt_stash_bytes(*s->static_field, P)));
  const bool zero_terms = (s->flags & NERFCA_STEP_ZERO_TERMS) != 0;
  if (fused) {
    // (the forward launch also clears the loss sums when asked to: no separate memset node in the step's graph)
    rc = tc_fields_forward(f, n, smp, raws, s->stash, s->workspace, (s->flags & NERFCA_STEP_PACKED) ? 0 : 1, zero_terms ? s->terms_out : nullptr, cs);
  } else {
    if (zero_terms) NERFCA_CUDA_OK(cudaMemsetAsync(s->terms_out, 0, NERFCA_N_LOSS_TERMS * sizeof(double), cs));
    ProfScope prof(NERFCA_K_FIELD_FWD, cs);
    if (wide) {
      rc = wide_field_forward(*f[0], smp, raws[0], s->stash, s->workspace, cs);
      if (!rc && dyn) rc = wide_field_forward(*f[1], smp, raws[1], stash_d, s->workspace, cs);
    } else {
      rc = simt_field_forward(*f[0], smp, raws[0], s->stash, s->workspace, cs);
      if (!rc && dyn) rc = simt_field_forward(*f[1], smp, raws[1], stash_d, s->workspace, cs);
    }
  }
  if (rc) return rc;
  rc = nerfca_composite_loss(s->raw_s, dyn ? s->raw_d : nullptr, smp.depth, s->i0, s->gt, s->wpix, s->gw_stride, smp.n_rays, smp.n_depth,
                             s->activation, s->loss, s->pix_out, s->terms_out, s->d_raw_s, dyn ? s->d_raw_d : nullptr, stream);
  if (rc) return rc;
  if (fused) return tc_fields_backward(f, n, smp, draws, s->stash, s->workspace, 0, g, cs);
  ProfScope prof(NERFCA_K_FIELD_BWD, cs);
  if (wide) {
    rc = wide_field_backward(*f[0], smp, draws[0], s->stash, s->workspace, *g[0], cs);
    if (!rc && dyn) rc = wide_field_backward(*f[1], smp, draws[1], stash_d, s->workspace, *g[1], cs);
    return rc;
  }
  rc = simt_field_backward(*f[0], smp, draws[0], s->stash, s->workspace, *g[0], cs);
  if (!rc && dyn) rc = simt_field_backward(*f[1], smp, draws[1], stash_d, s->workspace, *g[1], cs);
  return rc;
}
