// C-ABI plumbing: error state, argument validation and precision dispatch of the field entry points.
#include "common.cuh"

namespace nerfca {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

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
// bf16 tcgen05 path (mlp_tc.cu)
int tc_supported(const nerfca_field_t& f);
size_t tc_stash_bytes(const nerfca_field_t& f, long long P);
size_t tc_workspace_bytes(const nerfca_field_t& f, long long P, int backward);
int tc_field_forward(const nerfca_field_t& f, const nerfca_samples_t& s, float* raw_out, void* stash, void* workspace,
                     cudaStream_t st);
int tc_field_backward(const nerfca_field_t& f, const nerfca_samples_t& s, const float* d_raw, const void* stash,
                      void* workspace, const nerfca_field_grads_t& gr, cudaStream_t st);

}  // namespace nerfca

using namespace nerfca;

extern "C" const char* nerfca_last_error(void) { return g_last_error.c_str(); }
extern "C" int nerfca_abi_version(void) { return NERFCA_ABI_VERSION; }

extern "C" size_t nerfca_field_stash_bytes(const nerfca_field_t* field, int64_t n_points, int32_t precision) {
  if (!field || n_points <= 0) return 0;
  return precision == NERFCA_PREC_BF16 ? tc_stash_bytes(*field, n_points) : simt_stash_bytes(*field, n_points);
}

extern "C" size_t nerfca_field_workspace_bytes(const nerfca_field_t* field, int64_t n_points, int32_t precision,
                                               int32_t backward) {
  if (!field || n_points <= 0) return 0;
  return precision == NERFCA_PREC_BF16 ? tc_workspace_bytes(*field, n_points, backward)
                                       : simt_workspace_bytes(*field, n_points, backward);
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
    rc = tc_supported(*field);
    if (rc) return rc;
    return tc_field_forward(*field, *samples, raw_out, stash, workspace, (cudaStream_t)stream);
  }
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
    rc = tc_supported(*field);
    if (rc) return rc;
    return tc_field_backward(*field, *samples, d_raw, stash, workspace, *grads, (cudaStream_t)stream);
  }
  return simt_field_backward(*field, *samples, d_raw, stash, workspace, *grads, (cudaStream_t)stream);
}
