// Cone-beam ray generation (A2), depth jitter (A3), sample-point materialisation (A4) and the
// stand-alone positional-encoding entry (A5).  All HBM-bound elementwise kernels: one thread per output
// element, consecutive threads write consecutive addresses.
#include "common.cuh"

namespace nerfca {

struct Pose3x4 { float r[3][3]; float t[3]; };

// train/proj_helpers.py:73-84.  One thread per detector pixel (i = u index, j = v index, ray id i*H + j).
// Every op is individually rounded (no FMA contraction) so directions are bit-identical to torch's fp32 ops.
__global__ void gen_rays_kernel(Pose3x4 pose, int W, int H, float du, float dv, float ou, float ov, float dsd,
                                float half_w, float half_h, float* __restrict__ origins, float* __restrict__ dirs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W * H) return;
  const int i = idx / H, j = idx - i * H;
  const float u = __fadd_rn(__fmul_rn(__fsub_rn(__fadd_rn((float)i, 0.5f), half_w), du), ou);
  const float v = __fadd_rn(__fmul_rn(__fsub_rn(__fadd_rn((float)j, 0.5f), half_h), dv), ov);
  const float dx = __fdiv_rn(u, dsd), dy = __fdiv_rn(v, dsd);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float a = __fmul_rn(pose.r[k][0], dx), b = __fmul_rn(pose.r[k][1], dy);
    dirs[3 * (size_t)idx + k] = __fadd_rn(__fadd_rn(a, b), __fmul_rn(pose.r[k][2], 1.0f));
    origins[3 * (size_t)idx + k] = pose.t[k];
  }
}

// train/model_helpers.py:3-12
__global__ void jitter_depth_kernel(const float* __restrict__ z, const float* __restrict__ t, int n, float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float zk = z[k];
  const float lower = (k == 0) ? zk : __fmul_rn(0.5f, __fadd_rn(zk, z[k - 1]));
  const float upper = (k == n - 1) ? zk : __fmul_rn(0.5f, __fadd_rn(z[k + 1], zk));
  out[k] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t[k]));
}

__global__ void sample_points_kernel(SampleSrc src, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= src.n_points) return;
  float x, y, z;
  load_point(src, idx, x, y, z);
  out[3 * idx] = x; out[3 * idx + 1] = y; out[3 * idx + 2] = z;
}

// One thread per (sample, feature): fully coalesced stores of the [P, in_dim] first-layer input.
__global__ void encode_kernel(SampleSrc src, EncDesc enc, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = src.n_points * enc.in_dim;
  if (idx >= total) return;
  const long long p = idx / enc.in_dim;
  const int f = (int)(idx - p * enc.in_dim);
  float x, y, z;
  load_point(src, p, x, y, z);
  const int phase = (f >= enc.enc_dim) ? load_phase(src, p) : 0;
  out[idx] = enc_feature(enc, f, x, y, z, phase);
}

int launch_encode(const nerfca_field_t& field, const nerfca_samples_t& samples, long long p0, long long np, float* out,
                  cudaStream_t st) {
  const SampleSrc src = make_src(samples, p0, np);
  const EncDesc enc = make_enc(field);
  const long long total = np * enc.in_dim;
  if (total == 0) return NERFCA_OK;
  encode_kernel<<<div_up(total, 256), 256, 0, st>>>(src, enc, out);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

}  // namespace nerfca

using namespace nerfca;

extern "C" int nerfca_gen_rays(const float* pose_host, int32_t width, int32_t height, float du, float dv, float off_u,
                               float off_v, float dsd, float* origins, float* dirs, void* stream) {
  NERFCA_REQUIRE(pose_host && origins && dirs, NERFCA_E_ARG, "null pointer");
  NERFCA_REQUIRE(width > 0 && height > 0, NERFCA_E_ARG, "empty detector");
  Pose3x4 p;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) p.r[r][c] = pose_host[4 * r + c];
    p.t[r] = pose_host[4 * r + 3];
  }
  // `img_width / 2` is a python float that torch rounds to fp32 when it meets the fp32 tensor (proj_helpers.py:79)
  const float half_w = (float)((double)width / 2.0), half_h = (float)((double)height / 2.0);
  const int n = width * height;
  gen_rays_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(p, width, height, du, dv, off_u, off_v, dsd, half_w,
                                                                    half_h, origins, dirs);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

extern "C" int nerfca_jitter_depth(const float* z, const float* t_rand, int32_t n, float* out, void* stream) {
  NERFCA_REQUIRE(z && t_rand && out, NERFCA_E_ARG, "null pointer");
  if (n <= 0) return NERFCA_OK;
  jitter_depth_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(z, t_rand, n, out);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

// one thread per (ray, 16-byte half-row pair): a ray-table row is 12 doubles = 96 bytes = 6 x 16 B
__global__ void gather_batch_kernel(const double* __restrict__ table, const long long* __restrict__ phases, long long n_table,
                                    const long long* __restrict__ ids, int n_batch, double* __restrict__ rays_out,
                                    int* __restrict__ phases_out, int* __restrict__ err_flag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = t / 6, part = t - r * 6;
  if (r >= n_batch) return;
  const long long id = __ldg(ids + r);
  if (id < 0 || id >= n_table) {
    if (err_flag && part == 0) *err_flag = 1;
    return;
  }
  const double2 v = __ldg(reinterpret_cast<const double2*>(table + id * 12) + part);
  reinterpret_cast<double2*>(rays_out + (long long)r * 12)[part] = v;
  if (part == 0 && phases_out) phases_out[r] = phases ? (int)__ldg(phases + id) : 0;
}

extern "C" int nerfca_gather_batch(const double* rays_table, const int64_t* phases_table, int64_t n_table, const int64_t* ids,
                                   int32_t n_batch, double* rays_out, int32_t* phases_out, int32_t* err_flag, void* stream) {
  NERFCA_REQUIRE(n_table > 0 && n_batch >= 0, NERFCA_E_ARG, "empty ray table or negative batch size");
  if (n_batch == 0) return NERFCA_OK;
  NERFCA_REQUIRE(rays_table && ids && rays_out, NERFCA_E_ARG, "null pointer");
  NERFCA_REQUIRE(((uintptr_t)rays_table & 15) == 0 && ((uintptr_t)rays_out & 15) == 0, NERFCA_E_ARG, "ray rows must be 16-byte aligned");
  gather_batch_kernel<<<div_up((long long)n_batch * 6, 256), 256, 0, (cudaStream_t)stream>>>(
      rays_table, (const long long*)phases_table, (long long)n_table, (const long long*)ids, n_batch, rays_out, phases_out, err_flag);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

extern "C" int nerfca_sample_points(const nerfca_samples_t* samples, float* points_out, void* stream) {
  NERFCA_REQUIRE(samples && points_out, NERFCA_E_ARG, "null pointer");
  int rc = validate_samples(samples, false);
  if (rc) return rc;
  if (samples->n_points == 0) return NERFCA_OK;
  sample_points_kernel<<<div_up(samples->n_points, 256), 256, 0, (cudaStream_t)stream>>>(make_src(*samples), points_out);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

extern "C" int nerfca_encode(const nerfca_field_t* field, const nerfca_samples_t* samples, float* enc_out, void* stream) {
  NERFCA_REQUIRE(field && samples && enc_out, NERFCA_E_ARG, "null pointer");
  int rc = validate_field(field);
  if (rc) return rc;
  rc = validate_samples(samples, field->n_latent > 0);
  if (rc) return rc;
  if (samples->n_points == 0) return NERFCA_OK;
  return launch_encode(*field, *samples, 0, samples->n_points, enc_out, (cudaStream_t)stream);
}
