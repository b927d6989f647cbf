// Fused optimizer step: torch.optim.Adam + LinearLR of train/run_composite.py:209-215,305-308 over one
// flat fp32 buffer.  HBM-bound elementwise kernel: 4 reads + 3(4) writes of 4 B per parameter, float4-vectorised.
// The step counter is device-resident so the call is CUDA-graph capturable.
#include "adam.cuh"

namespace nerfca {

__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, const long long* __restrict__ step_dev, double lr, double b1, double b2, double eps,
                            double end_factor, long long decay, float grad_scale, int zero_grads) {
  __shared__ AdamScalars sa;
  if (threadIdx.x == 0) sa = adam_scalars(*step_dev, lr, b1, b2, eps, end_factor, decay, grad_scale);
  __syncthreads();
  const AdamScalars a = sa;
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    float4 P = *reinterpret_cast<float4*>(p + i4), G = *reinterpret_cast<float4*>(g + i4);
    float4 M = *reinterpret_cast<float4*>(m + i4), V = *reinterpret_cast<float4*>(v + i4);
    adam_one(P.x, G.x, M.x, V.x, a); adam_one(P.y, G.y, M.y, V.y, a);
    adam_one(P.z, G.z, M.z, V.z, a); adam_one(P.w, G.w, M.w, V.w, a);
    *reinterpret_cast<float4*>(p + i4) = P; *reinterpret_cast<float4*>(m + i4) = M; *reinterpret_cast<float4*>(v + i4) = V;
    if (zero_grads) *reinterpret_cast<float4*>(g + i4) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (long long i = i4; i < n; ++i) {
      adam_one(p[i], g[i], m[i], v[i], a);
      if (zero_grads) g[i] = 0.f;
    }
  }
}

// after every block has read the old counter (stream order: separate tiny launch)
__global__ void adam_bump_kernel(long long* step_dev) { *step_dev += 1; }

}  // namespace nerfca

using namespace nerfca;

extern "C" int nerfca_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t* step_dev,
                                const nerfca_adam_cfg_t* cfg, float grad_scale, int32_t zero_grads, void* stream) {
  NERFCA_REQUIRE(params && grads && exp_avg && exp_avg_sq && step_dev && cfg, NERFCA_E_ARG, "null pointer");
  NERFCA_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, NERFCA_E_ARG,
                 "buffers must be 16-byte aligned");
  if (n <= 0) return NERFCA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(NERFCA_K_ADAM, st);
  adam_kernel<<<div_up((n + 3) / 4, 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, (long long)n, (const long long*)step_dev,
                                                      cfg->lr, cfg->beta1, cfg->beta2, cfg->eps, cfg->lr_end_factor,
                                                      (long long)cfg->lr_decay_steps, grad_scale, zero_grads);
  NERFCA_LAUNCH_OK();
  adam_bump_kernel<<<1, 1, 0, st>>>((long long*)step_dev);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}
