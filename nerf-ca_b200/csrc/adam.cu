// Fused optimizer step: torch.optim.Adam + LinearLR of train/run_composite.py:209-215,305-308 over one flat
// fp32 buffer, the clearing of the gradient buffer for the next step's accumulation, and the bf16 re-pack of the updated
// parameters into the operand tiles of the tcgen05 kernels -- one launch.  HBM-bound elementwise kernel: 4 reads + 4 writes
// of 4 B per parameter, float4-vectorised.
#include "adam.cuh"

namespace nerfca {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, AdamScalars a, int zero_grads,
                                                   const __grid_constant__ RepackTable rt) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  if (i4 + 3 < n) {
    float4 P = *reinterpret_cast<float4*>(p + i4), G = *reinterpret_cast<float4*>(g + i4);
    float4 M = *reinterpret_cast<float4*>(m + i4), V = *reinterpret_cast<float4*>(v + i4);
    adam_one(P.x, G.x, M.x, V.x, a); adam_one(P.y, G.y, M.y, V.y, a);
    adam_one(P.z, G.z, M.z, V.z, a); adam_one(P.w, G.w, M.w, V.w, a);
    *reinterpret_cast<float4*>(p + i4) = P; *reinterpret_cast<float4*>(m + i4) = M; *reinterpret_cast<float4*>(v + i4) = V;
    if (zero_grads) *reinterpret_cast<float4*>(g + i4) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rt.n_segs > 0) {
      const float pv[4] = {P.x, P.y, P.z, P.w};
      repack_group(rt, i4, pv);
    }
  } else {
    float pv[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long i = i4; i < n; ++i) {
      adam_one(p[i], g[i], m[i], v[i], a);
      if (zero_grads) g[i] = 0.f;
      pv[i - i4] = p[i];
    }
    if (rt.n_segs > 0) repack_group(rt, i4, pv);
  }
}

int adam_repack_table(const nerfca_repack_t* repack, const float* params, long long n, RepackTable* rt) {
  rt->n_segs = 0;
  if (!repack) return NERFCA_OK;
  NERFCA_REQUIRE(repack->static_field && repack->workspace, NERFCA_E_ARG, "repack: field / workspace missing");
  const nerfca_field_t* f[2] = {repack->static_field, repack->dynamic_field};
  return make_repack_table(f, repack->dynamic_field ? 2 : 1, repack->workspace, params, n, rt);
}

}  // namespace nerfca

using namespace nerfca;

extern "C" int nerfca_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                const nerfca_adam_step_t* cfg, float grad_scale, int32_t zero_grads, const nerfca_repack_t* repack,
                                void* stream) {
  NERFCA_REQUIRE(params && grads && exp_avg && exp_avg_sq && cfg, NERFCA_E_ARG, "null pointer");
  NERFCA_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, NERFCA_E_ARG,
                 "buffers must be 16-byte aligned");
  NERFCA_REQUIRE(cfg->bias_correction1 > 0.0 && cfg->bias_correction2_sqrt > 0.0, NERFCA_E_ARG, "bias corrections must be positive");
  if (n <= 0) return NERFCA_OK;
  RepackTable rt;
  int rc = adam_repack_table(repack, params, (long long)n, &rt);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(NERFCA_K_ADAM, st);
  adam_kernel<<<div_up((n + 3) / 4, 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, (long long)n, adam_scalars(*cfg, grad_scale),
                                                      zero_grads, rt);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}
