// Adam + LinearLR scalar set-up and the per-element update, shared by the single-GPU optimizer kernel (adam.cu) and the
// all-reduce-fused multi-GPU one (allreduce.cu).
#pragma once
#include "common.cuh"

namespace nerfca {

struct AdamScalars { float w1, beta2, w2, step_size, bc2_sqrt, eps, grad_scale; };

// Same operation order / roundings as torch's foreach Adam kernels: lerp_(g, 1-b1) -> fma(w, g - m, m);
// v.mul_(b2).addcmul_(g, g, 1-b2) -> fma(w2 * g, g, b2 * v);  denom = sqrt(v) / bc2_sqrt + eps;  p += -step * (m / denom).
__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, const AdamScalars& a) {
  const float gs = g * a.grad_scale;
  m = fmaf(a.w1, gs - m, m);
  v = fmaf(__fmul_rn(a.w2, gs), gs, __fmul_rn(v, a.beta2));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), a.bc2_sqrt), a.eps);
  p = fmaf(-a.step_size, __fdiv_rn(m, denom), p);
}

// scalars of update t = *step_dev + 1 (LinearLR with start_factor 1, torch bias corrections)
__device__ __forceinline__ AdamScalars adam_scalars(long long step_now, double lr, double b1, double b2, double eps, double end_factor,
                                                    long long decay, float grad_scale) {
  AdamScalars sa;
  const long long t = step_now + 1;
  const double frac = decay > 0 ? (double)((t - 1 < decay) ? t - 1 : decay) / (double)decay : 0.0;
  const double lr_t = lr * (1.0 + (end_factor - 1.0) * frac);
  const double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
  sa.w1 = (float)(1.0 - b1); sa.beta2 = (float)b2; sa.w2 = (float)(1.0 - b2);
  sa.step_size = (float)(lr_t / bc1); sa.bc2_sqrt = (float)sqrt(bc2); sa.eps = (float)eps; sa.grad_scale = grad_scale;
  return sa;
}

}  // namespace nerfca
