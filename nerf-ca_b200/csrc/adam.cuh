// Adam + LinearLR scalar set-up and the per-element update, shared by the single-GPU optimizer kernel (adam.cu) and the
// all-reduce-fused multi-GPU one (allreduce.cu).
//
// Bit-for-bit torch.optim.Adam(foreach=True) (torch/optim/adam.py::_multi_tensor_adam, weight_decay 0, amsgrad off) for fp32
// parameters: the per-step scalars (learning rate of the LinearLR recursion, 1 - beta1^t, sqrt(1 - beta2^t)) are computed by the
// HOST in python double arithmetic exactly as torch computes them and arrive by value; the kernel applies torch's foreach kernels'
// operation order and roundings:
//   _foreach_lerp_(m, g, 1 - b1)          m = fma(w1, g - m, m)                      (ATen Lerp.h, weight < 0.5 branch)
//   _foreach_mul_(v, b2)                  v = v * b2
//   _foreach_addcmul_(v, g, g, 1 - b2)    v = fma(w2, g * g, v)                      (a + scalar * (b * c), ForeachFunctors.cuh)
//   d = _foreach_sqrt(v); d /= bc2_sqrt; d += eps
//   _foreach_addcdiv_(p, m, d, -lr/bc1)   p = fma(step, m / d, p)                    (a + scalar * (b / c))
#pragma once
#include "common.cuh"
#include "pack.cuh"

namespace nerfca {

struct AdamScalars { float w1, beta2, w2, step_size_neg, bc2_sqrt, eps, grad_scale; };

inline AdamScalars adam_scalars(const nerfca_adam_step_t& c, float grad_scale) {
  AdamScalars sa;
  sa.w1 = (float)(1.0 - c.beta1);
  sa.beta2 = (float)c.beta2;
  sa.w2 = (float)(1.0 - c.beta2);
  sa.step_size_neg = (float)((c.lr / c.bias_correction1) * -1.0);
  sa.bc2_sqrt = (float)c.bias_correction2_sqrt;
  sa.eps = (float)c.eps;
  sa.grad_scale = grad_scale;
  return sa;
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& a) {
  const float gs = (a.grad_scale == 1.f) ? g : __fmul_rn(g, a.grad_scale);
  m = __fmaf_rn(a.w1, __fsub_rn(gs, m), m);
  v = __fmaf_rn(a.w2, __fmul_rn(gs, gs), __fmul_rn(v, a.beta2));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), a.bc2_sqrt), a.eps);
  p = __fmaf_rn(a.step_size_neg, __fdiv_rn(m, denom), p);
}

}  // namespace nerfca
