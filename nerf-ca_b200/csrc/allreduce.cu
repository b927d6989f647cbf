// the step's only exchange -- the sum of the flat gradient buffers over the ranks -- fused with the optimizer step,
// over NVLink / NVSwitch peer memory.  Every rank's gradient buffer and a small signal pad are peer-mapped (the host side uses
// torch's symmetric memory for the mapping only).  One step, epoch e = 1, 2, ...:
//
//   allreduce_adam_kernel   block 0 raises   ready[rank] = e   in every peer's pad (the gradients were completed by the backward
//                           kernel before this launch); every block waits until its own pad shows ready[r] >= e for all r; each
//                           thread then loads its 4 elements from all `world` buffers in rank order (the same order on every rank,
//                           so the replicas stay bit-identical), applies Adam + LinearLR to the local parameters; the last block
//                           to finish raises   done[rank] = e   in every peer's pad.
//   grad_reset_kernel       waits until its own pad shows done[r] >= e for all r (nobody reads this rank's gradients any more)
//                           and clears them for the next step's accumulation.
//
// Compared with ncclAllReduce + the optimizer kernel this is one 612 KB read per peer straight into the update (no reduced copy is
// ever written), two flag round trips and no separate collective launch.  All waits are bounded and trap instead of hanging.
#include "adam.cuh"

namespace nerfca {

constexpr int PAD_READY = 0, PAD_DONE = 64, PAD_COUNTER = 128;   // uint32 slots of a signal pad (world_size <= 64)

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {   // peer memory: never through a possibly stale L1 line
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_flags(const uint32_t* flags, int world, uint32_t epoch) {
  if ((int)threadIdx.x < world) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
      __nanosleep(100);
      if (clock64() - t0 > 20000000000LL) {              // ~10 s: a peer never arrived
#ifdef NERFCA_TIMELINE_BUILD
        printf("peer flag timeout: block %d waits for rank %d, sees %u, epoch %u\n", (int)blockIdx.x, (int)threadIdx.x, ld_acquire_sys(flags + threadIdx.x), epoch);
#endif
        __trap();
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) allreduce_adam_kernel(const float* const* __restrict__ peer_grads, uint32_t* const* __restrict__ peer_pads,
                                                             int rank, int world, uint32_t epoch, float* __restrict__ p,
                                                             float* __restrict__ m, float* __restrict__ v, long long n, AdamScalars a,
                                                             long long terms_off, double* __restrict__ terms_out,
                                                             const __grid_constant__ RepackTable rt) {
  __shared__ int s_last;
  uint32_t* my_pad = peer_pads[rank];
  if (blockIdx.x == 0 && (int)threadIdx.x < world) st_release_sys(peer_pads[threadIdx.x] + PAD_READY + rank, epoch);
  wait_flags(my_pad + PAD_READY, world, epoch);
  // the step's 16 loss sums ride in the same peer-mapped buffer behind the gradients: summed in rank order
  // (entries 2, 3 are maxima), so every rank holds the same global terms without a second collective
  if (blockIdx.x == 0 && terms_out && threadIdx.x < NERFCA_N_LOSS_TERMS) {
    double acc = 0.0;
    for (int r = 0; r < world; ++r) {
      const double t = *reinterpret_cast<const volatile double*>(peer_grads[r] + terms_off + 2 * threadIdx.x);
      acc = (threadIdx.x == NERFCA_T_SIGMA_S_MAX || threadIdx.x == NERFCA_T_SIGMA_D_MAX) ? fmax(acc, t) : acc + t;
    }
    terms_out[threadIdx.x] = acc;
  }
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 g = ld_peer_f4(peer_grads[r] + i4);
      G.x += g.x; G.y += g.y; G.z += g.z; G.w += g.w;
    }
    float4 P = *reinterpret_cast<float4*>(p + i4), M = *reinterpret_cast<float4*>(m + i4), V = *reinterpret_cast<float4*>(v + i4);
    adam_one(P.x, G.x, M.x, V.x, a); adam_one(P.y, G.y, M.y, V.y, a);
    adam_one(P.z, G.z, M.z, V.z, a); adam_one(P.w, G.w, M.w, V.w, a);
    *reinterpret_cast<float4*>(p + i4) = P; *reinterpret_cast<float4*>(m + i4) = M; *reinterpret_cast<float4*>(v + i4) = V;
    if (rt.n_segs > 0) {
      const float pv[4] = {P.x, P.y, P.z, P.w};
      repack_group(rt, i4, pv);
    }
  } else if (i4 < n) {
    float pv[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long i = i4; i < n; ++i) {
      float g = 0.f;
      for (int r = 0; r < world; ++r) g += *reinterpret_cast<const volatile float*>(peer_grads[r] + i);
      adam_one(p[i], g, m[i], v[i], a);
      pv[i - i4] = p[i];
    }
    if (rt.n_segs > 0) repack_group(rt, i4, pv);
  }
  // the last block to get here tells every peer that this rank has finished reading
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(my_pad + PAD_COUNTER, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
    if (s_last) my_pad[PAD_COUNTER] = 0u;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < world) st_release_sys(peer_pads[threadIdx.x] + PAD_DONE + rank, epoch);
}

__global__ void __launch_bounds__(256) grad_reset_kernel(float* __restrict__ g, long long n, const uint32_t* __restrict__ done_flags, int world,
                                                         uint32_t epoch) {
  wait_flags(done_flags, world, epoch);
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) *reinterpret_cast<float4*>(g + i4) = make_float4(0.f, 0.f, 0.f, 0.f);
  else for (long long i = i4; i < n; ++i) g[i] = 0.f;
}

int adam_repack_table(const nerfca_repack_t* repack, const float* params, long long n, RepackTable* rt);   // adam.cu

}  // namespace nerfca

using namespace nerfca;

extern "C" int nerfca_allreduce_adam_step(const nerfca_peers_t* peers, uint32_t epoch, float* params, float* grads, float* exp_avg,
                                          float* exp_avg_sq, int64_t n, const nerfca_adam_step_t* cfg, const nerfca_repack_t* repack,
                                          int64_t terms_offset, double* terms_out, void* stream) {
  NERFCA_REQUIRE(peers && params && grads && exp_avg && exp_avg_sq && cfg, NERFCA_E_ARG, "null pointer");
  NERFCA_REQUIRE(peers->grads && peers->signals, NERFCA_E_ARG, "null peer tables");
  NERFCA_REQUIRE(peers->world_size >= 1 && peers->world_size <= 64 && peers->rank >= 0 && peers->rank < peers->world_size, NERFCA_E_ARG,
                 "rank / world size out of range (1 <= world <= 64)");
  NERFCA_REQUIRE(epoch != 0, NERFCA_E_ARG, "epochs start at 1 (the signal pads start zeroed)");
  NERFCA_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, NERFCA_E_ARG,
                 "buffers must be 16-byte aligned");
  NERFCA_REQUIRE(!terms_out || (terms_offset >= n && terms_offset % 2 == 0), NERFCA_E_ARG,
                 "the loss sums must sit behind the gradients in the peer-mapped buffer, 8-byte aligned");
  if (n <= 0) return NERFCA_OK;
  RepackTable rt;
  int rc = adam_repack_table(repack, params, (long long)n, &rt);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(NERFCA_K_ADAM, st);
  const unsigned grid = div_up((n + 3) / 4, 256);
  allreduce_adam_kernel<<<grid, 256, 0, st>>>(peers->grads, peers->signals, peers->rank, peers->world_size, epoch, params, exp_avg, exp_avg_sq,
                                             (long long)n, adam_scalars(*cfg, 1.f), (long long)terms_offset, terms_out, rt);
  NERFCA_LAUNCH_OK();
  // the pad of this rank: only the host knows its address through the peer table; it is passed separately to keep the kernel simple
  grad_reset_kernel<<<grid, 256, 0, st>>>(grads, (long long)n, peers->own_signals + PAD_DONE, peers->world_size, epoch);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}
