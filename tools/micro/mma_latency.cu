// Microbenchmark: latency from issuing K tcgen05.mma (N = 128) + tcgen05.commit to (a) the issuing thread and (b) another warp
// observing the mbarrier phase, with mbarrier.try_wait (suspending) vs mbarrier.test_wait (pure spin).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../nerf-ca_b200/csrc/tc_common.cuh"
using namespace nerfca::tc;

__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(256, 1) mma_latency(int ts, int k_mma, int spin, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar_go;
  __shared__ uint32_t tmem_ptr;
  __shared__ long long t_issue[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar_go), 1); mbar_init_fence(); }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_ptr), 512);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  const uint32_t b = smem_u32(&bar), bgo = smem_u32(&bar_go);
  long long sum_self = 0, sum_other = 0;
  for (int r = 0; r < reps; ++r) {
    const uint32_t ph = r & 1;
    if (warp == 0) {
      if (elect_one()) {
        const Desc da = kmajor(smem_u32(smem)), db = kmajor(smem_u32(smem + 32768));
        const uint32_t idesc = instr_desc(128, 128, 0, 0);
        // let the waiter settle into its wait first
        long long t = clock64(); while (clock64() - t < 3000) {}
        const long long t0 = clock64();
        t_issue[r] = t0;
        mbar_arrive(bgo);
        for (int i = 0; i < k_mma; ++i) {
          if (ts) umma_ts(tmem, tmem + 256, db.lo + (i & 7) * KSTEP_KMAJOR, db.hi, idesc, i > 0);
          else umma_lh(tmem, da.lo + (i & 7) * KSTEP_KMAJOR, da.hi, db.lo + (i & 7) * KSTEP_KMAJOR, db.hi, idesc, i > 0);
        }
        umma_commit(b);
        if (spin) { while (!mbar_test_wait(b, ph)) {} } else { while (!mbar_try_wait(b, ph)) {} }
        sum_self += clock64() - t0;
      }
      __syncwarp();
    } else if (warp == 1) {
      if (lane == 0) {
        while (!mbar_try_wait(bgo, ph)) {}
        if (spin) { while (!mbar_test_wait(b, ph)) {} } else { while (!mbar_try_wait(b, ph)) {} }
        const long long t1 = clock64();
        sum_other += t1 - t_issue[r];
      }
      __syncwarp();
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sum_self / reps;
  if (threadIdx.x == 32) out[1] = sum_other / reps;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_latency, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  for (int ts = 0; ts < 2; ++ts) for (int spin = 0; spin < 2; ++spin) for (int k : {1, 8, 16}) {
    cudaMemset(d, 0, 64);
    mma_latency<<<1, 256, 65536 + 1024>>>(ts, k, spin, 32, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%s %-9s K=%2d MMAs: issuing thread sees completion after %5lld cyc, another warp after %5lld cyc (floor %d)\n", ts ? "TS" : "SS", spin ? "test_wait" : "try_wait", k, h[0], h[1], 64 * k);
  }
  return 0;
}
