// Microbenchmark: cost of ISSUING tcgen05.mma from one thread (lane-0 branch vs elect.sync), SS vs TS operand form, and the
// round-trip latency of tcgen05.ld / tcgen05.st.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue mma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../nerf-ca_b200/csrc/tc_common.cuh"
using namespace nerfca::tc;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int TS>
__device__ __forceinline__ void issue64(uint32_t tmem, Desc a, Desc b, uint32_t idesc, int n_mma) {
  for (int i = 0; i < n_mma; i += 8) {
    if (TS) umma_ts_k<8, KSTEP_KMAJOR>(tmem + (i & 8 ? 128 : 0), tmem + 256, b, idesc, 0);
    else umma_k<8, KSTEP_KMAJOR, KSTEP_KMAJOR>(tmem + (i & 8 ? 128 : 0), a, b, idesc, 0);
  }
}

// style 0: if (threadIdx.x == 0)   1: warp 0, elect.sync
__global__ void __launch_bounds__(128, 1) mma_issue(int style, int ts, int n_mma, int N, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); mbar_init_fence(); }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_ptr), 512);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  const Desc a = kmajor(smem_u32(smem)), b = kmajor(smem_u32(smem + 32768));
  const uint32_t idesc = instr_desc(128, N, 0, 0);
  long long t0 = 0, t1 = 0, t2 = 0;
  if (style == 0) {
    if (threadIdx.x == 0) {
      t0 = clock64();
      if (ts) issue64<1>(tmem, a, b, idesc, n_mma); else issue64<0>(tmem, a, b, idesc, n_mma);
      t1 = clock64();
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
    }
  } else {
    if (warp == 0) {
      if (elect_one()) {
        t0 = clock64();
        if (ts) issue64<1>(tmem, a, b, idesc, n_mma); else issue64<0>(tmem, a, b, idesc, n_mma);
        t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  tc_fence_after();
  // TMEM round trips, one warp at a time would hide nothing: all 4 warps do the same thing on their own lanes
  const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16);
  uint32_t acc = 0;
  long long c0 = clock64();
  for (int r = 0; r < 16; ++r) { uint32_t v[16]; tmem_ld16(ta + (r & 3) * 16, v); tmem_ld_wait(); for (int j = 0; j < 16; ++j) acc += v[j]; }
  long long c1 = clock64();
  for (int r = 0; r < 16; ++r) { uint32_t v[32]; tmem_ld32(ta + (r & 3) * 32, v); tmem_ld_wait(); for (int j = 0; j < 32; ++j) acc += v[j]; }
  long long c2 = clock64();
  for (int r = 0; r < 16; ++r) { uint32_t v[32], w[32]; tmem_ld32(ta, v); tmem_ld32(ta + 32, w); tmem_ld_wait(); for (int j = 0; j < 32; ++j) acc += v[j] ^ w[j]; }
  long long c3 = clock64();
  for (int r = 0; r < 16; ++r) { uint32_t v[32]; for (int j = 0; j < 32; ++j) v[j] = acc + j; tmem_st32(ta + 256, v); tmem_st_wait(); }
  long long c4 = clock64();
  for (int r = 0; r < 16; ++r) { uint32_t v[2]; tmem_ld2(ta + r, v); tmem_ld_wait(); acc += v[0] + v[1]; }
  long long c5 = clock64();
  if (threadIdx.x == 0) { out[2] = (c1 - c0) / 16; out[3] = (c2 - c1) / 16; out[4] = (c3 - c2) / 16; out[5] = (c4 - c3) / 16; out[6] = (c5 - c4) / 16; out[7] = acc; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  for (int N : {128, 64, 16}) for (int ts = 0; ts < 2; ++ts) for (int style = 0; style < 2; ++style) {
    for (int rep = 0; rep < 2; ++rep) {
      mma_issue<<<1, 128, 65536 + 1024>>>(style, ts, 64, N, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("style %d: %s\n", style, cudaGetErrorString(e)); return 1; }
    }
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("N=%3d %s %-10s: issue %5.1f cyc/MMA, issue+exec %5.1f cyc/MMA | round trips (4 warps): ld.x16 %lld, ld.x32 %lld, 2 x ld.x32 %lld, st.x32 %lld, ld.x2 %lld cyc\n", N,
           ts ? "TS" : "SS", style ? "elect" : "tid==0", h[0] / 64.0, h[1] / 64.0, h[2], h[3], h[4], h[5], h[6]);
  }
  return 0;
}
