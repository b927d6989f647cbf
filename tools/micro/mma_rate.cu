// Microbenchmark: tcgen05.mma issue/execute rate for different shared-memory operand layouts, and TMEM load rate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../nerf-ca_b200/csrc/tc_common.cuh"
using namespace nerfca::tc;

__device__ __forceinline__ uint64_t desc_gen(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}

// mode 0: no-swizzle K-major A,B   1: SW128 K-major A,B   2: no-swizzle MN-major A,B   3: SW128 MN-major A,B
// 4: no-swz A K-major, B MN-major (dgrad)   5: SW128 same
__global__ void __launch_bounds__(128, 1) mma_rate(int mode, int n_mma, int N, int n_acc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
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
  const uint32_t A = smem_u32(smem), B = smem_u32(smem + 32768);
  if (threadIdx.x == 0) {
    const bool a_mn = (mode == 2 || mode == 3), b_mn = (mode >= 2);
    const bool swz = (mode & 1);
    const uint32_t idesc = instr_desc(128, N, a_mn ? 1 : 0, b_mn ? 1 : 0);
    // lean issue: descriptors built once, the K-step advance is an immediate add on the low word, loop unrolled by 8
    const uint64_t da0 = !swz ? (a_mn ? desc_gen(A, 128, 2048, 0) : desc_gen(A, 2048, 128, 0)) : (a_mn ? desc_gen(A, 16384, 1024, 2) : desc_gen(A, 16, 1024, 2));
    const uint64_t db0 = !swz ? (b_mn ? desc_gen(B, 128, 2048, 0) : desc_gen(B, 2048, 128, 0)) : (b_mn ? desc_gen(B, 16384, 1024, 2) : desc_gen(B, 16, 1024, 2));
    const uint32_t a_lo = (uint32_t)da0, a_hi = (uint32_t)(da0 >> 32), b_lo = (uint32_t)db0, b_hi = (uint32_t)(db0 >> 32);
    const uint32_t a_step = !swz ? (a_mn ? 16u : 256u) : (a_mn ? 128u : 2u);   // in 16-byte units per K step (sw128 K-major: 32 B; block switch ignored here)
    const uint32_t b_step = !swz ? (b_mn ? 16u : 256u) : (b_mn ? 128u : 2u);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t dcol = tmem + (uint32_t)((n_acc > 1) ? (kk & 1) * 256 : 0);
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(dcol), "r"(a_lo + kk * a_step), "r"(a_hi), "r"(b_lo + kk * b_step), "r"(b_hi), "r"(idesc), "r"((uint32_t)(i + kk > 1))
            : "memory");
      }
    }
    long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  tc_fence_after();
  // TMEM load rate: all 4 warps read their 32 lanes x 128 columns, 16 times
  long long t3 = clock64();
  uint32_t acc = 0;
  for (int r = 0; r < 16; ++r) {
    uint32_t v[32], w[32];
    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16);
    tmem_ld32(ta, v); tmem_ld32(ta + 32, w); tmem_ld_wait();
    for (int j = 0; j < 32; ++j) acc += v[j] ^ w[j];
    tmem_ld32(ta + 64, v); tmem_ld32(ta + 96, w); tmem_ld_wait();
    for (int j = 0; j < 32; ++j) acc += v[j] ^ w[j];
  }
  long long t4 = clock64();
  if (threadIdx.x == 0) { out[2] = t4 - t3; out[3] = acc; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
  const char* names[] = {"noswz K/K", "sw128 K/K", "noswz MN/MN", "sw128 MN/MN", "noswz K/MN", "sw128 K/MN"};
  for (int N : {256, 128, 64, 16}) for (int n_acc : {1, 2}) for (int mode = 0; mode < 6; ++mode) {
    if (N == 256 && mode >= 2) continue;
    for (int rep = 0; rep < 2; ++rep) {
      mma_rate<<<1, 128, 131072 + 1024>>>(mode, 64, N, n_acc, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("N=%3d n_acc=%d %-12s: issue %6lld cyc, issue+exec %6lld cyc for 64 MMAs -> %.1f cyc/MMA ; TMEM ld 16 x 64KB: %lld cyc -> %.0f cyc per 64 KB\n", N, n_acc, names[mode], h[0], h[1],
           h[1] / 64.0, h[2], h[2] / 16.0);
  }
  return 0;
}
