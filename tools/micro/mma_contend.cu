// Microbenchmark: does concurrent tcgen05.ld / tcgen05.st traffic from other warps slow tcgen05.mma down (SS vs TS form)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_contend mma_contend.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../nerf-ca_b200/csrc/tc_common.cuh"
using namespace nerfca::tc;

// hammer: 0 none, 1 tcgen05.ld.x32 loops, 2 tcgen05.st.x16 loops, 3 ld + st (epilogue-like), 4 LDS.128 loops; n_hammer warps (<= 8)
__global__ void __launch_bounds__(640, 1) mma_contend(int ts, int hammer, int n_hammer, int n_mma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) done = 0;
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
  if (warp == 0) {
    if (elect_one()) {
      const Desc a = kmajor(smem_u32(smem)), b = kmajor(smem_u32(smem + 32768));
      const uint32_t idesc = instr_desc(128, 128, 0, 0);
      long long t0 = clock64();
      for (int i = 0; i < n_mma; i += 8) {
        if (ts) umma_ts_k<8, KSTEP_KMAJOR>(tmem, tmem + 256, b, idesc, 0);
        else umma_k<8, KSTEP_KMAJOR, KSTEP_KMAJOR>(tmem, a, b, idesc, 0);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      long long t1 = clock64();
      out[0] = t1 - t0;
      done = 1;
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 4 + n_hammer) {
    const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128;   // columns 128.. (not the accumulator in use)
    uint32_t acc = 0;
    long long n = 0;
    while (!done) {
      if (hammer == 1 || hammer == 3) { uint32_t v[32]; tmem_ld32(ta + (n & 1) * 32, v); tmem_ld_wait(); for (int j = 0; j < 32; ++j) acc += v[j]; }
      if (hammer == 2 || hammer == 3) { uint32_t v[16]; for (int j = 0; j < 16; ++j) v[j] = acc + j; tmem_st16(ta + 192, v); tmem_st_wait(); }
      if (hammer == 5) { uint32_t v[32], w[32]; tmem_ld32(ta, v); tmem_ld32(ta + 32, w); tmem_ld_wait(); for (int j = 0; j < 32; ++j) acc += v[j] ^ w[j]; }
      if (hammer == 4) { for (int j = 0; j < 8; ++j) { uint32_t x0, x1, x2, x3; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(smem_u32(smem + 1024 * j + (warp & 3) * 16))); acc += x0 ^ x3; } }
      ++n;
    }
    if (lane == 0) { out[1 + (warp - 4)] = n; out[12] = acc; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 128);
  cudaFuncSetAttribute(mma_contend, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  const char* hn[] = {"none", "ld.x32", "st.x16", "ld+st", "lds", "2xld.x32"};
  for (int ts = 0; ts < 2; ++ts) for (int hammer : {0, 1, 5, 2}) for (int nh : {4, 8, 16}) {
    if (hammer == 0 && nh != 4) continue;
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d, 0, 128);
      mma_contend<<<1, 640, 65536 + 1024>>>(ts, hammer, nh, 512, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
    double bytes = (hammer == 5 ? 8192.0 : hammer == 1 ? 4096.0 : hammer == 2 ? 2048.0 : 0.0) * h[1] * nh;
    printf("%s hammer=%-8s x%2d warps: %6.1f cyc/MMA (512 MMAs N=128); iterations per warp %4lld -> %6.1f B/clk/SM TMEM traffic by the hammer warps\n", ts ? "TS" : "SS", hn[hammer], nh, h[0] / 512.0, h[1], bytes / h[0]);
  }
  return 0;
}
