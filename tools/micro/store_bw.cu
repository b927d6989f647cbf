// Microbenchmark: streaming write bandwidth to HBM as a function of the number of CTAs (1 per SM) and the store mechanism:
//   mode 0: st.global.cs.v4 from registers (what the forward's stash uses)      mode 1: plain st.global.v4
//   mode 2: cp.async.bulk shared -> global, 32 KB per operation                  mode 3: same, 4 KB per operation
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu && ./store_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) store_kernel(uint8_t* out, size_t bytes_per_cta, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* dst = out + (size_t)blockIdx.x * bytes_per_cta;
  if (mode <= 1) {
    const uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    for (size_t off = (size_t)threadIdx.x * 16; off < bytes_per_cta; off += (size_t)blockDim.x * 16) {
      if (mode == 0) __stcs(reinterpret_cast<uint4*>(dst + off), v);
      else *reinterpret_cast<uint4*>(dst + off) = v;
    }
  } else {
    const uint32_t chunk = mode == 2 ? 32768u : 4096u;
    for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, 1, 2, 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(smem);
      int inflight = 0;
      for (size_t off = 0; off < bytes_per_cta; off += chunk) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(src + (uint32_t)(off % 32768)), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
}

int main() {
  const size_t total = (size_t)2 << 30;   // 2 GiB per launch
  uint8_t* buf;
  cudaMalloc(&buf, total);
  cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[4] = {"st.global.cs.v4", "st.global.v4", "bulk 32 KB", "bulk 4 KB"};
  for (int mode = 0; mode < 4; ++mode)
    for (int ctas : {37, 74, 148}) {
      const size_t per = (total / ctas) & ~(size_t)32767;
      store_kernel<<<ctas, 512, 32768>>>(buf, per, mode);
      cudaEventRecord(e0);
      for (int r = 0; r < 3; ++r) store_kernel<<<ctas, 512, 32768>>>(buf, per, mode);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double gbs = 3.0 * per * ctas / (ms * 1e-3) / 1e9;
      printf("%-16s %3d CTAs: %7.1f GB/s  (%5.1f B/clk/SM at 1965 MHz)  %s\n", names[mode], ctas, gbs, gbs * 1e9 / ctas / 1.965e9,
             cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
