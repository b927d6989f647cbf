// bf16 tensor-core path for field shapes OUTSIDE the fused tcgen05 kernels of mlp_tc.cu (hidden != 128, first layers wider than 95
// features, other layer counts: BASELINE config 5, hidden 256 / 16 bands): the layer stack as individual tcgen05 GEMMs.
//
// Same orchestration as the fp32 SIMT path (mlp_simt.cu: forward = bias + ReLU GEMM per layer, backward = split-K weight-gradient
// GEMM + masked dgrad GEMM per layer), but
//   * activations and gradients live in HBM as bf16 ROW-MAJOR matrices [samples, features] (half the traffic of the fp32 path),
//     the weights are converted to bf16 row-major copies once per call,
//   * every GEMM is one kernel template with three modes
//         forward   C = relu(A W^T + b)   A K-major,  W K-major     (reduction along the contiguous dimension of both)
//         dgrad     C = (A W) * 1[h > 0]  A K-major,  W MN-major    (+ the column sums of C = the next bias gradient)
//         wgrad     C += A^T H            A MN-major, H MN-major    (reduction = samples, split over tiles, fp32 vector reductions)
//     wide_gemm2_kernel<MODE> (the default): persistent, one CTA per SM, 128 x 128 output tiles, warp-specialised -- one lane feeds a
//     5-stage ring of 32 KB stages with 128-byte-swizzled tensor-map copies (TMA), one lane issues tcgen05.mma (M128 N128 K16,
//     kind::f16) into one of two TMEM accumulators, 8 epilogue warps drain the other (tcgen05.ld -> bias / ReLU / mask -> 128-byte-swizzled
//     shared-memory tile -> tensor-map store; the ReLU-mask tile of a dgrad comes in the same way).  Depending on which matrix dimension is the reduction the same row-major bytes
//     are loaded as a K-major block (one 64 x 128 box) or an MN-major block (two 64 x 64 boxes), so no operand is ever transposed.
//     wide_gemm_kernel<MODE> (fallback when the driver has no cuTensorMapEncodeTiled, or NERFCA_WIDE_TMA=0): one tile per CTA,
//     operands staged by 16-byte cp.async copies straight into the UMMA no-swizzle canonical layout (tc_common.cuh), 3-stage ring.
//   * the small kernels around the GEMMs (encoder: one thread per sample; output layer forward / backward; latent-gradient scatter)
//     are single passes over their bf16 matrix.
// Config 5 (1024 rays x 256 samples, hidden 256, 16 bands): 2.3 ms per step = 22 % of the tensor roofline (the per-thread cp.async version
// of round 2a: 7.7 ms); the GEMMs run at 57-78 us against an HBM bound of 41-62 us per 262144 x 256 x 256 layer.
//
// Reference: model/CPPN.py:88-110, model/Temporal.py:113-151 and their autograd.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace nerfca {

using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int W_STAGES = 3;
constexpr uint32_t W_OPER_BYTES = 16384;               // one operand block: 128 x 64 (K-major) or 64 x 128 (MN-major) bf16
constexpr uint32_t W_STAGE_BYTES = 2 * W_OPER_BYTES;
constexpr size_t W_SMEM = (size_t)W_STAGES * W_STAGE_BYTES + (W_STAGES + 1) * 8 + 16;
enum { W_FWD = 0, W_DGRAD = 1, W_WGRAD = 2 };

struct WideGemm {
  const bf16* A; long long lda, a_rows, a_cols;       // row-major matrices as stored (cols are multiples of 8, zero padded)
  const bf16* B; long long ldb, b_rows, b_cols;
  long long M; int N; long long K;                    // C is [M, N], reduction length K
  bf16* C; long long ldc;                             // W_FWD / W_DGRAD output (bf16 rows)
  const float* bias;                                  // W_FWD: [N] or null
  const bf16* mask; long long ldm;                    // W_DGRAD: gradient passes where mask > 0 (same shape as C)
  float* colsum;                                      // W_DGRAD (persistent kernel only), optional: [N] += column sums of the bf16 output (the next bias gradient)
  float* C32; long long ldc32;                        // W_WGRAD output (fp32, atomically accumulated)
  long long k_split;                                  // W_WGRAD: reduction elements per blockIdx.z (multiple of 64)
};

constexpr int W_ROWS_PER_BLOCK = 128;
__global__ void wide_colsum_kernel(const bf16* __restrict__ dz, long long np, int H, float* __restrict__ db);

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Copies the [RB rows x 8 * CB cols] block at (r0, c0) of a row-major bf16 matrix into the canonical layout
// byte(row, col) = (col / 8) * (RB * 16) + row * 16 + (col % 8) * 2; everything outside the matrix is zero-filled.  A warp's 32 pieces
// are 16 consecutive rows x 2 adjacent chunks: full 32-byte sectors on the global side, conflict-free quarter-warps on the shared side.
template <int RB, int CB>
__device__ __forceinline__ void load_block(const bf16* mat, long long ld, long long n_rows, long long n_cols, long long r0, long long c0,
                                           uint32_t smem_base) {
  constexpr int PIECES = RB * CB;
  for (int p = threadIdx.x; p < PIECES; p += blockDim.x) {
    const int sub = p & 31, grp = p >> 5;
    const int row = (grp % (RB / 16)) * 16 + (sub & 15);
    const int chunk = (grp / (RB / 16)) * 2 + (sub >> 4);
    const long long r = r0 + row, c = c0 + chunk * 8;
    const bool ok = r < n_rows && c < n_cols;
    cp_async16(smem_base + (uint32_t)chunk * (RB * 16) + (uint32_t)row * 16, ok ? (const void*)(mat + r * ld + c) : (const void*)mat, ok ? 16u : 0u);
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) wide_gemm_kernel(WideGemm g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + (size_t)W_STAGES * W_STAGE_BYTES);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + W_STAGES + 1);
  const uint32_t bar0 = smem_u32(s_bar), bar_done = bar0 + 8 * W_STAGES;   // bar0 + 8 s: the MMAs that read stage s have completed
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s <= W_STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), 128);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const long long m0 = (long long)blockIdx.x * 128;
  const int n0 = blockIdx.y * 128;
  const long long kbeg = (MODE == W_WGRAD) ? (long long)blockIdx.z * g.k_split : 0;
  const long long kend = (MODE == W_WGRAD && kbeg + g.k_split < g.K) ? kbeg + g.k_split : g.K;
  const int nkb = (int)((kend - kbeg + 63) / 64);
  constexpr uint32_t idesc = (MODE == W_FWD) ? instr_desc(128, 128, 0, 0) : (MODE == W_DGRAD) ? instr_desc(128, 128, 0, 1) : instr_desc(128, 128, 1, 1);
  const uint32_t s_base = smem_u32(smem);

  auto issue_load = [&](int kb) {
    const int s = kb % W_STAGES;
    if (kb >= W_STAGES) mbar_wait(bar0 + 8 * s, (uint32_t)((kb / W_STAGES - 1) & 1));   // the MMAs of block kb - W_STAGES have left the stage
    const long long k0 = kbeg + (long long)kb * 64;
    const uint32_t sa = s_base + (uint32_t)s * W_STAGE_BYTES, sb = sa + W_OPER_BYTES;
    if (MODE == W_WGRAD) load_block<64, 16>(g.A, g.lda, (kend < g.a_rows ? kend : g.a_rows), g.a_cols, k0, m0, sa);   // rows = samples (reduction), cols = M
    else load_block<128, 8>(g.A, g.lda, g.a_rows, g.a_cols, m0, k0, sa);                                             // rows = M, cols = reduction
    if (MODE == W_FWD) load_block<128, 8>(g.B, g.ldb, g.b_rows, g.b_cols, n0, k0, sb);                               // rows = N, cols = reduction
    else load_block<64, 16>(g.B, g.ldb, (MODE == W_WGRAD && kend < g.b_rows) ? kend : g.b_rows, g.b_cols, k0, n0, sb);   // rows = reduction, cols = N
  };
  for (int kb = 0; kb < W_STAGES - 1; ++kb) {
    if (kb < nkb) issue_load(kb);
    cp_async_commit();
  }
  for (int kb = 0; kb < nkb; ++kb) {
    cp_async_wait<W_STAGES - 2>();        // this thread's pieces of block kb have landed
    fence_proxy_async();                   // ... and are visible to the tensor core's shared-memory reads
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const int s = kb % W_STAGES;
      const uint32_t sa = s_base + (uint32_t)s * W_STAGE_BYTES, sb = sa + W_OPER_BYTES;
      const long long kvalid = (kend - (kbeg + (long long)kb * 64) < 64) ? kend - (kbeg + (long long)kb * 64) : 64;
      const int ksteps = (int)((kvalid + 15) / 16);
      for (int kk = 0; kk < ksteps; ++kk) {
        // K-major [128 x 64] block: reduction step = 2 chunks of 2048 B; MN-major [64 x 128] block: reduction step = 16 rows of 16 B, chunks 1024 B apart
        const uint64_t da = (MODE == W_WGRAD) ? smem_desc(sa + kk * 256, 128, 1024) : smem_desc(sa + kk * 4096, 2048, 128);
        const uint64_t db = (MODE == W_FWD) ? smem_desc(sb + kk * 4096, 2048, 128) : smem_desc(sb + kk * 256, 128, 1024);
        umma_ss(tmem, da, db, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(bar0 + 8 * s);
      if (kb == nkb - 1) umma_commit(bar_done);
    }
    if (kb + W_STAGES - 1 < nkb) issue_load(kb + W_STAGES - 1);
    cp_async_commit();                     // (an empty group keeps the group count uniform)
  }
  cp_async_wait<0>();

  // ---- epilogue: thread = (accumulator row, column half) ----
  if (nkb > 0) {
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int q = warp & 3, ch = warp >> 2;
    const int row = q * 32 + lane;
    const long long m = m0 + row;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ch * 64 + half * 32, v);
      tmem_ld_wait();
      if (m >= g.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + ch * 64 + half * 32 + 8 * j;
        if (n >= g.N) continue;
        if (MODE == W_WGRAD) {
          float* dst = g.C32 + m * g.ldc32 + n;
          if (n + 8 <= g.N && (g.ldc32 & 3) == 0 && (reinterpret_cast<uintptr_t>(g.C32) & 15) == 0) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(v[8 * j])), "f"(__uint_as_float(v[8 * j + 1])),
                         "f"(__uint_as_float(v[8 * j + 2])), "f"(__uint_as_float(v[8 * j + 3])) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(__uint_as_float(v[8 * j + 4])), "f"(__uint_as_float(v[8 * j + 5])),
                         "f"(__uint_as_float(v[8 * j + 6])), "f"(__uint_as_float(v[8 * j + 7])) : "memory");
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (n + e < g.N) atomicAdd(dst + e, __uint_as_float(v[8 * j + e]));
          }
        } else {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
          uint4 o;
          if (MODE == W_FWD) {
            if (g.bias) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += __ldg(g.bias + n + e);
            }
            o = make_uint4(pack_relu_bf16x2(f[0], f[1]), pack_relu_bf16x2(f[2], f[3]), pack_relu_bf16x2(f[4], f[5]), pack_relu_bf16x2(f[6], f[7]));
          } else {
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(g.mask + m * g.ldm + n));
            o = make_uint4(mul_bf16x2(pack_bf16x2(f[0], f[1]), relu_mask_bf16x2(h.x)), mul_bf16x2(pack_bf16x2(f[2], f[3]), relu_mask_bf16x2(h.y)),
                           mul_bf16x2(pack_bf16x2(f[4], f[5]), relu_mask_bf16x2(h.z)), mul_bf16x2(pack_bf16x2(f[6], f[7]), relu_mask_bf16x2(h.w)));
          }
          *reinterpret_cast<uint4*>(g.C + m * g.ldc + n) = o;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- persistent TMA-fed version of the same GEMM (the default; NERFCA_WIDE_TMA=0 or a driver without cuTensorMapEncodeTiled falls back
// to wide_gemm_kernel above) --------------------------------------------------------------------------------------------------------
// One CTA per SM walks the output tiles (n fastest, so the CTAs that share an A block run side by side and the second fetch is an L2
// hit).  Warp 0: one lane issues the operand loads as 2-D tensor-map copies with the 128-byte swizzle: a box of 64 contiguous bf16 (one
// 128-byte swizzle row) x 128 rows is a K-major operand block, two boxes of 64 x 64 are an MN-major one (UMMA canonical layouts
// SWIZZLE_128B: K-major ((8,n),2):((8,SBO),1), MN-major ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units, SBO = 1024 B, LBO = 8192 B);
// out-of-range rows / columns are zero-filled by the copy engine.  (A 3-D map [cols / 8][rows][8] that lands the bytes in the no-swizzle
// chunk layout of wide_gemm_kernel works too but moves 16 bytes per request: 170 us per GEMM, r6b.)  Warp 1: one lane issues the MMAs of a tile into one of two 128-column accumulators and commits
// the stage's "empty" and the accumulator's "full" mbarriers.  Warps 2-9: epilogue of the OTHER accumulator meanwhile.  Six 32 KB
// stages keep ~190 KB of loads in flight per SM; the per-thread cp.async version above had 64 KB and a CTA-wide barrier + proxy
// fence per K block (r5w: 250 us per 262144 x 256 x 256 GEMM, tensor pipe 6.5 %).
// operand stages per mode: what the mask / output tiles leave of the shared memory (dgrad: 3 stages + 2 mask tiles + output tile,
// forward: 5 stages + output tile, wgrad: 6 stages)
__host__ __device__ constexpr int w2_stages(int mode) { return mode == W_DGRAD ? 3 : (mode == W_FWD ? 5 : 6); }
constexpr int W2_THREADS = 10 * 32;
// [128 x 128] bf16 tiles that enter or leave through the copy engine: two 64-column halves of [128 rows x 128 B], 128-byte swizzle (the
// 16-byte chunk j of row r sits at r * 128 + ((j ^ (r & 7)) << 4)): conflict-free for a thread that owns a row, and what a box of the tensor map is
constexpr uint32_t W2_IO_TILE = 32768;
__host__ __device__ constexpr uint32_t w2_io_bytes(int mode) {   // ReLU-mask tiles of the next two tiles (W_DGRAD) + the staged output tile
  return mode == W_DGRAD ? 3 * W2_IO_TILE : (mode == W_FWD ? W2_IO_TILE : 0);
}
constexpr uint32_t W2_CS_BYTES = 8 * 128 * 4;          // column-sum exchange: [8 row parts][128 columns] fp32
__host__ __device__ constexpr size_t w2_smem(int mode) {
  return (size_t)w2_stages(mode) * W_STAGE_BYTES + w2_io_bytes(mode) + W2_CS_BYTES + (2 * w2_stages(mode) + 8) * 8 + 16;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int col, int row, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(col), "r"(row)
               : "memory");
}
// operand block at (row0, col0) of a row-major matrix: K-major = [128 rows x 64 reduction cols] in one box, MN-major = [64 reduction rows x
// 128 cols] as two 64-column boxes 8 KB apart
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(col), "r"(row)
               : "memory");
}
__device__ __forceinline__ uint32_t io_chunk(uint32_t tile, int row, int half, int j) {
  return tile + (uint32_t)half * 16384u + (uint32_t)row * 128u + ((uint32_t)(j ^ (row & 7)) << 4);
}
__device__ __forceinline__ void load_kmajor(uint32_t dst, const CUtensorMap* tm, int row0, int col0, uint32_t bar) { tma_load_2d(dst, tm, col0, row0, bar); }
__device__ __forceinline__ void load_mnmajor(uint32_t dst, const CUtensorMap* tm, int row0, int col0, uint32_t bar) {
  tma_load_2d(dst, tm, col0, row0, bar);
  tma_load_2d(dst + 8192u, tm, col0 + 64, row0, bar);
}
// UMMA descriptors of those blocks (SWIZZLE_128B = layout type 2), reduction step kk of 16
__device__ __forceinline__ uint64_t desc_sw128_k(uint32_t base, int kk) { return smem_desc(base + kk * 32, 16, 1024) | (2ull << 61); }
__device__ __forceinline__ uint64_t desc_sw128_mn(uint32_t base, int kk) { return smem_desc(base + kk * 2048, 8192, 1024) | (2ull << 61); }

template <int MODE>
__global__ void __launch_bounds__(W2_THREADS, 1) wide_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                   const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmM,
                                                                   WideGemm g, int n_mt, int n_nt, int n_sp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int W2_STAGES = w2_stages(MODE);
  constexpr uint32_t W2_IO_BYTES = w2_io_bytes(MODE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  [[maybe_unused]] float* s_cs = reinterpret_cast<float*>(smem + (size_t)W2_STAGES * W_STAGE_BYTES + W2_IO_BYTES);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + (size_t)W2_STAGES * W_STAGE_BYTES + W2_IO_BYTES + W2_CS_BYTES);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * W2_STAGES + 8);
  [[maybe_unused]] const uint32_t s_mask = smem_u32(smem) + W2_STAGES * W_STAGE_BYTES;  // two ReLU-mask tiles (W_DGRAD), filled by the loader lane
  [[maybe_unused]] const uint32_t s_out = s_mask + (MODE == W_DGRAD ? 2 * W2_IO_TILE : 0);   // staged output tile, drained by a tensor-map store
  const uint32_t full0 = smem_u32(s_bar), empty0 = full0 + 8 * W2_STAGES, accf0 = empty0 + 8 * W2_STAGES, acce0 = accf0 + 16;
  [[maybe_unused]] const uint32_t mfull0 = acce0 + 16, mempty0 = mfull0 + 16;
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < W2_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(accf0 + 8 * b, 1); mbar_init(acce0 + 8 * b, 8); mbar_init(mfull0 + 8 * b, 1); mbar_init(mempty0 + 8 * b, 8); }
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t s_base = smem_u32(smem);
  const long long total = (long long)n_mt * n_nt * n_sp;
  // tile t -> (n tile, m tile, reduction split); the K range of a tile in 64-wide blocks
  auto k_range = [&](long long t, long long& m0, int& n0, long long& kbeg, long long& kend) {
    const int nt = (int)(t % n_nt);
    const long long rest = t / n_nt;
    const long long mt = rest % n_mt, sp = rest / n_mt;
    m0 = mt * 128; n0 = nt * 128;
    kbeg = (MODE == W_WGRAD) ? sp * g.k_split : 0;
    kend = (MODE == W_WGRAD && kbeg + g.k_split < g.K) ? kbeg + g.k_split : g.K;
  };

  if (warp == 0) {
    if (lane == 0) {   // ================= operand loads =================
      uint32_t it = 0, tl = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x, ++tl) {
        long long m0, kbeg, kend; int n0;
        k_range(t, m0, n0, kbeg, kend);
        const int nkb = (int)((kend - kbeg + 63) / 64);
        if (MODE == W_DGRAD) {      // the tile's ReLU-mask tile (two buffers: the epilogue released this one two tiles ago)
          const uint32_t b = tl & 1;
          if (tl >= 2) mbar_wait(mempty0 + 8 * b, ((tl >> 1) - 1) & 1);
          mbar_expect_tx(mfull0 + 8 * b, W2_IO_TILE);
          tma_load_2d(s_mask + b * W2_IO_TILE, &tmM, n0, (int)m0, mfull0 + 8 * b);
          tma_load_2d(s_mask + b * W2_IO_TILE + 16384u, &tmM, n0 + 64, (int)m0, mfull0 + 8 * b);
        }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % W2_STAGES;
          if (it >= W2_STAGES) mbar_wait(empty0 + 8 * s, ((it / W2_STAGES) - 1) & 1);   // the MMAs that read the stage's previous block are done
          const uint32_t sa = s_base + s * W_STAGE_BYTES, sb = sa + W_OPER_BYTES, bar = full0 + 8 * s;
          const int k0 = (int)(kbeg + (long long)kb * 64);
          mbar_expect_tx(bar, W_STAGE_BYTES);
          if (MODE == W_WGRAD) load_mnmajor(sa, &tmA, k0, (int)m0, bar);               // rows = reduction (samples), cols = M
          else load_kmajor(sa, &tmA, (int)m0, k0, bar);                                 // rows = M, cols = reduction
          if (MODE == W_FWD) load_kmajor(sb, &tmB, n0, k0, bar);                        // rows = N, cols = reduction
          else load_mnmajor(sb, &tmB, k0, n0, bar);                                     // rows = reduction, cols = N
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ================= MMA issue =================
      constexpr uint32_t idesc = (MODE == W_FWD) ? instr_desc(128, 128, 0, 0) : (MODE == W_DGRAD) ? instr_desc(128, 128, 0, 1) : instr_desc(128, 128, 1, 1);
      uint32_t it = 0, tl = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x, ++tl) {
        long long m0, kbeg, kend; int n0;
        k_range(t, m0, n0, kbeg, kend);
        const int nkb = (int)((kend - kbeg + 63) / 64);
        const uint32_t b = tl & 1;
        if (tl >= 2) mbar_wait(acce0 + 8 * b, ((tl >> 1) - 1) & 1);   // the epilogue has read the accumulator's previous tile
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % W2_STAGES;
          mbar_wait(full0 + 8 * s, (it / W2_STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = s_base + s * W_STAGE_BYTES, sb = sa + W_OPER_BYTES;
          const long long left = kend - (kbeg + (long long)kb * 64);
          const int ksteps = (int)(((left < 64 ? left : 64) + 15) / 16);
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t da = (MODE == W_WGRAD) ? desc_sw128_mn(sa, kk) : desc_sw128_k(sa, kk);
            const uint64_t db = (MODE == W_FWD) ? desc_sw128_k(sb, kk) : desc_sw128_mn(sb, kk);
            umma_ss(tmem + b * 128, da, db, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(empty0 + 8 * s);
        }
        umma_commit(accf0 + 8 * b);
      }
    }
  } else {
    // ================= epilogue: thread = (accumulator row, column half); warps 2-9 cover the four TMEM lane quadrants twice.  A thread
    // owns a ROW of the tile, so direct global accesses would touch 32 different lines per warp instruction: the ReLU-mask tile
    // (W_DGRAD) therefore comes in, and the bf16 output tile goes out, through the copy engine and a 128-byte-swizzled shared-memory
    // tile in which a row's eight 16-byte chunks are conflict-free.  (The fp32 weight-gradient tiles -- 512 per GEMM instead of 4096 --
    // go out as vector reductions straight from the registers.) =================
    const int q = warp & 3, ch = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int te = (int)threadIdx.x - 64;                     // 0 .. 255 among the epilogue threads
    float cs_acc = 0.f;                                       // running column sum of column (cs_n0 + te), te < 128, over this CTA's tiles
    int cs_n0 = -1;
    uint32_t tl = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++tl) {
      long long m0, kbeg, kend; int n0;
      k_range(t, m0, n0, kbeg, kend);
      const uint32_t b = tl & 1;
      const long long m = m0 + row;
      uint4 hmask[8];
      if (MODE == W_DGRAD) {
        mbar_wait(mfull0 + 8 * b, (tl >> 1) & 1);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(hmask[j].x), "=r"(hmask[j].y), "=r"(hmask[j].z), "=r"(hmask[j].w)
                       : "r"(io_chunk(s_mask + b * W2_IO_TILE, row, ch, j)) : "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(mempty0 + 8 * b);          // the loader lane may fetch the mask of the tile after next
      }
      mbar_wait(accf0 + 8 * b, (tl >> 1) & 1);
      tc_fence_after();
      if (MODE != W_WGRAD) {
        if (te == 0) bulk_wait_read_all();                     // the previous tile's store has drained the staged tile
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + b * 128 + ch * 64 + half * 32, v);
        tmem_ld_wait();
        if (MODE == W_WGRAD) {
          if (m >= g.M) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = n0 + ch * 64 + half * 32 + 8 * j;
            if (n >= g.N) continue;
            float* dst = g.C32 + m * g.ldc32 + n;
            if (n + 8 <= g.N && (g.ldc32 & 3) == 0 && (reinterpret_cast<uintptr_t>(g.C32) & 15) == 0) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(v[8 * j])), "f"(__uint_as_float(v[8 * j + 1])),
                           "f"(__uint_as_float(v[8 * j + 2])), "f"(__uint_as_float(v[8 * j + 3])) : "memory");
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(__uint_as_float(v[8 * j + 4])), "f"(__uint_as_float(v[8 * j + 5])),
                           "f"(__uint_as_float(v[8 * j + 6])), "f"(__uint_as_float(v[8 * j + 7])) : "memory");
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (n + e < g.N) atomicAdd(dst + e, __uint_as_float(v[8 * j + e]));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = n0 + ch * 64 + half * 32 + 8 * j;
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
            uint4 o;
            if (MODE == W_FWD) {
              if (g.bias && n < g.N) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + n)), b1 = __ldg(reinterpret_cast<const float4*>(g.bias + n + 4));
                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
              }
              o = make_uint4(pack_relu_bf16x2(f[0], f[1]), pack_relu_bf16x2(f[2], f[3]), pack_relu_bf16x2(f[4], f[5]), pack_relu_bf16x2(f[6], f[7]));
            } else {
              const uint4 h = hmask[half * 4 + j];
              o = make_uint4(mul_bf16x2(pack_bf16x2(f[0], f[1]), relu_mask_bf16x2(h.x)), mul_bf16x2(pack_bf16x2(f[2], f[3]), relu_mask_bf16x2(h.y)),
                             mul_bf16x2(pack_bf16x2(f[4], f[5]), relu_mask_bf16x2(h.z)), mul_bf16x2(pack_bf16x2(f[6], f[7]), relu_mask_bf16x2(h.w)));
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(io_chunk(s_out, row, ch, half * 4 + j)), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce0 + 8 * b);             // the accumulator has been read: the MMA lane may start the tile after next
      if (MODE != W_WGRAD) {
        fence_proxy_async();                                   // the staged tile is read by the copy engine
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (te == 0) {                                         // rows / columns beyond the matrix are clipped by the store
          tma_store_2d(&tmC, s_out, n0, (int)m0);
          tma_store_2d(&tmC, s_out + 16384u, n0 + 64, (int)m0);
          bulk_commit();
        }
        if (MODE == W_DGRAD && g.colsum) {
          // column sums of the staged tile (rows beyond M are zero): thread = (16-row part, 4 columns), then 128 threads add the 8 parts
          const int rpart = te >> 5, col4 = te & 31;
          float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int r = rpart * 16 + k;
            uint32_t lo, hi;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(io_chunk(s_out, r, col4 >> 4, (col4 & 15) >> 1) + (uint32_t)(col4 & 1) * 8u) : "memory");
            p0 += __uint_as_float(lo << 16); p1 += __uint_as_float(lo & 0xFFFF0000u);
            p2 += __uint_as_float(hi << 16); p3 += __uint_as_float(hi & 0xFFFF0000u);
          }
          *reinterpret_cast<float4*>(s_cs + rpart * 128 + col4 * 4) = make_float4(p0, p1, p2, p3);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (te < 128) {
            if (n0 != cs_n0) {
              if (cs_n0 >= 0 && cs_n0 + te < g.N) atomicAdd(g.colsum + cs_n0 + te, cs_acc);
              cs_n0 = n0; cs_acc = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) cs_acc += s_cs[k * 128 + te];
          }
        }
      }
    }
    if (MODE == W_DGRAD && g.colsum && te < 128 && cs_n0 >= 0 && cs_n0 + te < g.N) atomicAdd(g.colsum + cs_n0 + te, cs_acc);
    if (MODE != W_WGRAD && te == 0) bulk_wait_all();           // the last store has left shared memory (and reached global memory)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* e = getenv("NERFCA_WIDE_TMA");
    if (!(e && e[0] == '0')) {
      void* p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
      else cudaGetLastError();
    }
  }
  return fn;
}
// row-major bf16 matrix [n_rows, n_cols] (ld a multiple of 8) as a 2-D tensor map, box = 64 columns (128 B, one swizzle row) x box_rows
static bool make_operand_map(EncodeTiledFn fn, CUtensorMap* tm, const bf16* mat, long long ld, long long n_rows, long long n_cols, int box_rows) {
  if ((ld & 7) || (reinterpret_cast<uintptr_t>(mat) & 15) || n_rows <= 0 || n_cols <= 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)n_cols, (cuuint64_t)n_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(mat), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static int wide_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int MODE>
static int run_wide_gemm(const WideGemm& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return NERFCA_OK;
  static bool attr_done = false;
  if (!attr_done) {
    NERFCA_CUDA_OK(cudaFuncSetAttribute(wide_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W_SMEM));
    attr_done = true;
  }
  const long long splits = (MODE == W_WGRAD) ? (g.K + g.k_split - 1) / g.k_split : 1;
  if (EncodeTiledFn fn = encode_tiled_fn()) {
    CUtensorMap tmA, tmB;
    CUtensorMap tmC, tmM;
    const bool a_ok = make_operand_map(fn, &tmA, g.A, g.lda, g.a_rows, g.a_cols, (MODE == W_WGRAD) ? 64 : 128);
    const bool b_ok = make_operand_map(fn, &tmB, g.B, g.ldb, g.b_rows, g.b_cols, (MODE == W_FWD) ? 128 : 64);
    bool io_ok = true;
    if (MODE != W_WGRAD) io_ok = make_operand_map(fn, &tmC, g.C, g.ldc, g.M, g.N, 128);
    else tmC = tmA;
    if (MODE == W_DGRAD) io_ok = io_ok && make_operand_map(fn, &tmM, g.mask, g.ldm, g.M, g.N, 128);
    else tmM = tmA;
    if (a_ok && b_ok && io_ok) {
      static bool attr2_done = false;
      if (!attr2_done) {
        NERFCA_CUDA_OK(cudaFuncSetAttribute(wide_gemm2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w2_smem(MODE)));
        attr2_done = true;
      }
      const int n_mt = (int)((g.M + 127) / 128), n_nt = (g.N + 127) / 128;
      const long long total = (long long)n_mt * n_nt * splits;
      const unsigned grid2 = (unsigned)(total < wide_sm_count() ? total : wide_sm_count());
      wide_gemm2_kernel<MODE><<<grid2, W2_THREADS, w2_smem(MODE), st>>>(tmA, tmB, tmC, tmM, g, n_mt, n_nt, (int)splits);
      NERFCA_LAUNCH_OK();
      return NERFCA_OK;
    }
  }
  dim3 grid((unsigned)((g.M + 127) / 128), (unsigned)((g.N + 127) / 128), (unsigned)splits);
  wide_gemm_kernel<MODE><<<grid, 256, W_SMEM, st>>>(g);
  NERFCA_LAUNCH_OK();
  if (MODE == W_DGRAD && g.colsum) {
    wide_colsum_kernel<<<div_up(g.M, W_ROWS_PER_BLOCK), 128, 0, st>>>(g.C, g.M, g.N, g.colsum);
    NERFCA_LAUNCH_OK();
  }
  return NERFCA_OK;
}

// ---- small kernels around the GEMMs (bf16 row-major activations) --------------------------------------------------------------------
// first-layer input [np, Dp] bf16 (A4 + A5 in registers, zero padded to Dp = roundup8(in_dim)): one thread per sample -- the point
// (float64 ray algebra) and the phase are formed once, the row leaves in 16-byte pieces.  (One thread per (sample, feature) re-derived
// the point for every feature: 194 us per field at config 5.)
__global__ void wide_encode_kernel(SampleSrc src, EncDesc enc, int Dp, bf16* __restrict__ out) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= src.n_points) return;
  float x, y, z;
  load_point(src, p, x, y, z);
  const int phase = (enc.n_latent > 0) ? load_phase(src, p) : 0;
  uint4* dst = reinterpret_cast<uint4*>(out + p * Dp);
  for (int c = 0; c < Dp / 8; ++c) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f0 = c * 8 + 2 * j;
      const float a = (f0 < enc.in_dim) ? enc_feature(enc, f0, x, y, z, phase) : 0.f;
      const float b = (f0 + 1 < enc.in_dim) ? enc_feature(enc, f0 + 1, x, y, z, phase) : 0.f;
      w[j] = pack_bf16x2(a, b);
    }
    dst[c] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// every layer's W fp32 [H, K_l] -> bf16 [H, Kp_l] (zero padded), back to back: layer 0 is [H, Dp], the others [H, H]
struct WideWeights { const float* w[NERFCA_MAX_LAYERS]; };
__global__ void wide_cvt_weight_kernel(WideWeights ws, int H, int D, int Dp, int L, bf16* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n0 = (long long)H * Dp;
  if (idx >= n0 + (long long)(L - 1) * H * H) return;
  int l, K, Kp;
  long long e;
  if (idx < n0) { l = 0; K = D; Kp = Dp; e = idx; }
  else { l = 1 + (int)((idx - n0) / ((long long)H * H)); K = Kp = H; e = (idx - n0) % ((long long)H * H); }
  const int r = (int)(e / Kp), k = (int)(e - (long long)r * Kp);
  out[idx] = __float2bfloat16_rn(k < K ? __ldg(ws.w[l] + (size_t)r * K + k) : 0.f);
}
// output layer (hidden -> 1): one warp per sample
__global__ void wide_out_forward_kernel(const bf16* __restrict__ h, const float* __restrict__ w, const float* __restrict__ b, long long np, int H,
                                        float* __restrict__ raw) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= np) return;
  float acc = 0.f;
  for (int c = lane * 8; c < H; c += 256) {      // 16-byte pieces: a warp reads 512 contiguous bytes of the row
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(h + row * H + c));
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c)), w1 = __ldg(reinterpret_cast<const float4*>(w + c + 4));
    acc = fmaf(__uint_as_float(v.x << 16), w0.x, acc); acc = fmaf(__uint_as_float(v.x & 0xFFFF0000u), w0.y, acc);
    acc = fmaf(__uint_as_float(v.y << 16), w0.z, acc); acc = fmaf(__uint_as_float(v.y & 0xFFFF0000u), w0.w, acc);
    acc = fmaf(__uint_as_float(v.z << 16), w1.x, acc); acc = fmaf(__uint_as_float(v.z & 0xFFFF0000u), w1.y, acc);
    acc = fmaf(__uint_as_float(v.w << 16), w1.z, acc); acc = fmaf(__uint_as_float(v.w & 0xFFFF0000u), w1.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) raw[row] = acc + (b ? __ldg(b) : 0.f);
}
// backward of the output layer: dZ[p,k] = d_raw[p] w[k] 1[h>0];  dW[k] += sum_p d_raw[p] h[p,k];  db += sum_p d_raw[p];
// db_prev[k] += sum_p dZ[p,k] (the bias gradient of the last hidden layer).  Block = 128 rows; thread = (row lane 0..7, 8 columns): 16-byte
// loads / stores, 512 contiguous bytes per warp and row; the eight row lanes meet in shared memory before the atomics.
__global__ void __launch_bounds__(256) wide_out_backward_kernel(const float* __restrict__ d_raw, const bf16* __restrict__ h, const float* __restrict__ w,
                                                                long long np, int H, bf16* __restrict__ dz, float* __restrict__ dw, float* __restrict__ db,
                                                                float* __restrict__ db_prev) {
  __shared__ float s_red[2][8][256];
  const long long r0 = (long long)blockIdx.x * W_ROWS_PER_BLOCK;
  const long long r1 = (r0 + W_ROWS_PER_BLOCK < np) ? r0 + W_ROWS_PER_BLOCK : np;
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  float ab = 0.f;
  for (int c0 = 0; c0 < H; c0 += 256) {
    const int c = c0 + cl * 8;
    float aw[8], az[8], wk[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { aw[e] = az[e] = 0.f; wk[e] = (c + e < H) ? __ldg(w + c + e) : 0.f; }
    if (c < H) {
      for (long long r = r0 + rl; r < r1; r += 8) {
        const float g = __ldg(d_raw + r);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(h + r * H + c));
        const uint32_t hw[4] = {v.x, v.y, v.z, v.w};
        uint32_t zo[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float h0 = __uint_as_float(hw[e2] << 16), h1 = __uint_as_float(hw[e2] & 0xFFFF0000u);
          const bf16 z0 = __float2bfloat16_rn(h0 > 0.f ? g * wk[2 * e2] : 0.f), z1 = __float2bfloat16_rn(h1 > 0.f ? g * wk[2 * e2 + 1] : 0.f);
          az[2 * e2] += __bfloat162float(z0); az[2 * e2 + 1] += __bfloat162float(z1);
          aw[2 * e2] = fmaf(g, h0, aw[2 * e2]); aw[2 * e2 + 1] = fmaf(g, h1, aw[2 * e2 + 1]);
          zo[e2] = (uint32_t)__bfloat16_as_ushort(z0) | ((uint32_t)__bfloat16_as_ushort(z1) << 16);
        }
        *reinterpret_cast<uint4*>(dz + r * H + c) = make_uint4(zo[0], zo[1], zo[2], zo[3]);
        if (c0 == 0 && cl == 0) ab += g;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { s_red[0][rl][cl * 8 + e] = aw[e]; s_red[1][rl][cl * 8 + e] = az[e]; }
    __syncthreads();
    {
      const int k = c0 + (int)threadIdx.x;
      if (k < H) {
        float tw = 0.f, tz = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) { tw += s_red[0][q][threadIdx.x]; tz += s_red[1][q][threadIdx.x]; }
        atomicAdd(dw + k, tw);
        if (db_prev) atomicAdd(db_prev + k, tz);
      }
    }
    __syncthreads();
  }
  if (db) {                     // the eight row lanes' shares of sum_p d_raw[p]  (db is uniform over the block)
    if (cl == 0) s_red[0][0][rl] = ab;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int q = 0; q < 8; ++q) t += s_red[0][0][q];
      atomicAdd(db, t);
    }
  }
}
// db[n] += sum_p dz[p,n]
__global__ void wide_colsum_kernel(const bf16* __restrict__ dz, long long np, int H, float* __restrict__ db) {
  const long long r0 = (long long)blockIdx.x * W_ROWS_PER_BLOCK;
  const long long r1 = (r0 + W_ROWS_PER_BLOCK < np) ? r0 + W_ROWS_PER_BLOCK : np;
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    float a = 0.f;
    for (long long r = r0; r < r1; ++r) a += __bfloat162float(dz[r * H + k]);
    atomicAdd(db + k, a);
  }
}
// d time_latents[phase[p], t] += sum_n dz0[p,n] W0[n, enc_dim + t]   (scatter-add of Temporal.py:144-147's gather)
// A skinny GEMM [np x H] x [H x T] followed by a scatter by phase, bound by the one pass over dz0: a warp walks 32 consecutive samples,
// lane = 8 consecutive columns of the row (16-byte loads, 512 contiguous bytes per warp and row), the latent columns of W0 sit in shared
// memory as [H][T] fp32.  Consecutive samples belong to the same ray and hence the same phase almost always, so the lanes keep private
// partial sums while the phase does not change and only then reduce over the warp (shuffles) and add to the block's [phase][t] table.
constexpr int WL_ROWS_PER_WARP = 32;
constexpr int WL_MAX_T = 16;
__global__ void __launch_bounds__(256) wide_latent_grad_kernel(const bf16* __restrict__ dz0, const float* __restrict__ w0, SampleSrc src, int H, int D,
                                                               int enc_dim, int T, int n_phases, int use_smem, float* __restrict__ dlat) {
  extern __shared__ float wl_smem[];        // [roundup256(H) * T] latent columns of W0, then (use_smem) [n_phases * T] accumulators
  float* s_w = wl_smem;
  float* sacc = wl_smem + ((H + 255) & ~255) * T;
  // W0's latent columns, laid out so that the 32 lanes of a warp (lane = columns 8 lane .. 8 lane + 7 of a 256-column group) read 32
  // consecutive words: word(n, t) = (n / 256) * 256 T + (8 t + n % 8) * 32 + (n / 8) % 32
  for (int i = threadIdx.x; i < H * T; i += blockDim.x) {
    const int n = i / T, t = i - n * T;
    s_w[(n >> 8) * (T << 8) + (((t << 3) + (n & 7)) << 5) + ((n >> 3) & 31)] = __ldg(w0 + (size_t)n * D + enc_dim + t);
  }
  if (use_smem)
    for (int i = threadIdx.x; i < n_phases * T; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * WL_ROWS_PER_WARP;
  float acc[WL_MAX_T];
#pragma unroll
  for (int t = 0; t < WL_MAX_T; ++t) acc[t] = 0.f;
  int cur = -1;
  auto flush = [&]() {
    if (cur >= 0 && cur < n_phases) {
#pragma unroll
      for (int t = 0; t < WL_MAX_T; ++t) {
        if (t < T) {
          float v = acc[t];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) atomicAdd(use_smem ? &sacc[cur * T + t] : dlat + (size_t)cur * T + t, v);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < WL_MAX_T; ++t) acc[t] = 0.f;
  };
  if (H <= 256 && T <= 8) {
    // common case (config 5: H = 256, T = 8): the lane's 8 x T slice of W0 lives in registers
    float wreg[8][8];
    const bool live = lane * 8 < H;
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int t = 0; t < 8; ++t) wreg[e][t] = (live && t < T) ? s_w[(((t << 3) + e) << 5) + lane] : 0.f;
    for (int k0 = 0; k0 < WL_ROWS_PER_WARP; k0 += 4) {      // four rows' loads in flight at a time
      uint4 v4[4];
      int ph4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long p = r0 + k0 + u;
        const bool ok = p < src.n_points;
        ph4[u] = ok ? load_phase(src, p) : -2;                 // warp-uniform; -2: past the end
        v4[u] = (ok && live) ? __ldg(reinterpret_cast<const uint4*>(dz0 + p * H + lane * 8)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ph4[u] == -2) break;
        if (ph4[u] != cur) { flush(); cur = ph4[u]; }
        const uint32_t w4[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = __uint_as_float((e & 1) ? (w4[e >> 1] & 0xFFFF0000u) : (w4[e >> 1] << 16));
#pragma unroll
          for (int t = 0; t < 8; ++t) acc[t] = fmaf(d, wreg[e][t], acc[t]);
        }
      }
    }
  } else
  for (int k = 0; k < WL_ROWS_PER_WARP; ++k) {
    const long long p = r0 + k;
    if (p >= src.n_points) break;
    const int ph = load_phase(src, p);           // warp-uniform
    if (ph != cur) { flush(); cur = ph; }
    for (int c = lane * 8; c < H; c += 256) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(dz0 + p * H + c));
      const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = __uint_as_float((e & 1) ? (w4[e >> 1] & 0xFFFF0000u) : (w4[e >> 1] << 16));
        const float* wr = s_w + (c >> 8) * (T << 8) + (e << 5) + lane;
#pragma unroll
        for (int t = 0; t < WL_MAX_T; ++t)
          if (t < T) acc[t] = fmaf(d, wr[t << 8], acc[t]);
      }
    }
  }
  flush();
  if (!use_smem) return;
  __syncthreads();
  for (int i = threadIdx.x; i < n_phases * T; i += blockDim.x)
    if (sacc[i] != 0.f) atomicAdd(dlat + i, sacc[i]);
}

// ---- host side -------------------------------------------------------------------------------------------------------------------
constexpr long long WIDE_CHUNK = 262144;   // samples per pass over the layer stack
static int pad8(int n) { return (n + 7) & ~7; }
static size_t up256w(size_t n) { return (n + 255) & ~(size_t)255; }

int wide_supported(const nerfca_field_t& f) {
  NERFCA_REQUIRE(f.hidden % 8 == 0, NERFCA_E_UNSUPPORTED, "bf16 path: hidden must be a multiple of 8 (use precision fp32)");
  return NERFCA_OK;
}
// bf16 weight copies [layer][hidden, Kp_l] at the head of the workspace
static size_t wide_weight_bytes(const nerfca_field_t& f) {
  const int Dp = pad8(in_dim_of(f)), H = f.hidden, L = f.n_hidden + 1;
  return up256w(((size_t)H * Dp + (size_t)(L - 1) * H * H) * sizeof(bf16));
}
size_t wide_stash_bytes(const nerfca_field_t& f, long long P) {
  return (size_t)P * ((size_t)pad8(in_dim_of(f)) + (size_t)(f.n_hidden + 1) * f.hidden) * sizeof(bf16);
}
size_t wide_workspace_bytes(const nerfca_field_t& f, long long P, int backward) {
  const long long ch = P < WIDE_CHUNK ? P : WIDE_CHUNK;
  const size_t w = wide_weight_bytes(f);
  if (backward) return w + (size_t)ch * f.hidden * 2 * sizeof(bf16);
  return w + (size_t)ch * ((size_t)pad8(in_dim_of(f)) + 2 * (size_t)f.hidden) * sizeof(bf16);
}

static int wide_convert_weights(const nerfca_field_t& f, bf16* wb, cudaStream_t st) {
  const int D = in_dim_of(f), Dp = pad8(D), H = f.hidden, L = f.n_hidden + 1;
  WideWeights ws;
  for (int l = 0; l < NERFCA_MAX_LAYERS; ++l) ws.w[l] = f.weight[l];
  wide_cvt_weight_kernel<<<div_up((long long)H * Dp + (long long)(L - 1) * H * H, 256), 256, 0, st>>>(ws, H, D, Dp, L, wb);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}
static const bf16* wide_weight(const nerfca_field_t& f, const bf16* wb, int l) {
  const int Dp = pad8(in_dim_of(f)), H = f.hidden;
  return l == 0 ? wb : wb + (size_t)H * Dp + (size_t)(l - 1) * H * H;
}

int wide_field_forward(const nerfca_field_t& f, const nerfca_samples_t& s, float* raw_out, void* stash, void* workspace, cudaStream_t st) {
  const long long P = s.n_points;
  const int D = in_dim_of(f), Dp = pad8(D), H = f.hidden, L = f.n_hidden + 1;   // L layers with ReLU
  bf16* wb = (bf16*)workspace;
  int rc = wide_convert_weights(f, wb, st);
  if (rc) return rc;
  bf16* scratch = (bf16*)((uint8_t*)workspace + wide_weight_bytes(f));
  bf16* st_x0 = (bf16*)stash;
  bf16* st_h = stash ? st_x0 + (size_t)P * Dp : nullptr;
  const EncDesc enc = make_enc(f);
  for (long long c0 = 0; c0 < P; c0 += WIDE_CHUNK) {
    const long long np = (P - c0 < WIDE_CHUNK) ? P - c0 : WIDE_CHUNK;
    const long long ch = P < WIDE_CHUNK ? P : WIDE_CHUNK;
    bf16* x0 = stash ? st_x0 + (size_t)c0 * Dp : scratch;
    bf16* ping[2] = {scratch + (size_t)ch * Dp, scratch + (size_t)ch * Dp + (size_t)ch * H};
    wide_encode_kernel<<<div_up(np, 128), 128, 0, st>>>(make_src(s, c0, np), enc, Dp, x0);
    NERFCA_LAUNCH_OK();
    const bf16* in = x0;
    int Kp = Dp;
    for (int l = 0; l < L; ++l) {
      bf16* out = stash ? st_h + ((size_t)l * P + c0) * H : ping[l & 1];
      WideGemm g{};
      g.A = in; g.lda = Kp; g.a_rows = np; g.a_cols = Kp;
      g.B = wide_weight(f, wb, l); g.ldb = Kp; g.b_rows = H; g.b_cols = Kp;
      g.M = np; g.N = H; g.K = Kp;
      g.C = out; g.ldc = H; g.bias = f.bias[l];
      rc = run_wide_gemm<W_FWD>(g, st);
      if (rc) return rc;
      in = out; Kp = H;
    }
    wide_out_forward_kernel<<<div_up(np * 32, 256), 256, 0, st>>>(in, f.weight[L], f.bias[L], np, H, raw_out + c0);
    NERFCA_LAUNCH_OK();
  }
  return NERFCA_OK;
}

int wide_field_backward(const nerfca_field_t& f, const nerfca_samples_t& s, const float* d_raw, const void* stash, void* workspace,
                        const nerfca_field_grads_t& gr, cudaStream_t st) {
  const long long P = s.n_points;
  const int D = in_dim_of(f), Dp = pad8(D), H = f.hidden, L = f.n_hidden + 1;
  bf16* wb = (bf16*)workspace;
  int rc = wide_convert_weights(f, wb, st);     // (the forward's copies may live in a different workspace: the autograd path)
  if (rc) return rc;
  const bf16* st_x0 = (const bf16*)stash;
  const bf16* st_h = st_x0 + (size_t)P * Dp;
  const long long ch = P < WIDE_CHUNK ? P : WIDE_CHUNK;
  bf16* X = (bf16*)((uint8_t*)workspace + wide_weight_bytes(f));
  bf16* Y = X + (size_t)ch * H;
  for (long long c0 = 0; c0 < P; c0 += WIDE_CHUNK) {
    const long long np = (P - c0 < WIDE_CHUNK) ? P - c0 : WIDE_CHUNK;
    const unsigned rb = div_up(np, W_ROWS_PER_BLOCK);
    const bf16* h_last = st_h + ((size_t)(L - 1) * P + c0) * H;
    // (every kernel that produces a dZ also accumulates its column sums = the bias gradient of that layer)
    wide_out_backward_kernel<<<rb, 256, 0, st>>>(d_raw + c0, h_last, f.weight[L], np, H, X, gr.weight[L], gr.bias[L], gr.bias[L - 1]);
    NERFCA_LAUNCH_OK();
    for (int l = L - 1; l >= 0; --l) {
      const bf16* h_prev = (l > 0) ? st_h + ((size_t)(l - 1) * P + c0) * H : st_x0 + (size_t)c0 * Dp;
      const int Kin = (l > 0) ? H : D, Kinp = (l > 0) ? H : Dp;
      {  // wgrad: dW_l[n, k] += sum_p X[p, n] h_prev[p, k]
        WideGemm g{};
        g.A = X; g.lda = H; g.a_rows = np; g.a_cols = H;
        g.B = h_prev; g.ldb = Kinp; g.b_rows = np; g.b_cols = Kinp;
        g.M = H; g.N = Kin; g.K = np; g.k_split = 2048;
        g.C32 = gr.weight[l]; g.ldc32 = Kin;
        rc = run_wide_gemm<W_WGRAD>(g, st);
        if (rc) return rc;
      }
      if (l > 0) {  // dgrad: Y[p, k] = (sum_n X[p, n] W_l[n, k]) * 1[h_prev[p,k] > 0]
        WideGemm g{};
        g.A = X; g.lda = H; g.a_rows = np; g.a_cols = H;
        g.B = wide_weight(f, wb, l); g.ldb = H; g.b_rows = H; g.b_cols = H;
        g.M = np; g.N = H; g.K = H;
        g.C = Y; g.ldc = H; g.mask = h_prev; g.ldm = H; g.colsum = gr.bias[l - 1];
        rc = run_wide_gemm<W_DGRAD>(g, st);
        if (rc) return rc;
        bf16* tmp = X; X = Y; Y = tmp;
      } else if (f.n_latent > 0 && gr.latents) {
        const int T = f.n_latent;
        NERFCA_REQUIRE(T <= WL_MAX_T, NERFCA_E_UNSUPPORTED, "bf16 layer-wise path: more than 16 latent dims (use precision fp32)");
        const size_t w_bytes = (size_t)((H + 255) & ~255) * T * sizeof(float), acc_bytes = (size_t)f.n_phases * T * sizeof(float);
        const int use_smem = w_bytes + acc_bytes <= 40 * 1024;
        NERFCA_REQUIRE(w_bytes <= 40 * 1024, NERFCA_E_UNSUPPORTED, "bf16 layer-wise path: hidden x latent dims too large for the latent-gradient kernel");
        const size_t smem = w_bytes + (use_smem ? acc_bytes : 0);
        wide_latent_grad_kernel<<<div_up(np, 8 * WL_ROWS_PER_WARP), 256, smem, st>>>(X, f.weight[0], make_src(s, c0, np), H, D, enc_dim_of(f), T,
                                                                                  f.n_phases, use_smem, gr.latents);
        NERFCA_LAUNCH_OK();
      }
    }
  }
  return NERFCA_OK;
}

}  // namespace nerfca
