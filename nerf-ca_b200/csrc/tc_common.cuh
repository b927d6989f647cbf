// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, bulk async copy (TMA 1-D),
// tcgen05 MMA / TMEM load-store / commit / fences, and the UMMA shared-memory + instruction descriptors.
//
// Shared-memory operand layout used everywhere ("tile-canonical"): a [128 rows x K] bf16 tile is stored as
//     byte(row, k) = (k / 8) * 2048 + row * 16 + (k % 8) * 2
// i.e. K is cut into 8-element (16-byte) chunks and each chunk holds the 128 rows contiguously.  In UMMA terms this
// is the no-swizzle ("interleave") canonical layout made of 8x16B core matrices, and the SAME bytes can be read
//   * K-major   (rows = M/N index, k = reduction): core matrices 128 B apart along M/N, 2048 B apart along K;
//   * MN-major  (k    = M/N index, rows = reduction): core matrices 2048 B apart along M/N, 128 B apart along K,
// which is what lets one stored tile serve forward (K-major), dgrad (weights MN-major) and wgrad (both MN-major).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace nerfca {
namespace tc {

constexpr int TILE_M = 128;        // samples per tile == UMMA M == TMEM lanes
constexpr int CHUNK_BYTES = 2048;  // one 8-wide K chunk of a 128-row tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (A suspend-time hint as the fourth try_wait operand was measured 2-5 % SLOWER on the backward kernels than the plain form:
// r3j, 390 vs 396 us; the plain try_wait already sleeps in hardware for a short system-defined time.)
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (surfacing as a launch failure) instead of hanging the GPU.  No printf here: a
// call in the wait path would force every live accumulator register onto the stack around each wait.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
#ifdef NERFCA_TIMELINE_BUILD   // developer build only: say which wait gave up
      printf("mbar_wait timeout: block %d warp %d lane %d barrier smem+0x%x parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5),
             (int)(threadIdx.x & 31), bar, parity);
#endif
      __trap();
    }
  }
}

// one lane of a converged warp (elect.sync)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// 16-byte shared-memory load by 32-bit shared-space address (no generic -> shared conversion at the use site)
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// makes a value opaque to the optimizer so that it lives in a register instead of being recomputed at every use
__device__ __forceinline__ void pin(uint32_t& x) { asm volatile("" : "+r"(x)); }

// per-warpgroup register budget (all four warps of a warpgroup execute it): the kernel launches with the launch-bound count,
// light roles hand registers back and heavy roles take them
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ---- async proxy ------------------------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies reading smem)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, 16-B aligned addresses)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// L2 prefetch of a global range (no completion tracking): the later bulk copy of the same bytes then hits L2
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// 1-D bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]   (kind::f16: bf16 operands, fp32 accumulate), issued by ONE thread
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior tcgen05 async ops of this thread -> one arrival on the mbarrier
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes (lane field of taddr = 32 * (warp % 4)) x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
        "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&v)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr) : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 32 consecutive columns (used to preload an accumulator with the bias)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                         uint32_t a6, uint32_t a7) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a0), "r"(a1), "r"(a2),
               "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------------------
// UMMA shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64).  LBO = byte distance between core matrices adjacent
// along K, SBO = along M/N (both for K-major and MN-major operands in the no-swizzle canonical layouts).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// K-major view of a tile-canonical buffer, reduction step kk (16 elements = 2 chunks)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int kk) { return smem_desc(base + kk * 2 * CHUNK_BYTES, CHUNK_BYTES, 128); }
// MN-major view: M/N index = stored column (chunk direction), reduction = stored rows; step kk covers rows [16kk, 16kk+16)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int kk) { return smem_desc(base + kk * 2 * 128, 128, CHUNK_BYTES); }

// ---- lean MMA issue ---------------------------------------------------------------------------------------------
// Issuing a tcgen05.mma costs ~50 cycles even when nothing else is in the way (measured, tools/micro/mma_rate.cu), so the
// issue loop must not rebuild 64-bit descriptors per K step: the descriptor is split into (lo, hi) once per operand and
// every K step only adds a compile-time constant to the low word (the start-address field, in 16-byte units).
struct Desc { uint32_t lo, hi; };
__device__ __forceinline__ Desc desc_split(uint64_t d) { return Desc{(uint32_t)d, (uint32_t)(d >> 32)}; }
constexpr uint32_t KSTEP_KMAJOR = (2 * CHUNK_BYTES) >> 4;   // 16 reduction elements = 2 chunks
constexpr uint32_t KSTEP_MNMAJOR = (2 * 128) >> 4;          // 16 reduction rows of 16 B inside a chunk
__device__ __forceinline__ Desc kmajor(uint32_t base) { return desc_split(smem_desc(base, CHUNK_BYTES, 128)); }
__device__ __forceinline__ Desc mnmajor(uint32_t base) { return desc_split(smem_desc(base, 128, CHUNK_BYTES)); }
__device__ __forceinline__ Desc kmajor_rows(uint32_t base, uint32_t chunk_bytes) { return desc_split(smem_desc(base, chunk_bytes, 128)); }

__device__ __forceinline__ void umma_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D (+)= A * B over KSTEPS reduction steps of 16; the first step overwrites D unless acc_first != 0
template <int KSTEPS, uint32_t A_STEP, uint32_t B_STEP>
__device__ __forceinline__ void umma_k(uint32_t d_tmem, Desc a, Desc b, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int kk = 0; kk < KSTEPS; ++kk) umma_lh(d_tmem, a.lo + kk * A_STEP, a.hi, b.lo + kk * B_STEP, b.hi, idesc, kk > 0 ? 1u : acc_first);
}

// A from tensor memory (TS form): the [128 x K] bf16 A operand sits in TMEM with lane = row and two K-consecutive elements per
// 32-bit column (cute UMMA::tmem_frg, M = 128), so one K step of 16 is 8 columns.  A must be K-major in this form.
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
constexpr uint32_t KSTEP_TMEM = 8;   // 16 bf16 reduction elements = 8 TMEM columns
template <int KSTEPS, uint32_t B_STEP>
__device__ __forceinline__ void umma_ts_k(uint32_t d_tmem, uint32_t a_tmem, Desc b, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int kk = 0; kk < KSTEPS; ++kk) umma_ts(d_tmem, a_tmem + kk * KSTEP_TMEM, b.lo + kk * B_STEP, b.hi, idesc, kk > 0 ? 1u : acc_first);
}

// UMMA instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D fp32, A/B bf16
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// [rows x K] bf16 tile with FEWER than 128 rows stored chunk-major with `chunk_bytes` = rows * 16 per 8-wide K chunk
// (used for the 16-row output-weight tile): K-major view, reduction step kk.
__device__ __forceinline__ uint64_t desc_kmajor_rows(uint32_t base, int kk, uint32_t chunk_bytes) {
  return smem_desc(base + kk * 2 * chunk_bytes, chunk_bytes, 128);
}

// {lo, hi} -> bf16x2 with ReLU fused into the conversion (one F2FP instruction)
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// per-half (h > 0 ? 1.0 : 0.0) as bf16x2, and packed bf16x2 multiply: gradient masking by the ReLU pattern of h
__device__ __forceinline__ uint32_t relu_mask_bf16x2(uint32_t h) {
  uint32_t d;
  asm("set.gt.bf16x2.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(h), "r"(0u));
  return d;
}
__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// packed fp32 pair add (FADD2): {a.x + b.x, a.y + b.y}
__device__ __forceinline__ float2 add_f32x2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace tc
}  // namespace nerfca
