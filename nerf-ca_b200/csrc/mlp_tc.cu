// bf16 tensor-core path of the two coordinate MLPs on tcgen05 / TMEM (sm_100a): forward + two-pass backward.
//
// Shapes: hidden = 128, 4 hidden layers (5 ReLU layers L0..L4 + the 128 -> 1 output layer), first-layer input padded to
// kpad0 (multiple of 16, with one extra constant-1 column that carries the layer-0 bias).  One tile = 128 consecutive
// samples (= UMMA M = TMEM lanes).  All kernels are persistent: grid = #SMs, CTA b serves net (b % n_nets) and walks
// the tiles  worker, worker + n_workers, ...   Both fields (static CPPN, dynamic Temporal) run in the SAME launch.
//
// What lives where
//   shared memory : the bf16 weights the kernel needs (tile-canonical K-major bytes, read K-major by the forward GEMMs
//                   and MN-major by the dgrad GEMMs), activation / gradient tiles of the tile in flight
//   TMEM          : fp32 accumulators; the forward also chains its activations through TMEM (TS-form MMA); in the backward the
//                   weight-gradient accumulators stay resident for the CTA's whole lifetime and are flushed once with vector
//                   reductions
//   HBM           : per tile and net the forward stashes H0 .. H3 (bf16, 32 KB each, written straight from the epilogue registers)
//                   and the ReLU pattern of H4 as one bit per element (2 KB): 130 KB.  The backward recomputes nothing but X0.
//   L2            : the hand-off of dZ2 from the top backward role to the bottom role (a ring of 32 KB slots + flags)
//
// Kernels
//   tc_forward_kernel   X0 -> H0 .. H4 -> raw.   16 epilogue warps (2 tiles in flight x (row quadrant, column half)), each slot
//                       issues its own MMAs, 4 X0 producer warps.  Hidden-layer biases are added in the epilogue, ReLU is fused
//                       into the bf16 conversion, the 128 -> 1 layer is an N = 16 MMA against a (hi, lo) bf16 split of w_out.
//   tc_bwd_kernel       ONE launch, two CTA roles per net (31 : 43 of the 74 CTAs); the default while the latent gradient comes through one-hot columns (tc_bwd2_kernel, mlp_tc_bwd2.cuh, otherwise):
//     top role          layers 4, 3 (+ output layer): loads H2, H3, the H4 pattern, d_raw; dZ4' = d_raw 1[H4 > 0]; dgrad 4, 3;
//                       wgrad 4, 3 with the bias gradient folded in (N = 144) accumulate in TMEM; dZ2 -> L2 ring.
//     bottom role       layers 2, 1, 0: polls the ring, loads dZ2, H1, H0; dgrad 2, 1; wgrad 2, 1, 0; rebuilds X0; latent
//                       gradients through one-hot phase columns of X0 (or an explicit latent dgrad for > 12 phases).
//     Each role: 8 epilogue warps, one MMA-issuing warp inside the epilogue's named barriers, one load warp (bulk copies),
//     top: a publisher warp for the GPU-scope flags, bottom: 4 X0 producer warps.  NERFCA_BWD_MERGED=0 runs the roles as
//     two launches (tc_bwd_top_kernel / tc_bwd_bot_kernel, hand-off through HBM).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "pack.cuh"
#include "tc_common.cuh"

namespace nerfca {

using namespace tc;

constexpr int TC_H = 128;
constexpr uint32_t TILE_BYTES = 32768;           // one [128 x 128] bf16 tile
constexpr int TC_N_RELU = 5;
constexpr int STASH_TILES = 4;                   // H0 .. H3 of every tile (bf16, tile-canonical bytes)
constexpr uint32_t MASK_BYTES = 2048;            // + the ReLU pattern of H4, one BIT per element: row r owns 16 bytes = 4 words, word 2 ch + g covers
                                                 // columns [64 ch + 32 g, + 32): bit i = column 2i, bit 16 + i = column 2i + 1 (i < 16)
constexpr size_t STASH_STRIDE = (size_t)STASH_TILES * 32768 + MASK_BYTES;   // bytes per tile and net
constexpr size_t STASH_PATTERN4_OFF = (size_t)STASH_TILES * 32768;          // the H4 pattern sits behind the four tiles
// (Two ways to spare the backward's epilogue its shared-memory reads of the activation tiles for the ReLU masks -- they compete with
// the MMAs' operand fetches and take ~2 800 cycles per step in the two-tiles-in-flight kernel -- were measured and dropped:
// stashing the 1-bit pattern of EVERY layer made the top backward role 31 % faster per tile but cost the forward's epilogue +42 us
// for the bit packing (r3o); deriving the patterns in the backward from the stash in global memory doubled its DRAM reads, 470 us
// against 395 (r3r).)
constexpr int FWD_THREADS = 20 * 32;             // 16 epilogue / issue warps + 4 X0 producer warps
constexpr int FAST_FREQ = 12;                    // band count the register-resident encoder is specialised for

// shapes the fused kernels are built for (the layer-wise tcgen05 path of mlp_wide.cu serves every other bf16 request)
bool tc_shape_ok(const nerfca_field_t& f, bool training) {
  if (!(f.hidden == TC_H && f.n_hidden == TC_N_RELU - 1 && in_dim_of(f) + 1 <= 96 && f.n_latent <= 16)) return false;
  // the backward keeps the [phase][t] latent-gradient table in shared memory (forward-only calls, e.g. query_time's per-point latents, do not)
  return !training || f.n_latent == 0 || (size_t)f.n_phases * f.n_latent <= 256;
}
int tc_supported(const nerfca_field_t& f) {
  NERFCA_REQUIRE(f.hidden == TC_H, NERFCA_E_UNSUPPORTED, "the tcgen05 path is built for hidden == 128 (use precision fp32)");
  NERFCA_REQUIRE(f.n_hidden == TC_N_RELU - 1, NERFCA_E_UNSUPPORTED, "the tcgen05 path is built for 4 hidden layers (use precision fp32)");
  NERFCA_REQUIRE(in_dim_of(f) + 1 <= 96, NERFCA_E_UNSUPPORTED, "tcgen05 path: first-layer input wider than 95 (use precision fp32)");
  NERFCA_REQUIRE(f.n_latent <= 16, NERFCA_E_UNSUPPORTED, "tcgen05 path: more than 16 latent dims (use precision fp32)");
  return NERFCA_OK;
}

// ---- parameter packing ---------------------------------------------------------------------------------------------
struct PackNet {
  const float* w[NERFCA_MAX_LAYERS];
  const float* b[NERFCA_MAX_LAYERS];
  uint8_t* out;
  int in_dim, kpad0;
  uint32_t wout_off, f32_off;
};
struct PackArgs { PackNet net[2]; };

__global__ void pack_params_kernel(PackArgs pa) {
  const PackNet& a = pa.net[blockIdx.y];
  const int n_w0 = a.kpad0 * 128;
  const int total_w = n_w0 + 4 * 128 * 128;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < total_w) {
    int l, e;
    if (tid < n_w0) { l = 0; e = tid; }
    else { l = 1 + (tid - n_w0) / (128 * 128); e = (tid - n_w0) % (128 * 128); }
    const int K = (l == 0) ? a.in_dim : 128;
    // e enumerates the packed layout: chunk-major, then output row n, then position in the chunk
    const int chunk = e / (128 * 8), n = (e / 8) % 128, kk = e % 8;
    const int k = chunk * 8 + kk;
    float v = 0.f;
    if (k < K) v = __ldg(a.w[l] + (size_t)n * K + k);
    else if (l == 0 && k == K) v = a.b[0] ? __ldg(a.b[0] + n) : 0.f;   // bias column of layer 0
    const size_t base = (l == 0) ? 0 : (size_t)a.kpad0 * 256 + (size_t)(l - 1) * TILE_BYTES;
    reinterpret_cast<__nv_bfloat16*>(a.out + base)[e] = __float2bfloat16_rn(v);
  }
  if (tid < 16 * 128) {   // output-layer tile, 16 rows per 8-wide K chunk: byte(n, k) = (k / 8) * 256 + n * 16 + (k % 8) * 2
    const int chunk = tid / (16 * 8), n = (tid / 8) % 16, kk = tid % 8;
    const float wv = __ldg(a.w[TC_N_RELU] + chunk * 8 + kk);
    const __nv_bfloat16 hi = __float2bfloat16_rn(wv);
    const __nv_bfloat16 v = (n == 0) ? hi : ((n == 1) ? __float2bfloat16_rn(wv - __bfloat162float(hi)) : __float2bfloat16_rn(0.f));
    reinterpret_cast<__nv_bfloat16*>(a.out + a.wout_off)[tid] = v;
  }
  float* fb = reinterpret_cast<float*>(a.out + a.f32_off);
  const int nb = (int)f32_block_floats();
  if (tid < nb) {
    float v = 0.f;
    if (tid < TC_N_RELU * 128) { const int l = tid / 128; v = a.b[l] ? __ldg(a.b[l] + tid % 128) : 0.f; }
    else if (tid < TC_N_RELU * 128 + 128) v = __ldg(a.w[TC_N_RELU] + (tid - TC_N_RELU * 128));
    else if (tid == TC_N_RELU * 128 + 128) v = a.b[TC_N_RELU] ? __ldg(a.b[TC_N_RELU]) : 0.f;
    fb[tid] = v;
  }
}

static int pack_params(const nerfca_field_t* const* f, int n_nets, void* dst, cudaStream_t st) {
  PackArgs pa;
  size_t off = 0;
  int max_total = 0;
  for (int i = 0; i < n_nets; ++i) {
    const NetDims d = net_dims(*f[i]);
    PackNet& a = pa.net[i];
    for (int l = 0; l < NERFCA_MAX_LAYERS; ++l) { a.w[l] = f[i]->weight[l]; a.b[l] = f[i]->bias[l]; }
    a.out = (uint8_t*)dst + off;
    a.in_dim = d.in_dim; a.kpad0 = d.kpad0; a.wout_off = d.wout_off; a.f32_off = d.f32_off;
    off += pack_stride(d);
    const int total = d.kpad0 * 128 + 4 * 128 * 128;
    max_total = total > max_total ? total : max_total;
  }
  ProfScope prof(NERFCA_K_PACK, st);
  pack_params_kernel<<<dim3(div_up(max_total, 256), n_nets), 256, 0, st>>>(pa);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

// Table for the optimizer kernels (pack.cuh): where each tensor of `fields` sits inside the flat parameter buffer and where its
// packed copy goes.  Fields whose tensors are not all inside the buffer, or not of the packed shape, give an error.
int make_repack_table(const nerfca_field_t* const* fields, int n_nets, void* workspace, const float* params, long long n,
                      RepackTable* out) {
  out->n_segs = 0;
  NERFCA_REQUIRE(n_nets >= 1 && n_nets <= 2 && workspace && params, NERFCA_E_ARG, "bad repack arguments");
  size_t off = 0;
  int ns = 0;
  for (int i = 0; i < n_nets; ++i) {
    const nerfca_field_t& f = *fields[i];
    if (!tc_shape_ok(f, true)) { out->n_segs = 0; return NERFCA_OK; }    // layer-wise path: converts its weights itself, nothing to re-pack
    const NetDims d = net_dims(f);
    RepackNet& rn = out->net[i];
    rn.out = (uint8_t*)workspace + off;
    rn.in_dim = d.in_dim; rn.kpad0 = d.kpad0; rn.w0_bytes = d.w0_bytes; rn.wout_off = d.wout_off; rn.f32_off = d.f32_off;
    off += pack_stride(d);
    for (int l = 0; l <= TC_N_RELU; ++l) {
      for (int is_bias = 0; is_bias < 2; ++is_bias) {
        const float* t = is_bias ? f.bias[l] : f.weight[l];
        if (!t) continue;
        const int K = (l == 0) ? d.in_dim : 128;
        const int count = is_bias ? (l == TC_N_RELU ? 1 : 128) : (l == TC_N_RELU ? 128 : 128 * K);
        NERFCA_REQUIRE(t >= params && t + count <= params + n, NERFCA_E_ARG, "a field tensor lies outside the flat parameter buffer");
        NERFCA_REQUIRE(((t - params) & 3) == 0, NERFCA_E_ARG, "field tensors must start on 4-float boundaries of the flat buffer");
        RepackSeg& sg = out->seg[ns++];
        sg.begin = (long long)(t - params);
        sg.count = count;
        sg.kind = (short)(l == TC_N_RELU ? (is_bias ? SEG_BOUT : SEG_WOUT) : (is_bias ? SEG_BIAS : SEG_WEIGHT));
        sg.layer = (short)l; sg.K = (short)K; sg.net = (short)i;
      }
    }
  }
  out->n_segs = ns;
  return NERFCA_OK;
}

// ---- first-layer input tile X0 (bf16, tile-canonical) ------------------------------------------------------------
// Each tile row is built by the two threads (column halves ch = 0 / 1) that own it.  Fast path (BANDS, 12 bands):
// the sines / cosines of bands 0-5 and 6-11 come from one range-reduced MUFU sin/cos pair per coordinate (at band 0 resp.
// band 6, sincos_reduced) followed by double-angle steps, everything in registers, 16-byte stores.  Against the fp32
// reference expression the features differ by <= 2e-5 (ours) + 2.4e-4 (the reference's own "+ fl32(pi/2)" argument
// rounding at band 11, which the double-angle cosine does not reproduce) before the bf16 rounding applied to the tile
// (half-ulp 2e-3); tests/test_gpu_parity.py::test_tensor_core_x0_tile bounds it.  Other encodings take the generic path.
struct X0Desc {
  EncDesc enc;
  int kpad0;
  int fast;     // BANDS with n_freq == FAST_FREQ
  int onehot;   // > 0: columns in_dim + 1 + ph (ph < onehot) carry 1[phase == ph] (bottom backward pass: the layer-0 weight-
                // gradient GEMM then also yields the per-phase column sums of dZ0, from which the latent gradient follows)
};

__device__ __forceinline__ uint4 pack_chunk(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
// Where a finished 8-feature chunk of a tile row goes.
//   SmemSink: the tile-canonical shared-memory tile (B operand of the layer-0 weight-gradient GEMM, bottom backward pass)
//   TmemSink: the A operand of the layer-0 forward GEMM in tensor memory: lane = tile row, 32-bit column c holds the
//             features (2c, 2c+1), so chunk c is columns [4c, 4c+4) of this thread's lane (tcgen05.st, warp-collective)
struct SmemSink {
  uint8_t* tile;
  int row;
  __device__ __forceinline__ void chunk(int c, const float* v) const { *reinterpret_cast<uint4*>(tile + c * CHUNK_BYTES + row * 16) = pack_chunk(v); }
};
struct TmemSink {
  uint32_t taddr;   // lane quadrant base | first column of the A region
  __device__ __forceinline__ void chunk(int c, const float* v) const {
    const uint4 w = pack_chunk(v);
    tmem_st4(taddr + 4 * c, w.x, w.y, w.z, w.w);
  }
};

// The raw inputs of one tile row: fetched early (the loads stay in flight until first use), consumed by emit_x0_row.
struct RowIn {
  float x, y, z;
  int phase;
  bool valid;
};
__device__ __forceinline__ RowIn fetch_row(const X0Desc& xd, const SampleSrc& src, long long p, bool valid) {
  RowIn r;
  r.x = r.y = r.z = 0.f;
  r.phase = 0;
  r.valid = valid;
  if (valid) {
    load_point(src, p, r.x, r.y, r.z);
    if (xd.enc.n_latent > 0) r.phase = load_phase(src, p);
  }
  return r;
}

// sin / cos of an argument of a few hundred radians (2^6 |x|): two-term Cody-Waite reduction to [-pi, pi] (the power-of-two scaling
// of the argument is exact), then the MUFU approximations, whose absolute error inside that range is ~5e-7.  The five double-angle
// steps that follow double the error each: <= 2e-5 at the highest band, against a bf16 half-ulp of 2e-3 and the reference's own
// "+ fl32(pi/2)" argument rounding of 2.4e-4 there.
__device__ __forceinline__ void sincos_reduced(float a, float& s, float& c) {
  const float k = rintf(a * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, a);          // 2 pi = 6.2831854820251465 (fp32) - 1.7484555e-7
  r = fmaf(k, 1.7484555e-7f, r);
  __sincosf(r, &s, &c);
}

// band_w / lat_tab: per-band weights [n_freq] (or null) and the latent table [n_phases, n_latent]; either global or a
// shared-memory copy.  Every branch below is warp-uniform (ch, the encoding and kpad0 are), which the TMEM sink needs.
template <class Sink>
__device__ __forceinline__ void emit_x0_row(const X0Desc& xd, const RowIn& in, const float* band_w, const float* lat_tab, const Sink& sink,
                                            int ch) {
  const float x = in.x, y = in.y, z = in.z;
  const bool valid = in.valid;
  const EncDesc& e = xd.enc;
  int phase = in.phase;
  if (phase < 0 || phase >= e.n_phases) phase = 0;
  const float live = valid ? 1.f : 0.f;
  if (xd.fast) {
    constexpr int HB = FAST_FREQ / 2;            // bands per half
    constexpr int SPLIT = 40;                    // features [0, 40) belong to ch 0 (5 chunks), the rest to ch 1
    const float scale = ch ? (float)(1 << HB) : 1.f;
    float s[3], c[3];
    sincos_reduced(x * scale, s[0], c[0]);
    sincos_reduced(y * scale, s[1], c[1]);
    sincos_reduced(z * scale, s[2], c[2]);
    // features are produced in index order and leave in 16-byte chunks as soon as 8 of them exist; `cnt` is a
    // compile-time constant at every use once the loops are unrolled, so `buf` stays in registers
    float buf[8];
    int cnt = 0;
    const int n_chunks = xd.kpad0 / 8;
    const int c_base = ch ? SPLIT / 8 : 0;
#define NERFCA_PUSH(val)                                                                  \
    do {                                                                                    \
      buf[cnt & 7] = (val);                                                                 \
      ++cnt;                                                                                \
      if ((cnt & 7) == 0 && c_base + (cnt >> 3) - 1 < n_chunks) sink.chunk(c_base + (cnt >> 3) - 1, buf); \
    } while (0)
    if (ch == 0) {
      NERFCA_PUSH(x); NERFCA_PUSH(y); NERFCA_PUSH(z);
#pragma unroll
      for (int b = 0; b < HB; ++b) {
        const float w = (band_w ? band_w[b] : 1.f) * live;
#pragma unroll
        for (int k = 0; k < 3; ++k) NERFCA_PUSH(w * s[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) NERFCA_PUSH(w * c[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float t = 2.f * s[k] * c[k];
          c[k] = fmaf(-2.f * s[k], s[k], 1.f);
          s[k] = t;
        }
      }
      NERFCA_PUSH((band_w ? band_w[HB] : 1.f) * live * s[0]);   // feature 39 = sin(2^6 x) closes chunk 4
    } else {
#pragma unroll
      for (int b = 0; b < HB; ++b) {
        const float w = (band_w ? band_w[HB + b] : 1.f) * live;
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (b > 0 || k > 0) NERFCA_PUSH(w * s[k]);                                   // feature 39 belongs to ch 0
#pragma unroll
        for (int k = 0; k < 3; ++k) NERFCA_PUSH(w * c[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float t = 2.f * s[k] * c[k];
          c[k] = fmaf(-2.f * s[k], s[k], 1.f);
          s[k] = t;
        }
      }
      // latents, the constant-1 column (layer-0 bias / bias gradient), zero padding up to kpad0 (<= 96)
#pragma unroll
      for (int t = 0; t < 96 - (3 + 6 * FAST_FREQ); ++t) {
        float val = 0.f;
        if (t < e.n_latent) val = live * lat_tab[phase * e.n_latent + t];
        else if (t == e.n_latent) val = live;
        else if (t - e.n_latent - 1 < xd.onehot) val = (t - e.n_latent - 1 == phase) ? live : 0.f;
        NERFCA_PUSH(val);
      }
    }
#undef NERFCA_PUSH
    return;
  }
  // generic encodings: ch 0 writes chunks [0, kpad0/16), ch 1 the rest
  const int c_lo = ch ? xd.kpad0 / 16 : 0, c_hi = ch ? xd.kpad0 / 8 : xd.kpad0 / 16;
  for (int c = c_lo; c < c_hi; ++c) {
    float buf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int f = c * 8 + j;
      float val = 0.f;
      if (valid) {
        if (f < e.in_dim) val = enc_feature(e, f, x, y, z, phase);
        else if (f == e.in_dim) val = 1.f;
        else if (f - e.in_dim - 1 < xd.onehot) val = (f - e.in_dim - 1 == phase) ? 1.f : 0.f;
      }
      buf[j] = val;
    }
    sink.chunk(c, buf);
  }
}

// ---- shared-memory access by 32-bit shared-space address ------------------------------------------------------
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// ---- accumulator (this thread's lane, 64 columns) -> registers ---------------------------------------------------
__device__ __forceinline__ void ld_acc64(uint32_t taddr, uint32_t (&a)[32], uint32_t (&b)[32]) {
  tmem_ld32(taddr, a);
  tmem_ld32(taddr + 32, b);
  tmem_ld_wait();
}

// dZ = acc * 1[h > 0] for this thread's 64 columns: h comes from the bf16 activation tile in shared memory, the result
// goes to `out` (tile-canonical bf16).
__device__ __forceinline__ void masked_grad_store(const uint32_t (&va)[32], const uint32_t (&vb)[32], const uint8_t* h_tile, uint8_t* out,
                                                  int row, int ch) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = ch * 8 + half * 4 + j;
      const uint4 hv = *reinterpret_cast<const uint4*>(h_tile + c * CHUNK_BYTES + row * 16);
      const uint32_t* v = half ? vb : va;
      uint4 o;
      o.x = mul_bf16x2(pack_bf16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])), relu_mask_bf16x2(hv.x));
      o.y = mul_bf16x2(pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])), relu_mask_bf16x2(hv.y));
      o.z = mul_bf16x2(pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])), relu_mask_bf16x2(hv.z));
      o.w = mul_bf16x2(pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])), relu_mask_bf16x2(hv.w));
      *reinterpret_cast<uint4*>(out + c * CHUNK_BYTES + row * 16) = o;
    }
  }
}
// H = relu(acc + bias) for this thread's 64 columns -> bf16 tile
__device__ __forceinline__ void relu_bias_store(const uint32_t (&va)[32], const uint32_t (&vb)[32], const float* bias, uint8_t* out,
                                                int row, int ch) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = ch * 8 + half * 4 + j;
      const uint32_t* v = half ? vb : va;
      const float4 b0 = *reinterpret_cast<const float4*>(bias + c * 8), b1 = *reinterpret_cast<const float4*>(bias + c * 8 + 4);
      uint4 o;
      o.x = pack_relu_bf16x2(__uint_as_float(v[8 * j]) + b0.x, __uint_as_float(v[8 * j + 1]) + b0.y);
      o.y = pack_relu_bf16x2(__uint_as_float(v[8 * j + 2]) + b0.z, __uint_as_float(v[8 * j + 3]) + b0.w);
      o.z = pack_relu_bf16x2(__uint_as_float(v[8 * j + 4]) + b1.x, __uint_as_float(v[8 * j + 5]) + b1.y);
      o.w = pack_relu_bf16x2(__uint_as_float(v[8 * j + 6]) + b1.z, __uint_as_float(v[8 * j + 7]) + b1.w);
      *reinterpret_cast<uint4*>(out + c * CHUNK_BYTES + row * 16) = o;
    }
  }
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// =====================================================================================================================
// forward
// =====================================================================================================================
struct FwdNet {
  const uint8_t* pack;
  float* raw_out;          // per-sample field output, or null when only the ray sums are wanted
  float* ray_sum;          // optional [n_rays]: += act(raw) * 1e-2 * delta per sample (line integral fused into the output layer's epilogue)
  int act;
  uint8_t* stash;          // [n_tiles][STASH_TILES][TILE_BYTES] or null
  X0Desc x0;
  uint32_t w0_bytes, wout_off, f32_off, pack_bytes;
};
struct FwdArgs {
  SampleSrc src;
  FwdNet net[2];
  int n_nets;
  long long n_tiles;
  double* zero_terms;   // optional: NERFCA_N_LOSS_TERMS loss sums to clear (they are accumulated by the loss kernel that follows)
  long long* dbg;   // optional event timeline of one CTA (NERFCA_TIMELINE=fwd, NERFCA_TIMELINE_CTA=n)
  int dbg_cta;
};
// Each logging thread owns a region of 1000 (tag, clock) pairs selected by tag / 1000 (1: slot-0 epilogue, 2: slot-1
// epilogue, 3: MMA thread) and keeps its own count: no atomics, the stores are fire-and-forget.
#ifdef NERFCA_TIMELINE_BUILD   // make EXTRA=-DNERFCA_TIMELINE_BUILD: the logging costs registers and issue slots, so it is compiled out by default
#define NERFCA_TL(cond, tag)                                                              \
  do {                                                                                    \
    if (a.dbg && blockIdx.x == (unsigned)a.dbg_cta && (cond) && tl_n < 1000) {                              \
      long long* r__ = a.dbg + (size_t)((tag) / 1000) * 2000 + 2 * tl_n;                  \
      r__[0] = (tag); r__[1] = clock64();                                                 \
      ++tl_n;                                                                             \
    }                                                                                     \
  } while (0)
#else
#define NERFCA_TL(cond, tag) do { } while (0)
#endif

// TMEM map of the forward kernel (512 columns): two tile slots s = 0, 1
//   ACC  [128 s, 128 s + 128)       fp32 accumulator of the layer in flight
//   A    [256 + 64 s, .. + 64)      bf16 activations H_l = A operand of layer l + 1
//   X0   [384 + 48 s, .. + 48)      bf16 encoded input = A operand of layer 0, written one tile ahead by the producer warps
//   OUT  [480 + 16 s, .. + 16)      accumulator of the 128 -> 1 layer (columns 0 / 1: hi / lo part of w_out)
constexpr uint32_t FWD_ACC = 0, FWD_A = 256, FWD_X0 = 384, FWD_OUT = 480;
constexpr int FWD_X0_WARP0 = 16;   // warps 0-15: the two slots' epilogue warps, 16-19: X0 producers
constexpr uint32_t WOUT_KSTEP = (2 * 256) >> 4;

__device__ __forceinline__ void named_bar_sync(int id, int n_threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory"); }

// relu(acc [+ bias]) of 32 accumulator columns -> 16 packed bf16x2 words (word i = columns 2i, 2i+1)
template <bool BIAS>
__device__ __forceinline__ void relu_pack32(const uint32_t (&v)[32], const float4 (&b)[8], uint32_t (&w)[16]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float2 p0 = make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]));
    float2 p1 = make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    if (BIAS) {
      p0 = add_f32x2(p0, make_float2(b[j].x, b[j].y));
      p1 = add_f32x2(p1, make_float2(b[j].z, b[j].w));
    }
    w[2 * j] = pack_relu_bf16x2(p0.x, p0.y);
    w[2 * j + 1] = pack_relu_bf16x2(p1.x, p1.y);
  }
}

// one arrival per warp: every lane's preceding tcgen05 work is complete and fenced, the warp converges, lane 0 arrives
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// The forward keeps the whole activation chain of a tile in tensor memory: the epilogue threads read the fp32 accumulator
// (tcgen05.ld), apply bias + ReLU, and write the bf16 result back to TMEM (tcgen05.st) where the next layer's MMA takes it as
// its A operand (TS form).  Shared memory only carries the weights (B operands: 32 KB read per tile and layer instead of 64 KB
// read + 32 KB written), and H0 / H2 go to the HBM stash straight from registers.
//   * Two tiles (slots) are in flight per CTA, each served by 8 epilogue warps (TMEM lane quadrant x column half).  A slot is
//     self-contained: when its warps have written H_l they meet at a named barrier and one elected lane of the slot's first
//     warp issues layer l + 1 right there -- no hand-off to a separate MMA warp, one mbarrier round trip less per layer.
//   * The encoded input X0 of a tile is produced by four dedicated warps (one thread per sample: ray point in fp64, sines by
//     MUFU + double-angle steps) into its own TMEM region while the slot's previous tile is still in the layer chain.
//   * The 128 -> 1 output layer is a sixth MMA (N = 16) against a (hi, lo) bf16 split of w_out into its own accumulator, so
//     layer 0 of the slot's next tile is issued right behind it.
//   * The stash (130 KB per tile) leaves as 16-byte streaming stores straight from the epilogue registers.  That path tops out near
//     16 B/clk per SM (~4.7 TB/s over the chip) and is what bounds the training forward (the same kernel without a stash runs 3x
//     faster); routing it through shared-memory staging + bulk stores was measured slower both ways it was tried (per-warp 512-byte
//     bulk stores: 294 us, one 32 KB bulk store per slot and layer: 301 us, against 227 us): the staging traffic competes with the
//     MMAs' weight reads for shared-memory bandwidth.
// smem: [packed block][barriers][band weights 32 f32][latent table 256 f32]
__global__ void __launch_bounds__(FWD_THREADS, 1) tc_forward_kernel(FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int net_id = blockIdx.x % a.n_nets;
  const FwdNet& nt = a.net[net_id];
  const long long worker = blockIdx.x / a.n_nets, n_workers = gridDim.x / a.n_nets;
  const uint32_t pack_pad = (nt.pack_bytes + 127u) & ~127u;
  uint8_t* s_pack = smem;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + pack_pad);
  // barriers (slot s): [0] weights, [1+s] acc_full, [3+s] out_full, [5+s] x0_full, [7+s] x0_free
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 9);
  float* s_bw = reinterpret_cast<float*>(s_bar + 10);
  float* s_lt = s_bw + 32;
  const float* s_f = reinterpret_cast<const float*>(s_pack + nt.f32_off);
  const uint32_t bar_w = smem_u32(s_bar), bar_acc0 = bar_w + 8, bar_out0 = bar_w + 24, bar_x0full0 = bar_w + 40, bar_x0free0 = bar_w + 56;

  if (warp == FWD_X0_WARP0) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar_acc0 + 8 * s, 1);
        mbar_init(bar_out0 + 8 * s, 1);
        mbar_init(bar_x0full0 + 8 * s, 4);     // one arrival per producer warp
        mbar_init(bar_x0free0 + 8 * s, 1);
      }
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), 512);
  }
  if (a.zero_terms && blockIdx.x == 0 && threadIdx.x < NERFCA_N_LOSS_TERMS) a.zero_terms[threadIdx.x] = 0.0;
  {
    const EncDesc& e = nt.x0.enc;
    for (int i = threadIdx.x; i < 32; i += blockDim.x) s_bw[i] = (e.band_weight && i < e.n_freq) ? __ldg(e.band_weight + i) : 1.f;
    const int n_lt = e.n_latent > 0 ? e.n_phases * e.n_latent : 0;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lt[i] = (i < n_lt) ? __ldg(e.latents + i) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const long long n_my = (a.n_tiles > worker) ? (a.n_tiles - worker + n_workers - 1) / n_workers : 0;
  const bool stash_on = nt.stash != nullptr;
  const bool lt_in_smem = nt.x0.enc.n_phases * nt.x0.enc.n_latent <= 256;
  [[maybe_unused]] int tl_n = 0;

  if (warp >= FWD_X0_WARP0) {
    // ================= 4 producer warps: thread = tile row, tiles alternate between the slots =================
    reg_dealloc<64>();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (warp == FWD_X0_WARP0 && lane == 0) {   // the packed parameter block -> shared memory
      mbar_expect_tx(bar_w, nt.pack_bytes);
      for (uint32_t off = 0; off < nt.pack_bytes; off += 32768u) {
        const uint32_t n = (nt.pack_bytes - off < 32768u) ? nt.pack_bytes - off : 32768u;
        bulk_g2s(smem_u32(s_pack + off), nt.pack + off, n, bar_w);
      }
    }
    __syncwarp();
    const float* lat_tab = lt_in_smem ? s_lt : nt.x0.enc.latents;
    const float* band_w = nt.x0.enc.band_weight ? s_bw : nullptr;
    uint32_t ph_free[2] = {0, 0};
    RowIn rin;
    {
      const long long p0 = worker * TILE_M + row;
      rin = fetch_row(nt.x0, a.src, p0, n_my > 0 && p0 < a.src.n_points);
    }
    for (long long i = 0; i < n_my; ++i) {
      const int s = (int)(i & 1);
      const RowIn cur = rin;
      const long long pn = (worker + (i + 1) * n_workers) * TILE_M + row;
      rin = fetch_row(nt.x0, a.src, pn, i + 1 < n_my && pn < a.src.n_points);
      if (i >= 2) {              // layer 0 of the slot's previous tile has consumed the region
        mbar_wait(bar_x0free0 + 8 * s, ph_free[s]);
        ph_free[s] ^= 1;
        tc_fence_after();
      }
      const TmemSink sink{tmem + FWD_X0 + s * 48 + ((uint32_t)(q * 32) << 16)};
      emit_x0_row(nt.x0, cur, band_w, lat_tab, sink, 0);
      emit_x0_row(nt.x0, cur, band_w, lat_tab, sink, 1);
      tmem_st_wait();
      warp_arrive(bar_x0full0 + 8 * s, lane);
    }
  } else {
    // ================= 16 epilogue warps: slot = warp / 8, column half ch, TMEM lane quadrant q =================
    reg_alloc<104>();                // 16 x 104 + 4 x 64 = 20 x 96: the registers the producers hand back (the CTA pool starts empty)
    const int slot = warp >> 3, ch = (warp >> 2) & 1, q = warp & 3;
    const bool issuer_warp = (warp & 7) == 0;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t t_acc = t_lane + FWD_ACC + slot * 128 + ch * 64;    // this thread's 64 accumulator columns
    const uint32_t t_a = t_lane + FWD_A + slot * 64 + ch * 32;         // ... and the 32 words they become
    const uint32_t bar_acc = bar_acc0 + 8 * slot, bar_out = bar_out0 + 8 * slot, bar_x0full = bar_x0full0 + 8 * slot,
                   bar_x0free = bar_x0free0 + 8 * slot;
    // issue side (used by the elected lane of the slot's first warp)
    constexpr uint32_t idesc = instr_desc(128, 128, 0, 0), idesc_out = instr_desc(128, 16, 0, 0);
    const int k0steps = nt.x0.kpad0 / 16;
    const uint32_t w_base = smem_u32(s_pack);
    const uint32_t td_acc = tmem + FWD_ACC + slot * 128, td_a = tmem + FWD_A + slot * 64, td_x0 = tmem + FWD_X0 + slot * 48,
                   td_out = tmem + FWD_OUT + slot * 16;
    [[maybe_unused]] const int tl_base = slot ? 500 : 3000;
    uint32_t ph_acc = 0, ph_out = 0, ph_x0 = 0;
    mbar_wait(bar_w, 0);             // weights / biases are in shared memory
    const float b_out = s_f[TC_N_RELU * 128 + 128];
    // per-thread constants of the layer loop, pinned in registers (the compiler otherwise re-derives them from %tid and the
    // shared-window base, ~25-cycle S2R reads on the critical path of every layer)
    uint32_t k_acc = t_acc, k_a = t_a, k_bias = smem_u32(s_f) + (uint32_t)(ch * 64) * 4u, k_bar_acc = bar_acc;
    uint32_t k_stash_off = (uint32_t)(ch * 8) * CHUNK_BYTES + (uint32_t)row * 16u;
    uint32_t k_mask_off = (uint32_t)row * 16u + (uint32_t)ch * 8u;
    pin(k_acc); pin(k_a); pin(k_bias); pin(k_bar_acc); pin(k_stash_off); pin(k_mask_off);

    auto issue_layer0 = [&]() {      // whole warp; the X0 region of the slot's next tile feeds layer 0
      if (elect_one()) {
        mbar_wait(bar_x0full, ph_x0);
        tc_fence_after();
        NERFCA_TL(true, tl_base);
        const Desc d_w0 = kmajor(w_base);
        umma_ts_k<5, KSTEP_KMAJOR>(td_acc, td_x0, d_w0, idesc, 0);
        if (k0steps > 5) umma_ts(td_acc, td_x0 + 5 * KSTEP_TMEM, d_w0.lo + 5 * KSTEP_KMAJOR, d_w0.hi, idesc, 1);
        umma_commit(bar_x0free);
        umma_commit(bar_acc);
      }
      ph_x0 ^= 1;
      __syncwarp();
    };
    if (issuer_warp && slot < n_my) issue_layer0();

    for (long long i = slot; i < n_my; i += 2) {
      const long long tile = worker + i * n_workers;
      const long long p = tile * TILE_M + row;
      const bool valid = p < a.src.n_points;
#pragma unroll 1
      for (int l = 0; l < TC_N_RELU; ++l) {
        // H_l = relu(Z_l + b_l) -> bf16 A operand of the next layer (layer 0's bias came through the constant-1 column), in two
        // groups of 32 columns; the bias of the first group is fetched before the accumulator wait
        const uint32_t bias = k_bias + (uint32_t)l * 512u;
        // tile-canonical stash bytes: chunk c of the row at c * 2048 + row * 16 (a warp writes 512 contiguous bytes per chunk);
        // layer 4 leaves only the ReLU pattern of H4 (what the backward needs), one bit per element
        uint8_t* dst = nt.stash + (size_t)tile * STASH_STRIDE + (size_t)(l < 4 ? l : STASH_TILES) * TILE_BYTES + ((l < 4) ? k_stash_off : k_mask_off);
        mbar_wait(k_bar_acc, ph_acc);
        ph_acc ^= 1;
        tc_fence_after();
        NERFCA_TL(lane == 0 && (warp & 7) == 1, 1010 + slot * 1000 + l * 10);
        // both halves are requested at once: a tcgen05.ld round trip takes ~290 cycles while the other slot's MMAs run
        uint32_t v0[32], v1[32];
        tmem_ld32(k_acc, v0);
        tmem_ld32(k_acc + 32, v1);
        tmem_ld_wait();
        NERFCA_TL(lane == 0 && (warp & 7) == 1, 1012 + slot * 1000 + l * 10);
        uint32_t mbits[2] = {0u, 0u};
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t w[16];
          float4 b[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) b[j] = lds_f4(bias + (uint32_t)(32 * g + 4 * j) * 4u);
          if (l == 0) relu_pack32<false>(g ? v1 : v0, b, w);
          else relu_pack32<true>(g ? v1 : v0, b, w);
          NERFCA_TL(lane == 0 && (warp & 7) == 1, 1014 + 2 * g + slot * 1000 + l * 10);
          tmem_st16(k_a + 16 * g, w);
          NERFCA_TL(lane == 0 && (warp & 7) == 1, 1015 + 2 * g + slot * 1000 + l * 10);
          if (stash_on) {
            if (l < 4) {
#pragma unroll
              for (int c = 0; c < 4; ++c)      // (issued after the barrier, in the shadow of the next MMA, they measured 6 % slower: r5a, 254 vs 240 us)
                __stcs(reinterpret_cast<uint4*>(dst + (4 * g + c) * CHUNK_BYTES), make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
            } else {
              // every half of w is a non-negative bf16: adding 0x7FFF carries into the half's top bit exactly when it is nonzero (and
              // never beyond: the largest half is 0x7F80); word i's two flags end up at bits i and 16 + i
              uint32_t m = 0;
#pragma unroll
              for (int i2 = 0; i2 < 16; ++i2) m = (m >> 1) | ((w[i2] + 0x7FFF7FFFu) & 0x80008000u);
              mbits[g] = m;
            }
          }
        }
        if (stash_on && l == 4) __stcs(reinterpret_cast<uint2*>(dst), make_uint2(mbits[0], mbits[1]));
        tmem_st_wait();
        NERFCA_TL(lane == 0 && (warp & 7) == 1, 1011 + slot * 1000 + l * 10);
        tc_fence_before();
        named_bar_sync(1 + slot, 256);          // the slot's H_l is complete and its accumulator has been read
        if (issuer_warp) {
          tc_fence_after();
          if (elect_one()) {
            NERFCA_TL(true, tl_base + 10 + l * 10);
            if (l + 1 < TC_N_RELU) {
              umma_ts_k<8, KSTEP_KMAJOR>(td_acc, td_a, kmajor(w_base + nt.w0_bytes + (uint32_t)l * TILE_BYTES), idesc, 0);
              umma_commit(bar_acc);
            } else {
              umma_ts_k<8, WOUT_KSTEP>(td_out, td_a, kmajor_rows(w_base + nt.wout_off, 256), idesc_out, 0);
              umma_commit(bar_out);
            }
            NERFCA_TL(true, tl_base + 11 + l * 10);
          }
          __syncwarp();
          if (l + 1 == TC_N_RELU && i + 2 < n_my) issue_layer0();   // the accumulator is free: start the slot's next tile
        }
      }
      // output layer: raw = H4 . (w_hi + w_lo) + b_out sits in columns 0 (hi part) and 1 (lo part) of the OUT accumulator
      mbar_wait(bar_out, ph_out);
      ph_out ^= 1;
      tc_fence_after();
      if (ch == 0) {
        uint32_t v[2];
        tmem_ld2(t_lane + FWD_OUT + slot * 16, v);
        tmem_ld_wait();
        const float raw = (__uint_as_float(v[0]) + __uint_as_float(v[1])) + b_out;
        if (valid && nt.raw_out) nt.raw_out[p] = raw;
        if (nt.ray_sum) {
          // X-ray line integral (train/model_helpers.py:72-97) right here: sigma = act(raw) * 1e-2, weight = sigma * delta_s, summed per
          // ray.  The warp's 32 lanes are 32 consecutive samples of at most a few rays: segmented inclusive scan by warp shuffles
          // (a lane adds the value `o` lanes below it while that lane belongs to the same ray), the last lane of every segment
          // adds the segment's sum to the ray with one atomic -- one atomic per warp while the warp stays inside one ray.
          const int n = a.src.n_depth;
          const long long pa = p + a.src.base;
          const int ray = valid ? (int)(pa / n) : -1;
          float wgt = 0.f;
          if (valid) {
            const int sidx = (int)(pa - (long long)ray * n);
            const float delta = (sidx == n - 1) ? 1e-10f : __fsub_rn(__ldg(a.src.depth + sidx + 1), __ldg(a.src.depth + sidx));
            wgt = __fmul_rn(__fmul_rn(act_fwd(nt.act, raw), 0.01f), delta);
          }
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const float w2 = __shfl_up_sync(0xffffffffu, wgt, o);
            const int r2 = __shfl_up_sync(0xffffffffu, ray, o);
            if (lane >= o && r2 == ray) wgt += w2;
          }
          const int r_next = __shfl_down_sync(0xffffffffu, ray, 1);
          if (valid && (lane == 31 || r_next != ray)) atomicAdd(nt.ray_sum + ray, wgt);
        }
      }
      NERFCA_TL(lane == 0 && (warp & 7) == 1, 1060 + slot * 1000);
      // (the next output MMA of this slot is issued after five more named barriers of the slot: OUT has long been read)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == FWD_X0_WARP0) tmem_dealloc(tmem, 512);
}

// =====================================================================================================================
// backward, shared pieces
// =====================================================================================================================
// vector reduction into global memory (4 consecutive floats, 16-byte aligned)
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// flush a TMEM-resident [128 out x n_cols] weight-gradient accumulator: thread = (row, column half)
__device__ __forceinline__ void flush_wgrad(uint32_t t_lane, uint32_t col0, float* gw, int row, int ch, int n_cols, int k_in) {
  const bool vec_ok = (k_in & 3) == 0 && (reinterpret_cast<uintptr_t>(gw) & 15) == 0;
  for (int c0 = ch * 64; c0 < n_cols && c0 < ch * 64 + 64; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(t_lane + col0 + c0, v);
    tmem_ld_wait();
    float* dst = gw + (size_t)row * k_in + c0;
    if (vec_ok && c0 + 16 <= k_in) {
#pragma unroll
      for (int e = 0; e < 16; e += 4)
        red_add_v4(dst + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (c0 + e < k_in) atomicAdd(dst + e, __uint_as_float(v[e]));
    }
  }
}

struct BwdNet {
  const uint8_t* pack;
  const uint8_t* stash;
  uint8_t* handoff;          // [ring][TILE_BYTES]  dZ2 of tile t sits in slot t % ring (top role -> bottom role, through L2)
  uint32_t* produced;        // [n_tiles] 8 once the top role's dZ2 of the tile is visible GPU-wide
  uint32_t* consumed;        // [n_tiles] 1 once the bottom role's copy of the slot has landed in its shared memory
  const float* d_raw;
  float* g_w[NERFCA_MAX_LAYERS];
  float* g_b[NERFCA_MAX_LAYERS];   // may be null
  float* g_lat;
  const float* w0_f32;       // nn.Linear weight of layer 0, [128, in_dim] fp32
  const float* w4_f32;       // nn.Linear weight of layer 4, [128, 128] fp32
  X0Desc x0;
  uint32_t w0_bytes, f32_off;
  int in_dim, enc_dim, n_latent, n_phases;
};
struct BwdArgs {
  SampleSrc src;
  BwdNet net[2];
  int n_nets;
  long long n_tiles;
  int n_top, n_bot;          // CTAs per net in the top / bottom role (grid = n_nets * (n_top + n_bot))
  int ring;                  // hand-off slots per net
  int role;                  // 0: both roles in this launch, 1: every CTA runs the top role, 2: every CTA runs the bottom role
  long long* dbg;   // optional event timeline of one CTA (NERFCA_TIMELINE=top|bot, NERFCA_TIMELINE_CTA=n)
  int dbg_cta;
};

// ---- hand-off flags in global memory (top role -> bottom role, both resident in the same launch) ------------------------------
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// CTA-scope monotonic counter in shared memory (epilogue warps -> publisher warp): unlike an mbarrier phase it cannot be overrun
__device__ __forceinline__ void red_release_cta_shared_add(uint32_t addr, uint32_t v) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// generic-proxy global writes <-> async-proxy (bulk copy) reads of the same bytes
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Bounded spin on a flag: a protocol bug traps (launch failure) instead of hanging the GPU.  (Polling with relaxed loads + one
// acquire fence after the successful poll -- to spare the L1 invalidation an acquire load implies -- measured 4 % slower, r3j.)
__device__ __forceinline__ void wait_flag_ge(const uint32_t* p, uint32_t want) {
  if (ld_acquire_gpu(p) >= want) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(p) < want) {
    __nanosleep(64);
    if (clock64() - t0 > 4000000000LL) {
#ifdef NERFCA_TIMELINE_BUILD
      printf("flag wait timeout: block %d warp %d flag %p = %u, want %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), (const void*)p,
             ld_acquire_gpu(p), want);
#endif
      __trap();
    }
  }
}

// =====================================================================================================================
// backward, top pass: output layer, layers 4 and 3
// =====================================================================================================================
// Nothing is recomputed here: the forward stashed H2, H3 and the ReLU pattern of H4 (one byte per element), so a tile is two
// chain steps
//     R = dZ4' = d_raw 1[H4 > 0]  ->  dgrad 4  ->  S = dZ3 = (.) 1[H3 > 0]  ->  dgrad 3  ->  dZ2 = (.) 1[H2 > 0]  -> hand-off
// with the weight-gradient GEMMs (and their N = 16 bias-gradient companions against a constant-1 tile) filling the tensor pipe
// while the epilogue converts.
//   * The epilogue warps issue the MMAs themselves: after the tile of a step is written they meet at a named barrier and the
//     elected lane of warp 0 issues the next GEMMs.
//   * The dgrad GEMMs take their A operand (the gradient tile just produced) from tensor memory (TS form); the shared-memory copy
//     is only read by the weight-gradient GEMMs (MN-major views).
//   * R holds dZ4' WITHOUT the output weight (TMEM gets dZ4 = dZ4' w_out for dgrad 4); then
//       dW4 = diag(w_out) R^T H3,  db4 = w_out . colsum(R),  dw_out[j] = sum_k W4[j,k] (R^T H3)[j,k] + b4[j] colsum(R)[j]
//     (the last because H4 = relu(H3 W4^T + b4)), so H4 itself is never needed.
//   * Four 32 KB tile buffers rotate: role k of tile i (0: H2, 1: H3, 2: R, 3: S) lives in buffer (i + k) % 4, so the loads of
//     tile i + 1 go into the buffers H3 / R of tile i vacate when weight gradient 4 completes.
// TMEM: ACC [0,128) | WG4 [128,256) | WG3 [256,384) | BG4 [384,400) | BG3 [400,416) | A [416,480)
// (each weight-gradient accumulator is 144 columns wide: column 128 collects the bias gradient, see TOP_BUF_STRIDE)
constexpr uint32_t TOP_ACC = 0, TOP_WG4 = 128, TOP_WG3 = 272, TOP_BG4 = TOP_WG4 + 128, TOP_BG3 = TOP_WG3 + 128, TOP_A = 416;
// A tile buffer is followed by a constant [128 x 16] bf16 block whose column 0 is 1: read MN-major as the B operand, buffer + tail
// are a [128 x 144] matrix, so ONE N = 144 GEMM yields dZ^T H (columns 0-127) and colsum(dZ) (column 128).  (The separate N = 16
// bias GEMMs re-read the whole A operand and cost the tensor pipe about as much as the N = 128 ones.)
constexpr uint32_t TOP_BUF_STRIDE = TILE_BYTES + 4096;
constexpr int BWD_THREADS = 16 * 32;   // both roles: warps 0-7 epilogue (warp 0 holds the issuing lane); top: 8 loader; bottom: 8-11 X0, 12 loader
constexpr int BWD_RING = 128;          // hand-off slots per net (4 MB: lives in L2)

// 32 packed words (64 bf16 columns of one row) -> the row's 8 chunks of a tile-canonical shared-memory tile
__device__ __forceinline__ void sts_row64(uint32_t tile_row_addr, const uint32_t (&w)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) sts_u4(tile_row_addr + c * CHUNK_BYTES, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}
// dZ = acc * 1[h > 0] of this thread's 64 columns -> 32 packed words; h = the activation tile in shared memory (row address)
__device__ __forceinline__ void masked_grad_pack64(const uint32_t (&va)[32], const uint32_t (&vb)[32], uint32_t h_row_addr, uint32_t (&w)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 hv = lds_u4(h_row_addr + (uint32_t)c * CHUNK_BYTES);
    const uint32_t* v = (c < 4) ? va : vb;
    const int j = (c & 3) * 8;
    w[4 * c] = mul_bf16x2(pack_bf16x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), relu_mask_bf16x2(hv.x));
    w[4 * c + 1] = mul_bf16x2(pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), relu_mask_bf16x2(hv.y));
    w[4 * c + 2] = mul_bf16x2(pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])), relu_mask_bf16x2(hv.z));
    w[4 * c + 3] = mul_bf16x2(pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])), relu_mask_bf16x2(hv.w));
  }
}
// flush a TMEM-resident [128 out x 128 in] weight-gradient accumulator scaled per output row: thread = (row, column half)
__device__ __forceinline__ void flush_wgrad_scaled(uint32_t t_lane, uint32_t col0, float* gw, int row, int ch, float scale) {
  for (int c0 = ch * 64; c0 < ch * 64 + 64; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(t_lane + col0 + c0, v);
    tmem_ld_wait();
    float* dst = gw + (size_t)row * 128 + c0;
#pragma unroll
    for (int e = 0; e < 16; e += 4)
      red_add_v4(dst + e, scale * __uint_as_float(v[e]), scale * __uint_as_float(v[e + 1]), scale * __uint_as_float(v[e + 2]),
                 scale * __uint_as_float(v[e + 3]));
  }
}

// smem: [W3][W4][4 tile buffers][ones tile 4 KB][H4 pattern 2 KB][fp32: w_out (128)][bf16x2 w_out pairs (64 words)][d_raw 2 x 128][barriers]
__device__ __forceinline__ void bwd_top_role(const BwdArgs& a, const BwdNet& nt, const long long worker, const long long n_workers) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* s_w3 = smem;
  uint8_t* s_w4 = smem + TILE_BYTES;
  uint8_t* s_buf = smem + 2 * TILE_BYTES;            // 4 rotating tile buffers, each followed by its constant-1 block
  uint8_t* s_m4 = s_buf + 4 * TOP_BUF_STRIDE;        // ReLU pattern of H4 of the current tile (1 bit per element)
  float* s_wo = reinterpret_cast<float*>(s_m4 + MASK_BYTES);
  uint32_t* s_wo2 = reinterpret_cast<uint32_t*>(s_wo + 128);   // w_out as packed bf16 pairs
  float* s_g = reinterpret_cast<float*>(s_wo2 + 64);           // d_raw of the tile's rows, two tiles deep (filled by the load warp)
  float* s_gbout = s_g + 256;
  const uint32_t pub_cnt = smem_u32(s_gbout + 1);     // warp arrivals behind dZ2 stores, 8 per tile, never reset
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_gbout + 4);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 12);
  const uint32_t bar_w = smem_u32(s_bar), bar_ldh = bar_w + 8, bar_ldm = bar_w + 16, bar_acc = bar_w + 24, bar_half = bar_w + 32,
                 bar_mfree = bar_w + 40, bar_tile = bar_w + 48, bar_slot = bar_w + 56, bar_h3free = bar_w + 64;

  if (warp == 8) {
    if (lane == 0) {
      mbar_init(bar_w, 1); mbar_init(bar_ldh, 1); mbar_init(bar_ldm, 1); mbar_init(bar_acc, 1); mbar_init(bar_half, 1);
      mbar_init(bar_mfree, 8); mbar_init(bar_tile, 1);
      mbar_init(bar_slot, 1); mbar_init(bar_h3free, 8);
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), 512);
  }
  {
    const float* fb = reinterpret_cast<const float*>(nt.pack + nt.f32_off);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_wo[i] = __ldg(fb + 5 * 128 + i);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_wo2[i] = pack_bf16x2(__ldg(fb + 5 * 128 + 2 * i), __ldg(fb + 5 * 128 + 2 * i + 1));
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x)      // the constant-1 block behind each tile buffer
      reinterpret_cast<uint4*>(s_buf + (size_t)(i >> 8) * TOP_BUF_STRIDE + TILE_BYTES)[i & 255] =
          ((i & 255) < 128) ? make_uint4(0x00003F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) { s_gbout[0] = 0.f; reinterpret_cast<uint32_t*>(s_gbout)[1] = 0u; }
  }
  tc_fence_before();
  fence_proxy_async();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const long long n_my = (a.n_tiles > worker) ? (a.n_tiles - worker + n_workers - 1) / n_workers : 0;
  [[maybe_unused]] int tl_n = 0;

  if (warp >= 8) {
    if (warp >= 12) reg_dealloc<24>(); else reg_dealloc<64>();      // per warpgroup; warps 9-15 only hand their registers over
  }
  if (warp == 8) {
    // ================= load warp =================
    if (lane == 0) {
      mbar_expect_tx(bar_w, 2 * TILE_BYTES);
      bulk_g2s(smem_u32(s_w3), nt.pack + nt.w0_bytes + 2 * (size_t)TILE_BYTES, TILE_BYTES, bar_w);
      bulk_g2s(smem_u32(s_w4), nt.pack + nt.w0_bytes + 3 * (size_t)TILE_BYTES, TILE_BYTES, bar_w);
    }
    uint32_t ph_half = 0, ph_mfree = 0;
    // hand-off slot of tile j: free once the bottom role has copied out the tile that used it `ring` tiles earlier.  The slot
    // barrier gets its arrival for tile j only after step A of tile j has been passed (bar_mfree), i.e. after the epilogue
    // has consumed the arrival for tile j - 1: never more than one phase ahead of its only waiter.
    auto release_slot = [&](long long tile_j, uint32_t seen) {
      if (tile_j >= a.ring && seen == 0u) wait_flag_ge(nt.consumed + (tile_j - a.ring), 1u);
      mbar_arrive(bar_slot);
    };
    for (long long i = 0; i < n_my; ++i) {
      const long long tile = worker + i * n_workers;
      const uint8_t* st = nt.stash + (size_t)tile * STASH_STRIDE;
      // the flag of the previous tile's slot is requested now and looked at after this tile's loads have been issued
      uint32_t seen = 1u;
      if (lane == 0 && i > 0 && tile - n_workers >= a.ring) seen = ld_acquire_gpu(nt.consumed + (tile - n_workers - a.ring));
      // H4 pattern + d_raw of tile i: as soon as step A of tile i - 1 has consumed its own (plain loads for d_raw: the last tile
      // may be ragged; they are published by the arrival on bar_ldm below)
      float gv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long p = tile * TILE_M + lane * 4 + e;
        gv[e] = (p < a.src.n_points) ? __ldg(nt.d_raw + p) : 0.f;
      }
      *reinterpret_cast<float4*>(s_g + (i & 1) * 128 + lane * 4) = make_float4(gv[0], gv[1], gv[2], gv[3]);
      __syncwarp();
      if (lane == 0) {
        if (i > 0) mbar_wait(bar_mfree, ph_mfree);
        NERFCA_TL(true, 2001);
        mbar_expect_tx(bar_ldm, MASK_BYTES);
        bulk_g2s(smem_u32(s_m4), st + STASH_PATTERN4_OFF, MASK_BYTES, bar_ldm);
        // H2, H3 of tile i: into the buffers that H3 / R of tile i - 1 leave when its weight gradient 4 is complete AND step B of that
        // tile has read H3's ReLU pattern (bar_h3free; this also keeps bar_ldh from completing a second phase before every epilogue
        // warp has seen the first: the loads are needed a whole tile later, so the extra wait costs nothing)
        if (i > 0) { mbar_wait(bar_half, ph_half); mbar_wait(bar_h3free, ph_half); }
        NERFCA_TL(true, 2003);
        mbar_expect_tx(bar_ldh, 2 * TILE_BYTES);
        bulk_g2s(smem_u32(s_buf + (size_t)(i & 3) * TOP_BUF_STRIDE), st + 2 * (size_t)TILE_BYTES, TILE_BYTES, bar_ldh);
        bulk_g2s(smem_u32(s_buf + (size_t)((i + 1) & 3) * TOP_BUF_STRIDE), st + 3 * (size_t)TILE_BYTES, TILE_BYTES, bar_ldh);
        if (i > 0) release_slot(tile - n_workers, seen);
        NERFCA_TL(true, 2005);
      }
      if (i > 0) { ph_half ^= 1; ph_mfree ^= 1; }
      __syncwarp();
    }
    if (lane == 0 && n_my > 0) {
      mbar_wait(bar_mfree, ph_mfree);                   // step A of the last tile
      release_slot(worker + (n_my - 1) * n_workers, 0u);
    }
  } else if (warp == 9) {
    // ================= issuing warp: joins the epilogue warps' two named barriers per tile and issues, in the order the results are
    // needed, the dgrad GEMM everybody waits for and then the weight / bias gradient GEMMs nobody waits for.  (The tensor pipe runs
    // MMAs in issue order at ~68 cycles per 128x128x16 step and the issuing lane stalls while the queue is full: ~1300 cycles per
    // step.  As one of the epilogue warps -- the previous arrangement -- that lane made its warp ~500 cycles late for the step's
    // epilogue and the whole CTA waited for it at the next barrier.  Handing over through an mbarrier instead of sharing the named
    // barrier, or issuing from several lanes at once, were both measured slower.)
    const uint32_t w3 = smem_u32(s_w3), w4 = smem_u32(s_w4), bufs = smem_u32(s_buf);
    constexpr uint32_t KM = KSTEP_MNMAJOR;
    constexpr uint32_t id_dgrad = instr_desc(128, 128, 0, 1), id_wgrad = instr_desc(128, 144, 1, 1);
    const uint32_t td_acc = tmem + TOP_ACC, td_a = tmem + TOP_A;
    mbar_wait(bar_w, 0);
    for (long long i = 0; i < n_my; ++i) {
      const uint32_t par = (uint32_t)(i & 1), first = (i > 0) ? 1u : 0u;
      const uint32_t h2 = bufs + (uint32_t)(i & 3) * TOP_BUF_STRIDE, h3 = bufs + (uint32_t)((i + 1) & 3) * TOP_BUF_STRIDE,
                     R = bufs + (uint32_t)((i + 2) & 3) * TOP_BUF_STRIDE, S = bufs + (uint32_t)((i + 3) & 3) * TOP_BUF_STRIDE;
      named_bar_sync(1, 288);                         // step A: R (shared memory) and dZ4 (tensor memory) are complete
      tc_fence_after();
      if (elect_one()) {
        NERFCA_TL(true, 3010);
        umma_ts_k<8, KM>(td_acc, td_a, mnmajor(w4), id_dgrad, 0);                                    // dH3 = dZ4 W4
        umma_commit(bar_acc);
        mbar_wait(bar_ldh, par);
        tc_fence_after();
        umma_k<8, KM, KM>(tmem + TOP_WG4, mnmajor(R), mnmajor(h3), id_wgrad, first);                 // WG4 += R^T [H3 | 1]
        umma_commit(bar_half);
        NERFCA_TL(true, 3011);
      }
      __syncwarp();
      named_bar_sync(1, 288);                         // step B: S and dZ3 are complete
      tc_fence_after();
      if (elect_one()) {
        NERFCA_TL(true, 3020);
        umma_ts_k<8, KM>(td_acc, td_a, mnmajor(w3), id_dgrad, 0);                                    // dH2 = dZ3 W3
        umma_commit(bar_acc);
        umma_k<8, KM, KM>(tmem + TOP_WG3, mnmajor(S), mnmajor(h2), id_wgrad, first);                 // WG3 += S^T [H2 | 1]
        umma_commit(bar_tile);
        NERFCA_TL(true, 3021);
      }
      __syncwarp();
    }
  } else if (warp == 10) {
    // ================= publisher: the epilogue warps only bump a shared-memory counter behind their dZ2 stores (a CTA-scope release costs them
    // nothing); the GPU-scope release -- which has to wait until those stores have reached L2 -- is paid here, off the tile's path ====
    if (lane == 0) {                // one polling lane that sleeps between probes: the epilogue warps of its scheduler keep their issue slots
      for (long long i = 0; i < n_my; ++i) {
        const long long t0 = clock64();
        while (ld_acquire_cta_shared(pub_cnt) < 8u * (uint32_t)(i + 1)) {     // a counter, not an mbarrier phase: the release below may take
          __nanosleep(200);                                                      // longer than a tile and must not lose arrivals
          if (clock64() - t0 > 4000000000LL) {
#ifdef NERFCA_TIMELINE_BUILD
            printf("publisher timeout: block %d tile# %lld count %u\n", (int)blockIdx.x, i, ld_acquire_cta_shared(pub_cnt));
#endif
            __trap();
          }
        }
        red_release_gpu_add(nt.produced + (worker + i * n_workers), 8u);
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    reg_alloc<200>();
    // ================= 8 epilogue warps: thread = (row, column half); the elected lane of warp 0 issues the dgrad GEMMs =================
    const int q = warp & 3, ch = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t k_acc = t_lane + TOP_ACC + ch * 64, k_a = t_lane + TOP_A + ch * 32;
    uint32_t k_wo2 = smem_u32(s_wo2) + (uint32_t)(ch * 32) * 4u;
    uint32_t k_rowoff = (uint32_t)(ch * 8) * CHUNK_BYTES + (uint32_t)row * 16u;     // this thread's first chunk inside a tile
    uint32_t k_m4 = smem_u32(s_m4) + (uint32_t)row * 16u + (uint32_t)ch * 8u;
    uint32_t k_buf = smem_u32(s_buf);
    pin(k_acc); pin(k_a); pin(k_wo2); pin(k_rowoff); pin(k_m4); pin(k_buf);
    uint32_t slot = (uint32_t)(worker % a.ring);      // hand-off slot of the tile whose dZ2 is pending, advanced without a division per tile
    const uint32_t slot_step = (uint32_t)(n_workers % a.ring);
    // dZ2 of a tile is packed in step C but leaves one step later, right behind the next tile's step-A barrier: its 32 KB then drain
    // through the store path while the CTA waits for dgrad 4 anyway (issued at the end of step C they sat in front of the next step's
    // shared-memory / mbarrier traffic for ~1000 cycles).  One arrival per warp; the publisher warp raises the GPU-scope flag.
    // The 8 stores of a thread go out in two halves, one in each of the next tile's two MMA waits (issuing all 8 at once kept the warp
    // ~500 cycles in the store path, longer than the wait they were meant to hide in).  (This split was in and out once: with it the
    // round-1 soak test trapped about once per 10^4 steps.  The cause was not the split but the bar_slot / bar_mfree order in step A,
    // see there; the split only widened the skew between the warps.)
    uint32_t dz[32];
    uint8_t* dz_dst = nullptr;
    auto store_dz_half = [&](int half) {       // streaming stores straight to L2 (a plain store also goes through the L1 path and drains ~2x slower)
      if (half == 0) {
        dz_dst = nt.handoff + (size_t)slot * TILE_BYTES + k_rowoff;
        slot += slot_step;
        if (slot >= (uint32_t)a.ring) slot -= (uint32_t)a.ring;
      }
#pragma unroll
      for (int c = 4 * half; c < 4 * half + 4; ++c)
        __stcs(reinterpret_cast<uint4*>(dz_dst + c * CHUNK_BYTES), make_uint4(dz[4 * c], dz[4 * c + 1], dz[4 * c + 2], dz[4 * c + 3]));
      if (half == 1) {
        __syncwarp();
        if (lane == 0) red_release_cta_shared_add(pub_cnt, 1u);
      }
    };
    uint32_t ph_acc = 0;
    float gb_sum = 0.f;
    mbar_wait(bar_w, 0);
    for (long long i = 0; i < n_my; ++i) {
      const uint32_t par = (uint32_t)(i & 1);
      const uint32_t h2 = k_buf + (uint32_t)(i & 3) * TOP_BUF_STRIDE, h3 = k_buf + (uint32_t)((i + 1) & 3) * TOP_BUF_STRIDE,
                     R = k_buf + (uint32_t)((i + 2) & 3) * TOP_BUF_STRIDE, S = k_buf + (uint32_t)((i + 3) & 3) * TOP_BUF_STRIDE;
      uint32_t va[32], vb[32], w[32];
      // ---- step A: R = dZ4' = d_raw 1[H4 > 0] (shared memory, A of wgrad 4), A = dZ4 = dZ4' w_out (tensor memory, A of dgrad 4)
      mbar_wait(bar_ldm, par);
      NERFCA_TL(warp == 1 && lane == 0, 1009);
      if (i > 0) mbar_wait(bar_tile, par ^ 1);      // weight gradient 3 of the previous tile has released the R / S buffers
      NERFCA_TL(warp == 1 && lane == 0, 1010);
      {
        const float g_cur = s_g[par * 128 + row];
        const uint32_t gb = pack_bf16x2(g_cur, g_cur);
        if (ch == 0) gb_sum += g_cur;
        // pattern word g covers this thread's columns [32 g, 32 g + 32): bit i / 16 + i = columns 2i / 2i + 1.  (0 or 1 in each half) times
        // the 16 bf16 bits of d_raw gives the packed pair without a carry between the halves.
        const uint2 mb = lds_u2(k_m4);
        const uint32_t gbits = gb & 0xFFFFu;
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2) {
          w[i2] = ((mb.x >> i2) & 0x00010001u) * gbits;
          w[16 + i2] = ((mb.y >> i2) & 0x00010001u) * gbits;
        }
        sts_row64(R + k_rowoff, w);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 wo = lds_u4(k_wo2 + (uint32_t)c * 16u);
          w[4 * c] = mul_bf16x2(w[4 * c], wo.x);
          w[4 * c + 1] = mul_bf16x2(w[4 * c + 1], wo.y);
          w[4 * c + 2] = mul_bf16x2(w[4 * c + 2], wo.z);
          w[4 * c + 3] = mul_bf16x2(w[4 * c + 3], wo.w);
        }
        tmem_st32(k_a, w);
      }
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async();
      // The slot wait comes BEFORE the arrival on bar_mfree: the loader's next arrival on bar_slot follows its wait for bar_mfree of THIS
      // tile (inside the loop additionally bar_half / bar_h3free, after the loop nothing else), so with the two statements the other
      // way round a warp that was held up between them on the LAST tile could find bar_slot two phases on (arrival n_my - 2 consumed
      // by nobody yet, arrival n_my - 1 already in): its parity wait then never returns and the bounded wait traps.  That is the
      // "once per 10^4 steps" trap of round 1 and the r3l soak failure.
      if (i > 0) mbar_wait(bar_slot, par ^ 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_mfree);        // the pattern and d_raw of this tile have been consumed
      NERFCA_TL(warp == 1 && lane == 0, 1013);
      named_bar_sync(1, 288);       // 8 epilogue warps + the issuing warp
      NERFCA_TL(warp == 1 && lane == 0, 1014);
      // ---- step B: dZ3 = dH3 * 1[H3 > 0] -> S (shared memory, A of wgrad 3) and tensor memory (A of dgrad 3)
      mbar_wait(bar_ldh, par);
      if (i > 0) store_dz_half(0);                    // first half of the previous tile's dZ2: drains while dgrad 4 runs
      mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
      tc_fence_after();
      NERFCA_TL(warp == 1 && lane == 0, 1020);
      ld_acc64(k_acc, va, vb);
      masked_grad_pack64(va, vb, h3 + k_rowoff, w);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_h3free);        // H3's pattern has been read: its buffer may take H2 of the next tile
      NERFCA_TL(warp == 1 && lane == 0, 1021);
      tmem_st32(k_a, w);
      sts_row64(S + k_rowoff, w);
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async();
      NERFCA_TL(warp == 1 && lane == 0, 1023);
      named_bar_sync(1, 288);       // 8 epilogue warps + the issuing warp
      NERFCA_TL(warp == 1 && lane == 0, 1024);
      // ---- step C: dZ2 = dH2 * 1[H2 > 0] -> registers (stored behind the next tile's step-A barrier)
      if (i > 0) store_dz_half(1);                    // second half of the previous tile's dZ2: drains while dgrad 3 runs
      mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
      tc_fence_after();
      NERFCA_TL(warp == 1 && lane == 0, 1030);
      ld_acc64(k_acc, va, vb);
      masked_grad_pack64(va, vb, h2 + k_rowoff, dz);
      tc_fence_before();
      NERFCA_TL(warp == 1 && lane == 0, 1033);
    }
    if (n_my > 0) {                                   // the last tile's dZ2
      mbar_wait(bar_slot, (uint32_t)((n_my - 1) & 1));
      store_dz_half(0);
      store_dz_half(1);
    }
    // ---- flush the TMEM-resident accumulators
    if (ch == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) gb_sum += __shfl_xor_sync(0xffffffffu, gb_sum, o);
      if (lane == 0 && nt.g_b[5]) atomicAdd(s_gbout, gb_sum);
    }
    if (n_my > 0) {
      mbar_wait(bar_tile, (uint32_t)((n_my - 1) & 1));   // the last commit of the issuing lane: every MMA has completed
      tc_fence_after();
      const float* fb = reinterpret_cast<const float*>(nt.pack + nt.f32_off);
      const float wo_row = s_wo[row], b4_row = __ldg(fb + 4 * 128 + row);
      // dw_out[row] = sum_k W4[row, k] (R^T H3)[row, k] + b4[row] colsum(R)[row]   (this thread: its 64 columns, then the bias term)
      float dot = 0.f;
      for (int c0 = ch * 64; c0 < ch * 64 + 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_lane + TOP_WG4 + c0, v);
        tmem_ld_wait();
        const float* wrow = nt.w4_f32 + (size_t)row * 128 + c0;
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + e));
          dot = fmaf(wv.x, __uint_as_float(v[e]), dot);
          dot = fmaf(wv.y, __uint_as_float(v[e + 1]), dot);
          dot = fmaf(wv.z, __uint_as_float(v[e + 2]), dot);
          dot = fmaf(wv.w, __uint_as_float(v[e + 3]), dot);
        }
      }
      flush_wgrad_scaled(t_lane, TOP_WG4, nt.g_w[4], row, ch, wo_row);
      flush_wgrad_scaled(t_lane, TOP_WG3, nt.g_w[3], row, ch, 1.f);
      if (ch == 0) {
        uint32_t v4[16], v3[16];
        tmem_ld16(t_lane + TOP_BG4, v4);
        tmem_ld16(t_lane + TOP_BG3, v3);
        tmem_ld_wait();
        const float cs4 = __uint_as_float(v4[0]), cs3 = __uint_as_float(v3[0]);
        if (nt.g_b[4]) atomicAdd(nt.g_b[4] + row, wo_row * cs4);
        if (nt.g_b[3]) atomicAdd(nt.g_b[3] + row, cs3);
        dot = fmaf(b4_row, cs4, dot);
      }
      atomicAdd(nt.g_w[5] + row, dot);
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && nt.g_b[5] && n_my > 0) atomicAdd(nt.g_b[5], s_gbout[0]);
  if (warp == 8) tmem_dealloc(tmem, 512);
}

// =====================================================================================================================
// backward, bottom pass: layers 2, 1, 0 (+ latent gradients)
// =====================================================================================================================
// Loads dZ2 (hand-off from the top pass), H1 and H0 (forward stash); nothing is recomputed except the encoded input X0, which
// four producer warps rebuild one tile ahead.  Two chain steps per tile
//     dgrad 2 -> dZ1 = (.) 1[H1 > 0] -> dgrad 1 -> dZ0 = (.) 1[H0 > 0]
// plus the weight-gradient GEMMs (dZ2^T H1, dZ1^T H0, dZ0^T X0) and their bias companions, which run while the epilogue converts.
// Four 32 KB buffers rotate: role k of tile i (0: dZ2 and later dZ0, 1: H1, 2: H0, 3: dZ1) lives in buffer (i + k) % 4; the load
// of dZ2 of tile i + 1 goes into H1's buffer as soon as weight gradient 2 is complete and step B has read H1's ReLU pattern, the
// loads of H1 / H0 into the H0 / dZ1 buffers when weight gradient 1 is complete.
// TMEM: ACC [0,128) | WG2 [128,256) | WG1 [256,384) | WG0 [384,480) | BG2 [480,496) | BG1 [496,512)
constexpr uint32_t BOT_ACC = 0, BOT_WG2 = 128, BOT_WG1 = 256, BOT_WG0 = 384, BOT_BG2 = 480, BOT_BG1 = 496;

// smem: [W1][W2][4 tile buffers][X0: 96 * 256][W0 latent chunks 4 KB][ones tile 4 KB][latent acc 256 f32][barriers]
__device__ __forceinline__ void bwd_bot_role(const BwdArgs& a, const BwdNet& nt, const long long worker, const long long n_workers) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool onehot = nt.n_latent > 0 && nt.x0.onehot > 0;    // latent gradient through the one-hot columns of X0
  const bool has_lat = nt.n_latent > 0 && !onehot;            // fallback: explicit latent dgrad + scatter by phase
  const int kpad0 = nt.x0.kpad0;
  const int lat_c0 = nt.enc_dim / 8;                                        // first chunk holding latent columns
  const int lat_n = has_lat ? ((nt.enc_dim % 8 + nt.n_latent + 15) / 16) * 16 : 0;   // MMA N covering them
  uint8_t* s_w1 = smem;
  uint8_t* s_w2 = smem + TILE_BYTES;
  uint8_t* s_buf = smem + 2 * TILE_BYTES;
  uint8_t* s_x0 = smem + 6 * TILE_BYTES;
  uint8_t* s_w0lat = s_x0 + 96 * 256;
  uint8_t* s_ones = s_w0lat + 4096;
  float* s_lat = reinterpret_cast<float*>(s_ones + 4096);
  const int n_lat_acc = has_lat ? nt.n_phases * nt.n_latent : 0;            // <= 256 checked on the host
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_lat + 256);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 12);
  float* s_bw = reinterpret_cast<float*>(s_tmem + 4);          // band weights / latent table for the X0 warps (see mlp_tc_bwd2.cuh: global
  float* s_lt = s_bw + 32;                                     // loads miss L1 every time here, the flag polls keep invalidating it)
  const bool lt_in_smem = nt.x0.enc.n_phases * nt.x0.enc.n_latent <= 256;
  const uint32_t bar_w = smem_u32(s_bar), bar_ld_dz = bar_w + 8, bar_ld_h1 = bar_w + 16, bar_ld_h0 = bar_w + 24, bar_acc = bar_w + 32,
                 bar_free1 = bar_w + 40, bar_free2 = bar_w + 48, bar_x0 = bar_w + 56, bar_x0free = bar_w + 64;

  if (warp == 12) {
    if (lane == 0) {
      mbar_init(bar_w, 1); mbar_init(bar_ld_dz, 1); mbar_init(bar_ld_h1, 1); mbar_init(bar_ld_h0, 1); mbar_init(bar_acc, 1);
      mbar_init(bar_free1, 1); mbar_init(bar_free2, 1); mbar_init(bar_x0, 4); mbar_init(bar_x0free, 1);
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), 512);
  }
  {
    const EncDesc& e = nt.x0.enc;
    for (int i = threadIdx.x; i < 32; i += blockDim.x) s_bw[i] = (e.band_weight && i < e.n_freq) ? __ldg(e.band_weight + i) : 1.f;
    const int n_lt = e.n_latent > 0 ? e.n_phases * e.n_latent : 0;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lt[i] = (i < n_lt && lt_in_smem) ? __ldg(e.latents + i) : 0.f;
    for (int i = threadIdx.x; i < n_lat_acc; i += blockDim.x) s_lat[i] = 0.f;
    // ones tile: column 0 == 1 in every row (bias gradients = column sums of dZ)
    for (int i = threadIdx.x; i < 4096 / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(s_ones)[i] = (i < 128) ? make_uint4(0x00003F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  tc_fence_before();
  fence_proxy_async();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const long long n_my = (a.n_tiles > worker) ? (a.n_tiles - worker + n_workers - 1) / n_workers : 0;
  [[maybe_unused]] int tl_n = 0;

  if (warp >= 12) {
    // ================= warp 12: loader, warp 13: issuing warp (warps 14-15 only fill the warpgroup) =================
    reg_dealloc<48>();
    if (warp == 13) {
      // joins the epilogue warps' named barriers and issues every GEMM of the tile in the order the results are needed (see the top role)
      const uint32_t w1 = smem_u32(s_w1), w2 = smem_u32(s_w2), ones = smem_u32(s_ones), x0 = smem_u32(s_x0), w0lat = smem_u32(s_w0lat),
                     bufs = smem_u32(s_buf);
      constexpr uint32_t KK = KSTEP_KMAJOR, KM = KSTEP_MNMAJOR;
      constexpr uint32_t id_dgrad = instr_desc(128, 128, 0, 1), id_wgrad = instr_desc(128, 128, 1, 1), id_side = instr_desc(128, 16, 1, 1);
      const uint32_t id_wg0 = instr_desc(128, kpad0, 1, 1), id_lat = instr_desc(128, lat_n > 0 ? lat_n : 16, 0, 1);
      const uint32_t td_acc = tmem + BOT_ACC;
      mbar_wait(bar_w, 0);
      for (long long i = 0; i < n_my; ++i) {
        const uint32_t par = (uint32_t)(i & 1), first = (i > 0) ? 1u : 0u;
        const uint32_t dz2 = bufs + (uint32_t)(i & 3) * TILE_BYTES, h1 = bufs + (uint32_t)((i + 1) & 3) * TILE_BYTES,
                       h0 = bufs + (uint32_t)((i + 2) & 3) * TILE_BYTES, dz1 = bufs + (uint32_t)((i + 3) & 3) * TILE_BYTES, dz0 = dz2;
        // ---- tile start: dH1 = dZ2 W2 (the barrier that closed the previous tile released the accumulator)
        tc_fence_after();
        if (elect_one()) {
          mbar_wait(bar_ld_dz, par);
          tc_fence_after();
          NERFCA_TL(true, 3000);
          umma_k<8, KK, KM>(td_acc, kmajor(dz2), mnmajor(w2), id_dgrad, 0);
          umma_commit(bar_acc);
          mbar_wait(bar_ld_h1, par);
          tc_fence_after();
          umma_k<8, KM, KM>(tmem + BOT_WG2, mnmajor(dz2), mnmajor(h1), id_wgrad, first);             // WG2 += dZ2^T H1
          umma_k<8, KM, KM>(tmem + BOT_BG2, mnmajor(dz2), mnmajor(ones), id_side, first);            // BG2 += colsum(dZ2)
          NERFCA_TL(true, 3001);
        }
        __syncwarp();
        named_bar_sync(1, 288);                       // step B: dZ1 is complete
        tc_fence_after();
        if (elect_one()) {
          NERFCA_TL(true, 3010);
          umma_commit(bar_free1);      // wgrad 2 complete and H1's pattern read: H1's buffer may take dZ2 of the next tile
          umma_k<8, KK, KM>(td_acc, kmajor(dz1), mnmajor(w1), id_dgrad, 0);                          // dH0 = dZ1 W1
          umma_commit(bar_acc);
          mbar_wait(bar_ld_h0, par);
          tc_fence_after();
          umma_k<8, KM, KM>(tmem + BOT_WG1, mnmajor(dz1), mnmajor(h0), id_wgrad, first);             // WG1 += dZ1^T H0
          umma_k<8, KM, KM>(tmem + BOT_BG1, mnmajor(dz1), mnmajor(ones), id_side, first);            // BG1 += colsum(dZ1)
          NERFCA_TL(true, 3011);
        }
        __syncwarp();
        named_bar_sync(1, 288);                       // step C: dZ0 is complete
        tc_fence_after();
        if (elect_one()) {
          NERFCA_TL(true, 3020);
          umma_commit(bar_free2);      // wgrad 1 complete and H0's pattern read: the H0 / dZ1 buffers may be reloaded
          mbar_wait(bar_x0, par);
          tc_fence_after();
          umma_k<8, KM, KM>(tmem + BOT_WG0, mnmajor(dz0), mnmajor(x0), id_wg0, first);               // WG0 += dZ0^T X0 (same lane as the next
          umma_commit(bar_x0free);                                                                    // dgrad 2: its commit covers this read of dZ0)
          if (has_lat) {
            umma_k<8, KK, KM>(td_acc, kmajor(dz0), mnmajor(w0lat), id_lat, 0);                       // latent columns of dX0
            umma_commit(bar_acc);
          }
          NERFCA_TL(true, 3021);
        }
        __syncwarp();
        if (has_lat) named_bar_sync(1, 288);          // the epilogue has read the latent columns: the accumulator is free again
      }
    }
    if (warp == 12 && lane == 0) {
      const uint32_t lat_bytes = has_lat ? (uint32_t)(lat_n / 8) * CHUNK_BYTES : 0u;
      mbar_expect_tx(bar_w, 2 * TILE_BYTES + lat_bytes);
      bulk_g2s(smem_u32(s_w1), nt.pack + nt.w0_bytes, TILE_BYTES, bar_w);
      bulk_g2s(smem_u32(s_w2), nt.pack + nt.w0_bytes + (size_t)TILE_BYTES, TILE_BYTES, bar_w);
      if (has_lat) bulk_g2s(smem_u32(s_w0lat), nt.pack + (size_t)lat_c0 * CHUNK_BYTES, lat_bytes, bar_w);
      uint32_t ph1 = 0, ph2 = 0;
      uint32_t slot = (uint32_t)(worker % a.ring);    // hand-off slot of the current tile, advanced without a division per tile
      const uint32_t slot_step = (uint32_t)(n_workers % a.ring);
      for (long long i = 0; i < n_my; ++i) {
        const long long tile = worker + i * n_workers;
        const uint8_t* st = nt.stash + (size_t)tile * STASH_STRIDE;
        const uint32_t seen = ld_acquire_gpu(nt.produced + tile);   // requested now, looked at once the buffer is free
        if (i > 0) {
          mbar_wait(bar_free1, ph1); ph1 ^= 1;                  // H1's buffer of tile i - 1
          mbar_wait(bar_ld_dz, (uint32_t)((i - 1) & 1));        // (long complete) dZ2 of tile i - 1 has left its hand-off slot
          st_release_gpu(nt.consumed + (tile - n_workers), 1u);
        }
        NERFCA_TL(true, 2001);
        if (seen < 8u) wait_flag_ge(nt.produced + tile, 8u);    // all 8 epilogue warps of the top role have written the tile
        NERFCA_TL(true, 2002);
        fence_proxy_async_all();
        mbar_expect_tx(bar_ld_dz, TILE_BYTES);
        bulk_g2s(smem_u32(s_buf + (size_t)(i & 3) * TILE_BYTES), nt.handoff + (size_t)slot * TILE_BYTES, TILE_BYTES, bar_ld_dz);
        slot += slot_step;
        if (slot >= (uint32_t)a.ring) slot -= (uint32_t)a.ring;
        if (i > 0) { mbar_wait(bar_free2, ph2); ph2 ^= 1; }     // H0's and dZ1's buffers of tile i - 1
        NERFCA_TL(true, 2003);
        mbar_expect_tx(bar_ld_h1, TILE_BYTES);
        bulk_g2s(smem_u32(s_buf + (size_t)((i + 1) & 3) * TILE_BYTES), st + (size_t)TILE_BYTES, TILE_BYTES, bar_ld_h1);
        mbar_expect_tx(bar_ld_h0, TILE_BYTES);
        bulk_g2s(smem_u32(s_buf + (size_t)((i + 2) & 3) * TILE_BYTES), st, TILE_BYTES, bar_ld_h0);
        // H0 / H1 of the NEXT tile can only be copied once this tile's step C has released their buffers, and step B of the next tile
        // then waits for H1: pull them into L2 now, so that copy is an L2 hit instead of an HBM round trip
        if (i + 1 < n_my) {
          const uint8_t* nx = nt.stash + (size_t)(tile + n_workers) * STASH_STRIDE;
          bulk_prefetch_l2(nx, TILE_BYTES);
          bulk_prefetch_l2(nx + TILE_BYTES, TILE_BYTES);
        }
      }
      if (n_my > 0) {
        mbar_wait(bar_ld_dz, (uint32_t)((n_my - 1) & 1));
        st_release_gpu(nt.consumed + (worker + (n_my - 1) * n_workers), 1u);
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ================= 4 X0 warps: thread = tile row; the encoded input is only needed by the last GEMM of a tile, so these
    // warps run about one tile ahead of the rest of the CTA =================
    reg_dealloc<64>();
    const int row = (warp - 8) * 32 + lane;
    uint32_t ph_x0free = 0;
    RowIn rin;
    {
      const long long p0 = worker * TILE_M + row;
      rin = fetch_row(nt.x0, a.src, p0, n_my > 0 && p0 < a.src.n_points);
    }
    for (long long i = 0; i < n_my; ++i) {
      const RowIn cur = rin;
      const long long pn = (worker + (i + 1) * n_workers) * TILE_M + row;
      rin = fetch_row(nt.x0, a.src, pn, i + 1 < n_my && pn < a.src.n_points);
      if (i > 0) { mbar_wait(bar_x0free, ph_x0free); ph_x0free ^= 1; }
      emit_x0_row(nt.x0, cur, nt.x0.enc.band_weight ? s_bw : nullptr, lt_in_smem ? s_lt : nt.x0.enc.latents, SmemSink{s_x0, row}, 0);
      emit_x0_row(nt.x0, cur, nt.x0.enc.band_weight ? s_bw : nullptr, lt_in_smem ? s_lt : nt.x0.enc.latents, SmemSink{s_x0, row}, 1);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_x0);
    }
  } else {
    // ================= 8 epilogue warps: thread = (row, column half); the elected lane of warp 0 issues =================
    reg_alloc<200>();
    const int q = warp & 3, ch = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t k_acc = t_lane + BOT_ACC + ch * 64;
    uint32_t k_rowoff = (uint32_t)(ch * 8) * CHUNK_BYTES + (uint32_t)row * 16u;     // this thread's first chunk inside a tile
    uint32_t k_buf = smem_u32(s_buf);
    pin(k_acc); pin(k_rowoff); pin(k_buf);
    uint32_t ph_acc = 0;
    mbar_wait(bar_w, 0);
    for (long long i = 0; i < n_my; ++i) {
      const long long tile = worker + i * n_workers;
      const long long p = tile * TILE_M + row;
      const bool valid = p < a.src.n_points;
      const uint32_t par = (uint32_t)(i & 1);
      const uint32_t dz2 = k_buf + (uint32_t)(i & 3) * TILE_BYTES, h1 = k_buf + (uint32_t)((i + 1) & 3) * TILE_BYTES,
                     h0 = k_buf + (uint32_t)((i + 2) & 3) * TILE_BYTES, dz1 = k_buf + (uint32_t)((i + 3) & 3) * TILE_BYTES, dz0 = dz2;
      uint32_t va[32], vb[32], w[32];
      // ---- step B: dZ1 = dH1 * 1[H1 > 0] -> shared memory (A of dgrad 1 and of wgrad 1)
      mbar_wait(bar_ld_h1, par);
      mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
      tc_fence_after();
      NERFCA_TL(warp == 1 && lane == 0, 1010);
      ld_acc64(k_acc, va, vb);
      masked_grad_pack64(va, vb, h1 + k_rowoff, w);
      NERFCA_TL(warp == 1 && lane == 0, 1011);
      sts_row64(dz1 + k_rowoff, w);
      tc_fence_before();
      fence_proxy_async();
      named_bar_sync(1, 288);       // 8 epilogue warps + the issuing warp
      NERFCA_TL(warp == 1 && lane == 0, 1014);
      // ---- step C: dZ0 = dH0 * 1[H0 > 0] -> shared memory (A of wgrad 0), into the buffer dZ2 occupied
      mbar_wait(bar_ld_h0, par);
      mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
      tc_fence_after();
      NERFCA_TL(warp == 1 && lane == 0, 1020);
      ld_acc64(k_acc, va, vb);
      masked_grad_pack64(va, vb, h0 + k_rowoff, w);
      NERFCA_TL(warp == 1 && lane == 0, 1021);
      mbar_wait(bar_free1, par);       // weight gradient 2 no longer reads dZ2
      sts_row64(dz0 + k_rowoff, w);
      tc_fence_before();
      fence_proxy_async();
      named_bar_sync(1, 288);       // 8 epilogue warps + the issuing warp
      NERFCA_TL(warp == 1 && lane == 0, 1024);
      // ---- latent gradient (fallback): columns [enc_dim, enc_dim + T) of dX0 sit at accumulator columns enc_dim - 8 * lat_c0 + t
      if (has_lat) {
        mbar_wait(bar_acc, ph_acc); ph_acc ^= 1;
        tc_fence_after();
        if (ch == 0) {
          // a warp's 32 rows are consecutive samples, almost always of one ray (one phase): reduce over the warp
          // first and add once; rows of a warp that straddles two rays fall back to per-lane atomics
          const int phase = valid ? load_phase(a.src, p) : -1;
          const int ph0 = __shfl_sync(0xffffffffu, phase, 0);
          const bool uniform = __all_sync(0xffffffffu, phase == ph0 || phase < 0);
          const bool ok = phase >= 0 && phase < nt.n_phases;
          uint32_t v[16];
          for (int c0 = 0; c0 < lat_n; c0 += 16) {
            tmem_ld16(t_lane + BOT_ACC + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int t = c0 + e - (nt.enc_dim - 8 * lat_c0);
              if (t < 0 || t >= nt.n_latent) continue;      // warp-uniform
              float val = ok ? __uint_as_float(v[e]) : 0.f;
              if (uniform) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                if (lane == 0 && ph0 >= 0 && ph0 < nt.n_phases) atomicAdd(&s_lat[ph0 * nt.n_latent + t], val);
              } else if (ok) {
                atomicAdd(&s_lat[phase * nt.n_latent + t], val);
              }
            }
          }
        }
        tc_fence_before();
        named_bar_sync(1, 288);        // (+ the issuing warp) the accumulator has been read: the next tile's dgrad 2 may overwrite it
      }
      // (without the fallback the accumulator was last read before the barrier of step C)
    }
    if (n_my > 0) {
      mbar_wait(bar_x0free, (uint32_t)((n_my - 1) & 1));   // the last commit of the issuing lane: every weight-gradient MMA has completed
      tc_fence_after();
      flush_wgrad(t_lane, BOT_WG2, nt.g_w[2], row, ch, 128, 128);
      flush_wgrad(t_lane, BOT_WG1, nt.g_w[1], row, ch, 128, 128);
      flush_wgrad(t_lane, BOT_WG0, nt.g_w[0], row, ch, kpad0, nt.in_dim);
      if (ch == 0) {
        uint32_t v[16];
        tmem_ld16(t_lane + BOT_BG2, v);
        tmem_ld_wait();
        if (nt.g_b[2]) atomicAdd(nt.g_b[2] + row, __uint_as_float(v[0]));
        tmem_ld16(t_lane + BOT_BG1, v);
        tmem_ld_wait();
        if (nt.g_b[1]) atomicAdd(nt.g_b[1] + row, __uint_as_float(v[0]));
      }
      if (ch == 1 && nt.g_b[0]) {   // bias 0 = the constant-1 column (index in_dim) of wgrad 0
        const int c0 = nt.in_dim & ~15;
        uint32_t v[16];
        tmem_ld16(t_lane + BOT_WG0 + c0, v);
        tmem_ld_wait();
        float val = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (c0 + e == nt.in_dim) val = __uint_as_float(v[e]);
        atomicAdd(nt.g_b[0] + row, val);
      }
      if (onehot && nt.g_lat) {
        // S[n][ph] = sum over the samples of phase ph of dZ0[.][n] sits in columns in_dim + 1 + ph of wgrad 0 (all inside the
        // last 16-column group); d latents[ph][t] = sum_n S[n][ph] * W0[n][enc_dim + t].  S goes through shared memory (the
        // tile buffers are idle now) so that one thread per (ph, t) can run the 128-term dot product.
        float* s_S = reinterpret_cast<float*>(s_buf);
        const int cg = kpad0 - 16;
        if (ch == (cg >> 6)) {
          uint32_t v[16];
          tmem_ld16(t_lane + BOT_WG0 + cg, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int ph = cg + e - (nt.in_dim + 1);
            if (ph >= 0 && ph < nt.x0.onehot) s_S[ph * 128 + row] = __uint_as_float(v[e]);
          }
        }
        tc_fence_before();
        named_bar_sync(1, 256);
        const int tid = (int)threadIdx.x;
        if (tid < nt.x0.onehot * nt.n_latent) {
          const int ph = tid / nt.n_latent, t = tid - ph * nt.n_latent;
          float g = 0.f;
          for (int n = 0; n < 128; ++n) g = fmaf(s_S[ph * 128 + n], __ldg(nt.w0_f32 + (size_t)n * nt.in_dim + nt.enc_dim + t), g);
          atomicAdd(nt.g_lat + tid, g);
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (nt.g_lat && has_lat)
    for (int i = threadIdx.x; i < n_lat_acc; i += blockDim.x)
      if (s_lat[i] != 0.f) atomicAdd(nt.g_lat + i, s_lat[i]);
  if (warp == 12) tmem_dealloc(tmem, 512);
}

// One launch, two roles: of the n_top + n_bot CTAs of a net the first n_top run the top pass and hand every tile's dZ2 to the rest
// through a ring of slots that stays in L2 (flags: `produced` / `consumed` per tile).  Both roles walk the tiles in increasing
// order (worker, worker + n_workers, ...), and all CTAs are resident at once (grid <= #SMs, one CTA per SM), so the smallest
// tile not yet produced is always being worked on: the hand-off cannot deadlock for any ring size >= 1.
__global__ void __launch_bounds__(BWD_THREADS, 1) tc_bwd_kernel(BwdArgs a) {
  const int net_id = blockIdx.x % a.n_nets;
  const int r = blockIdx.x / a.n_nets;
  if (a.role == 1 || (a.role == 0 && r < a.n_top)) bwd_top_role(a, a.net[net_id], r, a.n_top);
  else bwd_bot_role(a, a.net[net_id], a.role == 2 ? r : r - a.n_top, a.n_bot);
}
// single-role launches (two-launch mode)
__global__ void __launch_bounds__(BWD_THREADS, 1) tc_bwd_top_kernel(BwdArgs a) {
  bwd_top_role(a, a.net[blockIdx.x % a.n_nets], blockIdx.x / a.n_nets, a.n_top);
}
__global__ void __launch_bounds__(BWD_THREADS, 1) tc_bwd_bot_kernel(BwdArgs a) {
  bwd_bot_role(a, a.net[blockIdx.x % a.n_nets], blockIdx.x / a.n_nets, a.n_bot);
}

#include "mlp_tc_bwd2.cuh"

// =====================================================================================================================
// host side
// =====================================================================================================================
static size_t fwd_smem_bytes(const NetDims& d) { return (((size_t)d.pack_bytes + 127) & ~(size_t)127) + 10 * 8 + (32 + 256) * 4; }
constexpr size_t TOP_SMEM = 2 * (size_t)TILE_BYTES + 4 * (size_t)TOP_BUF_STRIDE + MASK_BYTES + 128 * 4 + 64 * 4 + 256 * 4 + 16 + 12 * 8 + 16;
constexpr size_t BOT_SMEM = 6 * (size_t)TILE_BYTES + 96 * 256 + 4096 + 4096 + 256 * 4 + 12 * 8 + 16 + (32 + 256) * 4;
constexpr size_t BWD_SMEM = TOP_SMEM > BOT_SMEM ? TOP_SMEM : BOT_SMEM;
static_assert(BWD_SMEM <= 227 * 1024, "backward kernel exceeds the shared memory of an SM");

static size_t n_tiles_of(long long P) { return (size_t)((P + TILE_M - 1) / TILE_M); }

// stash: [net][tile][2][32 KB]
size_t tc_stash_bytes_n(int n_nets, long long P) { return (size_t)n_nets * n_tiles_of(P) * STASH_STRIDE; }
// workspace: [packed blocks][hand-off ring: net, slot, 32 KB][hand-off flags: net, {produced, consumed}, tile]   (the last two: backward only)
static size_t tc_pack_bytes_n(const nerfca_field_t* const* f, int n_nets) {
  size_t n = 0;
  for (int i = 0; i < n_nets; ++i) n += pack_stride(net_dims(*f[i]));
  return n;
}
size_t tc_workspace_bytes_n(const nerfca_field_t* const* f, int n_nets, long long P, int backward) {
  size_t n = tc_pack_bytes_n(f, n_nets);
  // (sized for the two-launch mode, NERFCA_BWD_MERGED=0, whose hand-off holds every tile; the one-launch mode uses BWD_RING slots per net)
  if (backward) n += (size_t)n_nets * (n_tiles_of(P) > (size_t)BWD_RING ? n_tiles_of(P) : (size_t)BWD_RING) * TILE_BYTES +
                     (size_t)n_nets * 2 * n_tiles_of(P) * sizeof(uint32_t);
  return n;
}

static int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static bool timeline_wanted(const char* which) {
  const char* e = getenv("NERFCA_TIMELINE");
  return e && strcmp(e, which) == 0;
}
static int timeline_dump(long long* dbg, cudaStream_t st) {
  std::vector<long long> h(8008);
  NERFCA_CUDA_OK(cudaMemcpyAsync(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
  NERFCA_CUDA_OK(cudaStreamSynchronize(st));
  for (size_t i = 0; i + 1 < h.size(); i += 2)
    if (h[i] != 0) fprintf(stderr, "TL %lld %lld\n", h[i], h[i + 1]);
  cudaFree(dbg);
  return NERFCA_OK;
}

static X0Desc make_x0(const nerfca_field_t& f, const NetDims& d) {
  X0Desc x;
  x.enc = make_enc(f);
  x.kpad0 = d.kpad0;
  x.fast = (x.enc.mode == NERFCA_ENC_BANDS && f.n_freq == FAST_FREQ) ? 1 : 0;
  x.onehot = 0;
  return x;
}

__global__ void __launch_bounds__(BWD_THREADS, 1) tc_bwd_kernel(BwdArgs a);
// can `grid` CTAs of the one-launch backward be resident at the same time on this device (and may it be launched cooperatively)?
static bool bwd_coresident(int grid) {
  static int cached_grid = -1, cached = 0;
  if (cached_grid == grid) return cached != 0;
  int dev = 0, coop = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  bool ok = coop != 0 && cudaFuncSetAttribute(tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tc_bwd_kernel, BWD_THREADS, BWD_SMEM) == cudaSuccess &&
            (long long)per_sm * sm_count() >= grid;
  if (!ok) cudaGetLastError();
  cached_grid = grid;
  cached = ok ? 1 : 0;
  return ok;
}

static unsigned grid_for(int n_nets, long long n_tiles) {
  long long g = sm_count() / n_nets * n_nets;
  const long long need = n_tiles * n_nets;
  if (need < g) g = need;
  return (unsigned)g;
}

// Forward of n_nets (1 or 2) fields over the same sample set in one launch.  pack != 0: (re)pack the parameters into
// the head of `workspace` first.
int tc_fields_forward(const nerfca_field_t* const* f, int n_nets, const nerfca_samples_t& s, float* const* raw_out, void* stash,
                      void* workspace, int pack, double* zero_terms, cudaStream_t st, float* const* ray_sum, int act) {
  if (pack) {
    int rc = pack_params(f, n_nets, workspace, st);
    if (rc) return rc;
  }
  FwdArgs a;
  a.src = make_src(s);
  a.n_nets = n_nets;
  a.n_tiles = (long long)n_tiles_of(s.n_points);
  a.zero_terms = zero_terms;
  size_t off = 0, smem = 0;
  for (int i = 0; i < n_nets; ++i) {
    const NetDims d = net_dims(*f[i]);
    FwdNet& n = a.net[i];
    n.pack = (const uint8_t*)workspace + off;
    off += pack_stride(d);
    n.raw_out = raw_out ? raw_out[i] : nullptr;
    n.ray_sum = ray_sum ? ray_sum[i] : nullptr;
    n.act = act;
    n.stash = stash ? (uint8_t*)stash + (size_t)i * a.n_tiles * STASH_STRIDE : nullptr;
    n.x0 = make_x0(*f[i], d);
    n.w0_bytes = d.w0_bytes; n.wout_off = d.wout_off; n.f32_off = d.f32_off; n.pack_bytes = d.pack_bytes;
    const size_t sm = fwd_smem_bytes(d);
    smem = sm > smem ? sm : smem;
  }
  NERFCA_REQUIRE(smem <= 227 * 1024, NERFCA_E_UNSUPPORTED, "field does not fit the forward kernel's shared memory");
  NERFCA_CUDA_OK(cudaFuncSetAttribute(tc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const bool timeline = timeline_wanted("fwd");   // developer aid: event timeline of one CTA to stderr
  a.dbg = nullptr;
  a.dbg_cta = getenv("NERFCA_TIMELINE_CTA") ? atoi(getenv("NERFCA_TIMELINE_CTA")) : 0;
  if (timeline) {
    NERFCA_CUDA_OK(cudaMalloc(&a.dbg, 8008 * sizeof(long long)));
    NERFCA_CUDA_OK(cudaMemsetAsync(a.dbg, 0, 8008 * sizeof(long long), st));
  }
  {
    ProfScope prof(NERFCA_K_FIELD_FWD, st);
    tc_forward_kernel<<<grid_for(n_nets, a.n_tiles), FWD_THREADS, smem, st>>>(a);
    NERFCA_LAUNCH_OK();
  }
  if (timeline) {
    int rc = timeline_dump(a.dbg, st);
    if (rc) return rc;
  }
  return NERFCA_OK;
}

int tc_fields_backward(const nerfca_field_t* const* f, int n_nets, const nerfca_samples_t& s, const float* const* d_raw,
                       const void* stash, void* workspace, int pack, const nerfca_field_grads_t* const* gr, cudaStream_t st) {
  if (pack) {
    int rc = pack_params(f, n_nets, workspace, st);
    if (rc) return rc;
  }
  BwdArgs a;
  a.src = make_src(s);
  a.n_nets = n_nets;
  a.n_tiles = (long long)n_tiles_of(s.n_points);
  size_t off = 0;
  uint8_t* handoff = (uint8_t*)workspace + tc_pack_bytes_n(f, n_nets);
  // The one-launch mode makes top-role and bottom-role CTAs wait on each other's flags: that is only deadlock-free if every CTA of
  // the grid is resident at once.  It is therefore launched COOPERATIVELY (the driver co-schedules the whole grid or refuses), and
  // only if the occupancy query says grid <= resident CTAs (MPS partitions, MIG slices or green contexts shrink that); otherwise
  // the roles run as two launches with the hand-off through HBM.
  const int merged = env_flag("NERFCA_BWD_MERGED", 1) && bwd_coresident(sm_count() / n_nets * n_nets);
  const size_t ring = merged ? (size_t)BWD_RING : ((size_t)a.n_tiles > (size_t)BWD_RING ? (size_t)a.n_tiles : (size_t)BWD_RING);
  const size_t ring_alloc = (size_t)a.n_tiles > (size_t)BWD_RING ? (size_t)a.n_tiles : (size_t)BWD_RING;
  uint32_t* flags = reinterpret_cast<uint32_t*>(handoff + (size_t)n_nets * ring_alloc * TILE_BYTES);
  for (int i = 0; i < n_nets; ++i) {
    const NetDims d = net_dims(*f[i]);
    BwdNet& n = a.net[i];
    n.pack = (const uint8_t*)workspace + off;
    off += pack_stride(d);
    n.stash = (const uint8_t*)stash + (size_t)i * a.n_tiles * STASH_STRIDE;
    n.handoff = handoff + (size_t)i * ring * TILE_BYTES;
    n.produced = flags + (size_t)(2 * i) * a.n_tiles;
    n.consumed = flags + (size_t)(2 * i + 1) * a.n_tiles;
    n.d_raw = d_raw[i];
    for (int l = 0; l < NERFCA_MAX_LAYERS; ++l) { n.g_w[l] = gr[i]->weight[l]; n.g_b[l] = gr[i]->bias[l]; }
    n.g_lat = gr[i]->latents;
    n.w0_f32 = f[i]->weight[0];
    n.w4_f32 = f[i]->weight[4];
    n.x0 = make_x0(*f[i], d);
    if (f[i]->n_latent > 0 && f[i]->n_phases <= d.kpad0 - d.in_dim - 1) n.x0.onehot = f[i]->n_phases;
    n.w0_bytes = d.w0_bytes; n.f32_off = d.f32_off;
    n.in_dim = d.in_dim; n.enc_dim = d.enc_dim; n.n_latent = f[i]->n_latent; n.n_phases = f[i]->n_phases;
    NERFCA_REQUIRE(f[i]->n_latent == 0 || (d.enc_dim % 8 + f[i]->n_latent + 15) / 16 * 16 + d.enc_dim / 8 * 8 <= d.kpad0,
                   NERFCA_E_UNSUPPORTED, "tcgen05 backward: latent columns do not fit the padded first layer");
    NERFCA_REQUIRE(f[i]->n_latent == 0 || (size_t)f[i]->n_phases * f[i]->n_latent <= 256, NERFCA_E_UNSUPPORTED,
                   "tcgen05 backward: latent table larger than 256 floats (use precision fp32)");
  }
  // role split: n_top + n_bot CTAs per net, all resident at once (one CTA per SM).  NERFCA_BWD_SPLIT="n_top,n_bot" overrides it.
  const int per_net = sm_count() / n_nets;
  // measured optimum 31 : 43 of 74 (the bottom role has three layers).  The two counts are kept coprime: tile t goes from top worker
  // t % n_top to bottom worker t % n_bot, and with a common divisor g the CTAs fall into g closed groups that run in lockstep
  // (34 : 40 and 36 : 38 measured 6-10 % slower than 33 : 41 and 35 : 39).
  int n_top = (per_net * 21 + 25) / 50, n_bot = per_net - n_top;
  {
    auto gcd = [](int x, int y) { while (y) { const int r = x % y; x = y; y = r; } return x; };
    while (n_top > 1 && gcd(n_top, per_net - n_top) != 1) --n_top;
    n_bot = per_net - n_top;
  }
  if (!merged) n_top = n_bot = per_net;          // two launches: every SM runs the top role, then every SM the bottom role
  else if (const char* e = getenv("NERFCA_BWD_SPLIT")) {
    int t = 0, b = 0;
    if (sscanf(e, "%d,%d", &t, &b) == 2 && t > 0 && b > 0 && t + b <= per_net) { n_top = t; n_bot = b; }
  }
  if (n_top > a.n_tiles) n_top = (int)a.n_tiles;
  if (n_bot > a.n_tiles) n_bot = (int)a.n_tiles;
  NERFCA_REQUIRE(n_top >= 1 && n_bot >= 1, NERFCA_E_UNSUPPORTED, "tcgen05 backward needs at least two SMs per net");
  a.n_top = n_top; a.n_bot = n_bot; a.ring = (int)ring;
  NERFCA_CUDA_OK(cudaMemsetAsync(flags, 0, (size_t)n_nets * 2 * a.n_tiles * sizeof(uint32_t), st));
  // Two generations of the one-launch kernel.  tc_bwd_kernel (one tile in flight, prefetched loads, dgrad A operands in tensor memory)
  // is the faster one while the latent gradient comes through the one-hot columns (<= 12 phases; r5b: 390 vs 402-418 us);
  // tc_bwd2_kernel (two tiles in flight on one shared accumulator, mlp_tc_bwd2.cuh) has the atomic-free latent fallback and wins
  // 2.6x when that is needed (config 3, 30 phases: 0.45 vs 1.19 ms).  NERFCA_BWD_V1=1 / 0 forces one or the other.
  // (tc_bwd_kernel trapped once per ~10^4 steps in round 1 and once in a 4 000-step soak under graph replay: its top role waited
  // for bar_slot AFTER arriving on bar_mfree, see the comment there; with the two swapped it ran 2 x 40 000 + 2 x 70 000 steps clean.)
  bool fallback_lat = false;
  for (int i = 0; i < n_nets; ++i) fallback_lat = fallback_lat || (f[i]->n_latent > 0 && a.net[i].x0.onehot == 0);
  if (merged && !env_flag("NERFCA_BWD_V1", fallback_lat ? 0 : 1)) {
    if (!getenv("NERFCA_BWD_SPLIT")) {
      // per tile the top role moves ~460 KB and the bottom role ~630 KB through shared memory (what bounds both): 29 : 45 of 74
      // measured best (r3i), kept coprime (see below)
      auto gcd2 = [](int x, int y) { while (y) { const int r = x % y; x = y; y = r; } return x; };
      // (with the latent fallback the bottom role carries an extra GEMM + scatter per tile: 25 : 49 measured best at 30 phases, r5x)
      n_top = fallback_lat ? (per_net + 1) / 3 : (per_net * 2 + 2) / 5;
      while (n_top > 1 && gcd2(n_top, per_net - n_top) != 1) --n_top;
      n_bot = per_net - n_top;
      if (n_top > a.n_tiles) n_top = (int)a.n_tiles;
      if (n_bot > a.n_tiles) n_bot = (int)a.n_tiles;
      a.n_top = n_top; a.n_bot = n_bot;
    }
    a.role = 0;
    a.dbg = nullptr;
    a.dbg_cta = getenv("NERFCA_TIMELINE_CTA") ? atoi(getenv("NERFCA_TIMELINE_CTA")) : 0;
    const bool tl2 = timeline_wanted("top") || timeline_wanted("bot");     // developer aid (make TL=1)
    if (tl2) {
      NERFCA_CUDA_OK(cudaMalloc(&a.dbg, 8008 * sizeof(long long)));
      NERFCA_CUDA_OK(cudaMemsetAsync(a.dbg, 0, 8008 * sizeof(long long), st));
    }
    NERFCA_CUDA_OK(cudaFuncSetAttribute(tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD2_SMEM));
    {
      ProfScope prof(NERFCA_K_FIELD_BWD, st);
      void* kargs[] = {&a};
      NERFCA_CUDA_OK(cudaLaunchCooperativeKernel((const void*)tc_bwd2_kernel, dim3((unsigned)(n_nets * (n_top + n_bot))), dim3(BWD2_THREADS), kargs,
                                                 BWD2_SMEM, st));
      NERFCA_LAUNCH_OK();
    }
    if (tl2) return timeline_dump(a.dbg, st);
    return NERFCA_OK;
  }
  NERFCA_CUDA_OK(cudaFuncSetAttribute(tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
  NERFCA_CUDA_OK(cudaFuncSetAttribute(tc_bwd_top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TOP_SMEM));
  NERFCA_CUDA_OK(cudaFuncSetAttribute(tc_bwd_bot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BOT_SMEM));
  const bool tl_top = timeline_wanted("top"), tl_bot = timeline_wanted("bot");   // NERFCA_TIMELINE_CTA picks a CTA of that role
  a.dbg = nullptr;
  a.dbg_cta = getenv("NERFCA_TIMELINE_CTA") ? atoi(getenv("NERFCA_TIMELINE_CTA")) : 0;
  if (tl_top || tl_bot) {
    NERFCA_CUDA_OK(cudaMalloc(&a.dbg, 8008 * sizeof(long long)));
    NERFCA_CUDA_OK(cudaMemsetAsync(a.dbg, 0, 8008 * sizeof(long long), st));
    if (merged && tl_bot && !getenv("NERFCA_TIMELINE_CTA")) a.dbg_cta = n_nets * n_top;
  }
  {
    ProfScope prof(NERFCA_K_FIELD_BWD, st);
    if (merged) {
      a.role = 0;
      void* kargs[] = {&a};
      NERFCA_CUDA_OK(cudaLaunchCooperativeKernel((const void*)tc_bwd_kernel, dim3((unsigned)(n_nets * (n_top + n_bot))), dim3(BWD_THREADS), kargs,
                                                 BWD_SMEM, st));
      NERFCA_LAUNCH_OK();
    } else {
      BwdArgs b = a;
      if (!tl_top) b.dbg = nullptr;
      b.role = 1;
      tc_bwd_top_kernel<<<(unsigned)(n_nets * n_top), BWD_THREADS, TOP_SMEM, st>>>(b);
      NERFCA_LAUNCH_OK();
      b = a;
      if (!tl_bot) b.dbg = nullptr;
      b.role = 2;
      tc_bwd_bot_kernel<<<(unsigned)(n_nets * n_bot), BWD_THREADS, BOT_SMEM, st>>>(b);
      NERFCA_LAUNCH_OK();
    }
  }
  if (tl_top || tl_bot) return timeline_dump(a.dbg, st);
  return NERFCA_OK;
}

// ---- parity / debug entry: the first-layer input tile exactly as the tensor-core kernels build it ------------------------------
struct GlobalSink {
  uint16_t* row_out;   // this row's kpad0 bf16 values
  __device__ __forceinline__ void chunk(int c, const float* v) const { *reinterpret_cast<uint4*>(row_out + c * 8) = pack_chunk(v); }
};
__global__ void x0_debug_kernel(X0Desc xd, SampleSrc src, uint16_t* out) {
  const long long p = (long long)blockIdx.x * TILE_M + threadIdx.x;
  const RowIn in = fetch_row(xd, src, p, p < src.n_points);
  const GlobalSink sink{out + (size_t)p * xd.kpad0};
  emit_x0_row(xd, in, xd.enc.band_weight, xd.enc.latents, sink, 0);
  emit_x0_row(xd, in, xd.enc.band_weight, xd.enc.latents, sink, 1);
}
int tc_debug_x0(const nerfca_field_t& f, const nerfca_samples_t& s, int onehot, uint16_t* out, int* kpad0_out, cudaStream_t st) {
  const NetDims d = net_dims(f);
  if (kpad0_out) *kpad0_out = d.kpad0;
  if (!out || s.n_points == 0) return NERFCA_OK;
  X0Desc xd = make_x0(f, d);
  if (onehot && f.n_latent > 0 && f.n_phases <= d.kpad0 - d.in_dim - 1) xd.onehot = f.n_phases;
  x0_debug_kernel<<<(unsigned)n_tiles_of(s.n_points), TILE_M, 0, st>>>(xd, make_src(s), out);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

// ---- single-field entry points (module-level CPPN.forward / Temporal.forward_composite and their autograd) ----------
size_t tc_stash_bytes(const nerfca_field_t& f, long long P) { (void)f; return tc_stash_bytes_n(1, P); }
size_t tc_workspace_bytes(const nerfca_field_t& f, long long P, int backward) {
  const nerfca_field_t* fs[1] = {&f};
  return tc_workspace_bytes_n(fs, 1, P, backward);
}
int tc_field_forward(const nerfca_field_t& f, const nerfca_samples_t& s, float* raw_out, void* stash, void* workspace,
                     cudaStream_t st) {
  const nerfca_field_t* fs[1] = {&f};
  float* outs[1] = {raw_out};
  return tc_fields_forward(fs, 1, s, outs, stash, workspace, 1, nullptr, st, nullptr, 0);
}
int tc_field_backward(const nerfca_field_t& f, const nerfca_samples_t& s, const float* d_raw, const void* stash, void* workspace,
                      const nerfca_field_grads_t& gr, cudaStream_t st) {
  const nerfca_field_t* fs[1] = {&f};
  const float* ds[1] = {d_raw};
  const nerfca_field_grads_t* gs[1] = {&gr};
  return tc_fields_backward(fs, 1, s, ds, stash, workspace, 1, gs, st);
}

}  // namespace nerfca
