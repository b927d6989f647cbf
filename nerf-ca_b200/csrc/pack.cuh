// Packed parameter block of one field as the tcgen05 kernels read it (bf16 operand tiles + fp32 biases), and the table that lets
// the optimizer kernels keep it current: every parameter the update writes is also stored, converted, at its packed position, so
// the training step needs no separate re-pack launch.
//
//   [W0: kpad0 * 256 B][W1..W4: 4 x 32768 B][w_out tile: 4096 B][fp32: bias[5][128], w_out[128], b_out, pad]
//   weights: tile-canonical K-major bytes   byte(row n, k) = (k / 8) * 2048 + n * 16 + (k % 8) * 2   (tc_common.cuh)
//   layer 0 is padded to kpad0 = roundup16(in_dim + 1) input columns, column in_dim carries the layer-0 bias
//   w_out tile: [16 rows x 128] bf16, 256 B per 8-wide K chunk; row 0 = hi part of w_out, row 1 = lo part (w - hi), rest 0
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace nerfca {

constexpr int PACK_H = 128;                      // hidden width the packed layout is built for
constexpr int PACK_N_RELU = 5;                   // layers with a ReLU (first + 4 hidden)
constexpr uint32_t PACK_TILE_BYTES = 32768;      // one [128 x 128] bf16 tile

struct NetDims {
  int in_dim, enc_dim, kpad0;
  uint32_t w0_bytes, w_bytes, wout_off, f32_off, pack_bytes;
};
__host__ __device__ inline uint32_t f32_block_floats() { return PACK_N_RELU * 128 + 128 + 32; }
inline NetDims net_dims(const nerfca_field_t& f) {
  NetDims d;
  d.in_dim = in_dim_of(f);
  d.enc_dim = enc_dim_of(f);
  d.kpad0 = (d.in_dim + 1 + 15) / 16 * 16;
  d.w0_bytes = (uint32_t)d.kpad0 * 256u;
  d.w_bytes = d.w0_bytes + 4u * PACK_TILE_BYTES;
  d.wout_off = d.w_bytes;
  d.f32_off = d.w_bytes + 4096u;
  d.pack_bytes = d.f32_off + f32_block_floats() * 4u;
  return d;
}
// packed blocks of the nets sit back to back at the head of the step workspace, each rounded up to 256 B
inline size_t pack_stride(const NetDims& d) { return ((size_t)d.pack_bytes + 255) & ~(size_t)255; }

// ---- flat parameter buffer -> packed positions -----------------------------------------------------------------------
enum { SEG_WEIGHT = 0, SEG_BIAS = 1, SEG_WOUT = 2, SEG_BOUT = 3 };
struct RepackSeg {
  long long begin;       // first float of the tensor inside the flat parameter buffer
  int count;             // elements
  short kind, layer;     // SEG_*, layer index (0 .. 4) for weights / biases
  short K, net;          // row length of a weight tensor; which packed block
};
struct RepackNet {
  uint8_t* out;          // packed block of the net
  int in_dim, kpad0;
  uint32_t w0_bytes, wout_off, f32_off;
};
constexpr int REPACK_MAX_SEGS = 2 * (2 * (PACK_N_RELU + 1));
struct RepackTable {
  int n_segs;            // 0: nothing to keep current
  RepackNet net[2];
  RepackSeg seg[REPACK_MAX_SEGS];
};

__device__ __forceinline__ void repack_bf16(uint8_t* base, uint32_t elem, float v) {
  reinterpret_cast<__nv_bfloat16*>(base)[elem] = __float2bfloat16_rn(v);
}
// element e of segment s has just been updated to `v`
__device__ __forceinline__ void repack_store(const RepackTable& t, const RepackSeg& s, int e, float v) {
  const RepackNet& n = t.net[s.net];
  float* f32 = reinterpret_cast<float*>(n.out + n.f32_off);
  if (s.kind == SEG_WEIGHT) {
    const int row = e / s.K, k = e - row * s.K;
    uint8_t* base = n.out + (s.layer == 0 ? 0u : n.w0_bytes + (uint32_t)(s.layer - 1) * PACK_TILE_BYTES);
    repack_bf16(base, (uint32_t)(k >> 3) * 1024u + (uint32_t)row * 8u + (uint32_t)(k & 7), v);
  } else if (s.kind == SEG_BIAS) {
    f32[s.layer * 128 + e] = v;
    if (s.layer == 0) repack_bf16(n.out, (uint32_t)(n.in_dim >> 3) * 1024u + (uint32_t)e * 8u + (uint32_t)(n.in_dim & 7), v);
  } else if (s.kind == SEG_WOUT) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(n.out + n.wout_off);
    const uint32_t at = (uint32_t)(e >> 3) * 128u + (uint32_t)(e & 7);       // row 0; row r adds 8 r
    tile[at] = hi;
    tile[at + 8] = __float2bfloat16_rn(v - __bfloat162float(hi));
    f32[PACK_N_RELU * 128 + e] = v;
  } else {
    f32[PACK_N_RELU * 128 + 128] = v;
  }
}
// flat index i (and the three behind it: tensors start on 4-float boundaries, so a float4 group lies in one tensor or in padding)
__device__ __forceinline__ void repack_group(const RepackTable& t, long long i, const float (&v)[4]) {
  for (int k = 0; k < t.n_segs; ++k) {
    const RepackSeg& s = t.seg[k];
    if (i < s.begin || i >= s.begin + s.count) continue;
    const int e0 = (int)(i - s.begin);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (e0 + j < s.count) repack_store(t, s, e0 + j, v[j]);
    return;
  }
}

// Builds the table for the fields whose parameter tensors lie inside [params, params + n); returns an empty table (n_segs = 0)
// with rc != 0 when a field is not of the packed shape.  (mlp_tc.cu)
int make_repack_table(const nerfca_field_t* const* fields, int n_nets, void* workspace, const float* params, long long n,
                      RepackTable* out);

}  // namespace nerfca
