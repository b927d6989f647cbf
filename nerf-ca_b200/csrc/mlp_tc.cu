// bf16 tensor-core path of the two coordinate MLPs on tcgen05 / TMEM (sm_100a).
//
// Forward (A5-A7): one persistent CTA per SM evaluates ONE field for a stream of 128-sample tiles.
//   * all layer weights (bf16, tile-canonical K-major) stay resident in shared memory for the CTA's lifetime
//     (loaded once with 1-D bulk async copies), biases / output weights in fp32;
//   * warp 8 lane 0 issues tcgen05.mma (M=128 samples, N=128 features, K=16 per instruction), accumulating each
//     layer's [128 x 128] fp32 pre-activation in TMEM;
//   * two epilogue warpgroups (warps 0-3 and 4-7) each own one of two in-flight tiles ("slots"): they build the
//     layer-0 input (sample point + positional encoding, in registers), and after every layer read the accumulator
//     with tcgen05.ld, add bias, ReLU, round to bf16 and write the next layer's A operand back to shared memory;
//     the last layer's 128 -> 1 projection is a per-thread dot product (thread == sample row), so no MMA with N = 1;
//   * the two slots ping-pong: while one tile's epilogue runs on the CUDA cores the other tile's layer runs on the
//     tensor core.  mbarriers: act_full[slot] (128 epilogue arrivals) -> MMA, acc_full[slot] (tcgen05.commit) -> epilogue.
//   * training: every layer's bf16 input/activation tile is also written to the stash in tile-canonical layout so
//     the backward kernels can bulk-copy it straight back into shared memory as a UMMA operand.
#include "common.cuh"
#include "tc_common.cuh"

namespace nerfca {

using namespace tc;

constexpr int TC_H = 128;            // hidden width the tensor-core kernels are built for
constexpr int TC_THREADS = 288;      // 2 epilogue warpgroups + 1 MMA / control warp
constexpr int TC_MAX_RELU_LAYERS = 5;

struct TcDims {
  int in_dim, kpad0, n_relu;         // n_relu = n_hidden + 1 layers that end in ReLU
  size_t w_bytes;                    // packed bf16 weights of the ReLU layers
  size_t x0_bytes;                   // one tile of layer-0 input  (kpad0 * 256)
  size_t tile_stash_bytes;           // x0 + n_relu activation tiles
};
static TcDims tc_dims(const nerfca_field_t& f) {
  TcDims d;
  d.in_dim = in_dim_of(f);
  d.kpad0 = (d.in_dim + 15) / 16 * 16;
  d.n_relu = f.n_hidden + 1;
  d.w_bytes = (size_t)d.kpad0 * 256 + (size_t)f.n_hidden * 32768;
  d.x0_bytes = (size_t)d.kpad0 * 256;
  d.tile_stash_bytes = d.x0_bytes + (size_t)d.n_relu * 32768;
  return d;
}
// workspace: [packed weights][bias: n_relu * 128 f32][w_out: 128 f32][b_out: 4 f32]
static size_t tc_param_bytes(const TcDims& d) { return d.w_bytes + ((size_t)d.n_relu * 128 + 128 + 4) * sizeof(float); }

int tc_supported(const nerfca_field_t& f) {
  NERFCA_REQUIRE(f.hidden == TC_H, NERFCA_E_UNSUPPORTED, "the tcgen05 path is built for hidden == 128 (use precision fp32)");
  NERFCA_REQUIRE(f.n_hidden + 1 <= TC_MAX_RELU_LAYERS, NERFCA_E_UNSUPPORTED, "tcgen05 path: at most 4 hidden layers fit in shared memory");
  NERFCA_REQUIRE(in_dim_of(f) <= 128, NERFCA_E_UNSUPPORTED, "tcgen05 path: first-layer input wider than 128");
  return NERFCA_OK;
}

// ---- parameter packing: fp32 nn.Linear tensors -> bf16 tile-canonical + fp32 bias block ------------------------
struct PackArgs {
  const float* w[NERFCA_MAX_LAYERS];
  const float* b[NERFCA_MAX_LAYERS];
  int in_dim, kpad0, n_relu;
};
__global__ void pack_params_kernel(PackArgs a, uint8_t* __restrict__ out, size_t w_bytes) {
  const int total_w = a.kpad0 * 128 + (a.n_relu - 1) * 128 * 128;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < total_w) {
    int l, e;
    if (tid < a.kpad0 * 128) { l = 0; e = tid; }
    else { l = 1 + (tid - a.kpad0 * 128) / (128 * 128); e = (tid - a.kpad0 * 128) % (128 * 128); }
    const int K = (l == 0) ? a.in_dim : 128;
    // e enumerates the packed layout: chunk-major, then output row n, then position in chunk
    const int chunk = e / (128 * 8), n = (e / 8) % 128, kk = e % 8;
    const int k = chunk * 8 + kk;
    const float v = (k < K) ? __ldg(a.w[l] + (size_t)n * K + k) : 0.f;
    const size_t base = (l == 0) ? 0 : (size_t)a.kpad0 * 256 + (size_t)(l - 1) * 32768;
    reinterpret_cast<__nv_bfloat16*>(out + base)[e] = __float2bfloat16_rn(v);
  }
  float* fb = reinterpret_cast<float*>(out + w_bytes);
  const int nb = a.n_relu * 128 + 128 + 4;
  if (tid < nb) {
    float v = 0.f;
    if (tid < a.n_relu * 128) { const int l = tid / 128; v = a.b[l] ? __ldg(a.b[l] + tid % 128) : 0.f; }
    else if (tid < a.n_relu * 128 + 128) v = __ldg(a.w[a.n_relu] + (tid - a.n_relu * 128));
    else if (tid == a.n_relu * 128 + 128) v = a.b[a.n_relu] ? __ldg(a.b[a.n_relu]) : 0.f;
    fb[tid] = v;
  }
}

static int pack_params(const nerfca_field_t& f, const TcDims& d, void* dst, cudaStream_t st) {
  PackArgs a;
  for (int l = 0; l < NERFCA_MAX_LAYERS; ++l) { a.w[l] = f.weight[l]; a.b[l] = f.bias[l]; }
  a.in_dim = d.in_dim; a.kpad0 = d.kpad0; a.n_relu = d.n_relu;
  const int total = d.kpad0 * 128 + f.n_hidden * 128 * 128;
  pack_params_kernel<<<div_up(total, 256), 256, 0, st>>>(a, (uint8_t*)dst, d.w_bytes);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

// ---- forward kernel ---------------------------------------------------------------------------------------------
struct FwdArgs {
  SampleSrc src;
  EncDesc enc;
  const uint8_t* params;   // packed weights + fp32 block
  float* raw_out;
  uint8_t* stash;          // or null
  long long n_tiles;
  int kpad0, n_relu;
  uint32_t w_bytes, tile_stash_bytes;
};

// dynamic smem map (bytes):  [weights][act slot 0: 32768][act slot 1: 32768][fp32 block][barriers / tmem ptr]
__global__ void __launch_bounds__(TC_THREADS, 1) tc_forward_kernel(FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* s_w = smem;
  uint8_t* s_act = s_w + a.w_bytes;
  float* s_f = reinterpret_cast<float*>(s_act + 2 * 32768);
  const int n_f = a.n_relu * 128 + 128 + 4;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_f + ((n_f + 3) & ~3));
  // s_bar[0] weights, [1..2] act_full, [3..4] acc_full, then the TMEM base pointer
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 5);
  const uint32_t bar_w = smem_u32(s_bar), bar_act0 = smem_u32(s_bar + 1), bar_acc0 = smem_u32(s_bar + 3);

  if (warp == 8) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_act0, 128); mbar_init(bar_act0 + 8, 128);
      mbar_init(bar_acc0, 1); mbar_init(bar_acc0 + 8, 1);
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), 256);
  }
  for (int i = threadIdx.x; i < n_f; i += blockDim.x) s_f[i] = __ldg(reinterpret_cast<const float*>(a.params + a.w_bytes) + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const long long n_my = (a.n_tiles > (long long)blockIdx.x) ? (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const float* s_wout = s_f + a.n_relu * 128;
  const float b_out = s_f[a.n_relu * 128 + 128];

  if (warp == 8) {
    // ================= control warp: weight load + MMA issue =================
    if (lane == 0) {
      mbar_expect_tx(bar_w, a.w_bytes);
      for (uint32_t off = 0; off < a.w_bytes; off += 32768) {
        const uint32_t n = (a.w_bytes - off < 32768u) ? a.w_bytes - off : 32768u;
        bulk_g2s(smem_u32(s_w + off), a.params + off, n, bar_w);
      }
      mbar_wait(bar_w, 0);
      constexpr uint32_t idesc = instr_desc(128, 128, 0, 0);
      uint32_t ph_act[2] = {0, 0};
      for (long long i0 = 0; i0 < n_my; i0 += 2) {
        const int nslots = (n_my - i0 >= 2) ? 2 : 1;
        for (int l = 0; l < a.n_relu; ++l) {
          const int ksteps = (l == 0 ? a.kpad0 : 128) / 16;
          const uint32_t wl = smem_u32(s_w) + (l == 0 ? 0u : (uint32_t)a.kpad0 * 256u + (uint32_t)(l - 1) * 32768u);
          for (int s = 0; s < nslots; ++s) {
            mbar_wait(bar_act0 + 8 * s, ph_act[s]);
            ph_act[s] ^= 1;
            tc_fence_after();
            const uint32_t act = smem_u32(s_act + s * 32768);
            for (int kk = 0; kk < ksteps; ++kk)
              umma_ss(tmem + s * 128, desc_kmajor(act, kk), desc_kmajor(wl, kk), idesc, kk > 0);
            umma_commit(bar_acc0 + 8 * s);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warpgroups =================
    const int g = warp >> 2;                       // slot owned by this warpgroup
    const int row = (warp & 3) * 32 + lane;        // tile row == TMEM lane
    uint8_t* act = s_act + g * 32768;
    const uint32_t t_acc = tmem + g * 128 + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ph_acc = 0;
    for (long long i = g; i < n_my; i += 2) {
      const long long tile = blockIdx.x + i * gridDim.x;
      const long long p = tile * TILE_M + row;     // local sample index
      const bool valid = p < a.src.n_points;
      uint8_t* st_tile = a.stash ? a.stash + (size_t)tile * a.tile_stash_bytes : nullptr;
      // ---- layer-0 input: point + encoding (+ latent) -> bf16 row of the A tile
      {
        float x = 0.f, y = 0.f, z = 0.f;
        int phase = 0;
        if (valid) {
          load_point(a.src, p, x, y, z);
          if (a.enc.n_latent > 0) phase = load_phase(a.src, p);
          if (phase < 0 || phase >= a.enc.n_phases) phase = 0;
        }
        for (int c = 0; c < a.kpad0 / 8; ++c) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int f = c * 8 + e;
            v[e] = (valid && f < a.enc.in_dim) ? enc_feature(a.enc, f, x, y, z, phase) : 0.f;
          }
          const uint4 q = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          *reinterpret_cast<uint4*>(act + c * CHUNK_BYTES + row * 16) = q;
          if (st_tile) *reinterpret_cast<uint4*>(st_tile + c * CHUNK_BYTES + row * 16) = q;
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_act0 + 8 * g);

      for (int l = 0; l < a.n_relu; ++l) {
        mbar_wait(bar_acc0 + 8 * g, ph_acc);
        ph_acc ^= 1;
        tc_fence_after();
        const bool last = (l == a.n_relu - 1);
        const float* bias = s_f + l * 128;
        uint8_t* st_l = st_tile ? st_tile + (size_t)a.kpad0 * 256 + (size_t)l * 32768 : nullptr;
        float dot = 0.f;
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t v[32];
          tmem_ld32(t_acc + cb * 32, v);
          tmem_ld_wait();
          float h[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) h[j] = fmaxf(__uint_as_float(v[j]) + bias[cb * 32 + j], 0.f);
          if (last) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dot = fmaf(h[j], s_wout[cb * 32 + j], dot);
          }
          if (!last || st_l) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint4 q = make_uint4(pack_bf16x2(h[q4 * 8 + 0], h[q4 * 8 + 1]), pack_bf16x2(h[q4 * 8 + 2], h[q4 * 8 + 3]),
                                         pack_bf16x2(h[q4 * 8 + 4], h[q4 * 8 + 5]), pack_bf16x2(h[q4 * 8 + 6], h[q4 * 8 + 7]));
              const int off = (cb * 4 + q4) * CHUNK_BYTES + row * 16;
              if (!last) *reinterpret_cast<uint4*>(act + off) = q;
              if (st_l) *reinterpret_cast<uint4*>(st_l + off) = q;
            }
          }
        }
        tc_fence_before();
        if (!last) {
          fence_proxy_async();
          mbar_arrive(bar_act0 + 8 * g);
        } else if (valid) {
          a.raw_out[p] = dot + b_out;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

static size_t fwd_smem_bytes(const TcDims& d) {
  const int n_f = d.n_relu * 128 + 128 + 4;
  return d.w_bytes + 2 * 32768 + (size_t)((n_f + 3) & ~3) * sizeof(float) + 8 * sizeof(uint64_t);
}

size_t tc_stash_bytes(const nerfca_field_t& f, long long P) {
  const TcDims d = tc_dims(f);
  return (size_t)((P + TILE_M - 1) / TILE_M) * d.tile_stash_bytes;
}

// backward workspace: [params block][dZ hand-off between layer-group passes: n_tiles * 32 KB]
size_t tc_workspace_bytes(const nerfca_field_t& f, long long P, int backward) {
  const TcDims d = tc_dims(f);
  size_t n = (tc_param_bytes(d) + 255) & ~(size_t)255;
  if (backward) n += (size_t)((P + TILE_M - 1) / TILE_M) * 32768;
  return n;
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

int tc_field_forward(const nerfca_field_t& f, const nerfca_samples_t& s, float* raw_out, void* stash, void* workspace,
                     cudaStream_t st) {
  const TcDims d = tc_dims(f);
  int rc = pack_params(f, d, workspace, st);
  if (rc) return rc;
  FwdArgs a;
  a.src = make_src(s);
  a.enc = make_enc(f);
  a.params = (const uint8_t*)workspace;
  a.raw_out = raw_out;
  a.stash = (uint8_t*)stash;
  a.n_tiles = (s.n_points + TILE_M - 1) / TILE_M;
  a.kpad0 = d.kpad0; a.n_relu = d.n_relu;
  a.w_bytes = (uint32_t)d.w_bytes; a.tile_stash_bytes = (uint32_t)d.tile_stash_bytes;
  const size_t smem = fwd_smem_bytes(d);
  NERFCA_REQUIRE(smem <= 227 * 1024, NERFCA_E_UNSUPPORTED, "field does not fit the forward kernel's shared memory");
  NERFCA_CUDA_OK(cudaFuncSetAttribute(tc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long grid = a.n_tiles < sm_count() ? a.n_tiles : sm_count();
  tc_forward_kernel<<<(unsigned)grid, TC_THREADS, smem, st>>>(a);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

// ---- backward kernel -----------------------------------------------------------------------------------------------
// One launch handles a group of at most two consecutive layers (hi, hi-1) for a stream of tiles; the weight-gradient
// accumulators of those layers stay in TMEM for the CTA's whole lifetime and are flushed once with atomics.
//   per tile and per layer l of the group (top first):
//     dgrad   accD[128 x 128]   = dZ_l (K-major A)        x W_l  (MN-major B)      -> epilogue: * 1[H_{l-1} > 0] -> dZ_{l-1}
//     wgrad   accW_l[out x in] += dZ_l^T (MN-major A)     x H_{l-1} (MN-major B)       (H_{-1} = encoded input X0)
//     bgrad   accB_l[out x 16] += dZ_l^T (MN-major A)     x side tile (column 2 == 1)
//   top group only: dZ_{L-1} = d_raw * w_out * 1[H_{L-1} > 0] on the CUDA cores, and
//     wout    accO[feat x 16]  += H_{L-1}^T (MN-major A)  x side tile (columns 0,1 = bf16 hi / lo split of d_raw)
//   bottom layer of a Temporal field: latent dgrad accD[128 x 16/32] = dZ_0 x W_0[:, latent columns] -> scatter by phase.
// Groups hand dZ over through global memory in tile-canonical layout (bulk-copied back into smem by the next launch).
struct BwdArgs {
  SampleSrc src;
  const uint8_t* params;     // packed weights + fp32 block (w_out at fp32 offset n_relu*128)
  const uint8_t* stash;
  uint8_t* handoff;          // [n_tiles][32768]
  const float* d_raw;
  float* g_w[2];             // fp32 gradient of weight[l_hi], weight[l_hi-1]
  float* g_b[2];             // may be null
  float* g_wout; float* g_bout; float* g_lat;
  long long n_tiles;
  int kpad0, n_relu, in_dim, enc_dim, n_latent, n_phases;
  int l_hi, n_layers;        // layers l_hi, l_hi-1 (n_layers in {1,2})
  int from_raw;              // l_hi == n_relu-1
  uint32_t w_bytes, tile_stash_bytes;
};

constexpr int BW_COL_D = 0, BW_COL_W0 = 128, BW_COL_W1 = 256, BW_COL_B0 = 384, BW_COL_B1 = 400, BW_COL_O = 416;

__global__ void __launch_bounds__(TC_THREADS, 1) tc_backward_kernel(BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_lat = a.n_latent > 0;
  const int l_lo = a.l_hi - a.n_layers + 1;
  // smem map: [W(l_hi)][W(l_hi-1)] (32 KB each; layer 0: kpad0*256)  [P][Q][hbuf0][hbuf1][side 4 KB][w_out 512 B][lat acc][bars]
  uint32_t w_off[2], w_len[2];
  uint32_t cur = 0;
  for (int j = 0; j < 2; ++j) {
    const int l = a.l_hi - j;
    const bool need = j < a.n_layers && (l > 0 || has_lat);
    w_off[j] = cur;
    w_len[j] = need ? (l == 0 ? (uint32_t)a.kpad0 * 256u : 32768u) : 0u;
    cur += w_len[j];
  }
  uint8_t* s_w = smem;
  uint8_t* s_dz = smem + cur;            // P = s_dz, Q = s_dz + 32768
  uint8_t* s_h = s_dz + 2 * 32768;       // hbuf0, hbuf1
  uint8_t* s_side = s_h + 2 * 32768;
  float* s_wout = reinterpret_cast<float*>(s_side + 4096);
  float* s_lat = s_wout + 128;
  const int n_lat_acc = (has_lat && l_lo == 0) ? a.n_phases * a.n_latent : 0;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_lat + ((n_lat_acc + 3) & ~3));
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 6);
  const uint32_t bar_w = smem_u32(s_bar), bar_load = bar_w + 8, bar_dz = bar_w + 16, bar_d = bar_w + 24, bar_done = bar_w + 32,
                 bar_free = bar_w + 40;

  if (warp == 8) {
    if (lane == 0) {
      mbar_init(bar_w, 1); mbar_init(bar_load, 1); mbar_init(bar_dz, 256); mbar_init(bar_d, 1); mbar_init(bar_done, 1);
      mbar_init(bar_free, 256);
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), 512);
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x)
    s_wout[i] = __ldg(reinterpret_cast<const float*>(a.params + a.w_bytes) + a.n_relu * 128 + i);
  for (int i = threadIdx.x; i < n_lat_acc; i += blockDim.x) s_lat[i] = 0.f;
  for (int i = threadIdx.x; i < 4096 / 16; i += blockDim.x) reinterpret_cast<uint4*>(s_side)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const long long n_my = (a.n_tiles > (long long)blockIdx.x) ? (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int lat_c0 = a.enc_dim / 8;                                        // first chunk holding latent columns
  const int lat_n = ((a.enc_dim % 8 + a.n_latent + 15) / 16) * 16;         // MMA N covering them

  if (warp == 8) {
    if (lane == 0) {
      // ---- resident weights of the group
      uint32_t wtot = w_len[0] + w_len[1];
      if (wtot) {
        mbar_expect_tx(bar_w, wtot);
        for (int j = 0; j < 2; ++j) {
          if (!w_len[j]) continue;
          const int l = a.l_hi - j;
          const uint8_t* src = a.params + (l == 0 ? 0 : (size_t)a.kpad0 * 256 + (size_t)(l - 1) * 32768);
          bulk_g2s(smem_u32(s_w + w_off[j]), src, w_len[j], bar_w);
        }
        mbar_wait(bar_w, 0);
      }
      uint32_t ph_dz = 0, ph_done = 0, ph_free = 0, ph_load_c = 0;
      for (long long i = 0; i < n_my; ++i) {
        const long long tile = blockIdx.x + i * gridDim.x;
        const uint8_t* st_tile = a.stash + (size_t)tile * a.tile_stash_bytes;
        if (i > 0) {   // every MMA and every epilogue read of the previous tile's buffers has finished
          mbar_wait(bar_done, ph_done); ph_done ^= 1;
          mbar_wait(bar_free, ph_free); ph_free ^= 1;
        }
        // ---- tile loads
        uint32_t bytes = 32768;
        for (int j = 0; j < a.n_layers; ++j) bytes += (a.l_hi - j > 0) ? 32768u : (uint32_t)a.kpad0 * 256u;
        mbar_expect_tx(bar_load, bytes);
        if (a.from_raw) bulk_g2s(smem_u32(s_dz + 32768), st_tile + (size_t)a.kpad0 * 256 + (size_t)a.l_hi * 32768, 32768, bar_load);
        else            bulk_g2s(smem_u32(s_dz), a.handoff + (size_t)tile * 32768, 32768, bar_load);
        for (int j = 0; j < a.n_layers; ++j) {
          const int l = a.l_hi - j;
          if (l > 0) bulk_g2s(smem_u32(s_h + j * 32768), st_tile + (size_t)a.kpad0 * 256 + (size_t)(l - 1) * 32768, 32768, bar_load);
          else       bulk_g2s(smem_u32(s_h + j * 32768), st_tile, (uint32_t)a.kpad0 * 256u, bar_load);
        }
        mbar_wait(bar_load, ph_load_c); ph_load_c ^= 1;   // the issuing thread observes the bulk copies itself as well
        const uint32_t acc_first = (i > 0) ? 1u : 0u;
        for (int j = 0; j < a.n_layers; ++j) {
          const int l = a.l_hi - j;
          const uint32_t dz = smem_u32(s_dz + (j & 1) * 32768), other = smem_u32(s_dz + ((j & 1) ^ 1) * 32768);
          const uint32_t hb = smem_u32(s_h + j * 32768), side = smem_u32(s_side), wl = smem_u32(s_w + w_off[j]);
          mbar_wait(bar_dz, ph_dz); ph_dz ^= 1;    // dZ_l (and the side tile) are in shared memory
          tc_fence_after();
          if (j == 0 && a.from_raw)
            for (int kk = 0; kk < 8; ++kk)
              umma_ss(tmem + BW_COL_O, desc_mnmajor(other, kk), desc_mnmajor(side, kk), instr_desc(128, 16, 1, 1), acc_first | (kk > 0));
          if (l > 0) {
            for (int kk = 0; kk < 8; ++kk)
              umma_ss(tmem + BW_COL_D, desc_kmajor(dz, kk), desc_mnmajor(wl, kk), instr_desc(128, 128, 0, 1), kk > 0);
            umma_commit(bar_d);
          } else if (has_lat) {
            for (int kk = 0; kk < 8; ++kk)
              umma_ss(tmem + BW_COL_D, desc_kmajor(dz, kk), desc_mnmajor(wl + lat_c0 * CHUNK_BYTES, kk), instr_desc(128, lat_n, 0, 1), kk > 0);
            umma_commit(bar_d);
          }
          const int n_in = (l > 0) ? 128 : a.kpad0;
          for (int kk = 0; kk < 8; ++kk)
            umma_ss(tmem + (j == 0 ? BW_COL_W0 : BW_COL_W1), desc_mnmajor(dz, kk), desc_mnmajor(hb, kk), instr_desc(128, n_in, 1, 1),
                    acc_first | (kk > 0));
          for (int kk = 0; kk < 8; ++kk)
            umma_ss(tmem + (j == 0 ? BW_COL_B0 : BW_COL_B1), desc_mnmajor(dz, kk), desc_mnmajor(side, kk), instr_desc(128, 16, 1, 1),
                    acc_first | (kk > 0));
        }
        umma_commit(bar_done);
      }
    }
    __syncwarp();
  } else {
    // ================= 8 epilogue warps: thread = (row, column half) =================
    const int row = (warp & 3) * 32 + lane, ch = warp >> 2;
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ph_load = 0, ph_d = 0;
    for (long long i = 0; i < n_my; ++i) {
      const long long tile = blockIdx.x + i * gridDim.x;
      const long long p = tile * TILE_M + row;
      const bool valid = p < a.src.n_points;
      mbar_wait(bar_load, ph_load); ph_load ^= 1;
      float g = 0.f;
      if (a.from_raw) {
        // dZ_top = d_raw * w_out * 1[H_top > 0]   (H_top sits in Q)
        g = valid ? __ldg(a.d_raw + p) : 0.f;
        const uint8_t* hq = s_dz + 32768;
        for (int c = ch * 8; c < ch * 8 + 8; ++c) {
          const uint4 hv = *reinterpret_cast<const uint4*>(hq + c * CHUNK_BYTES + row * 16);
          const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lo = (hw[e] & 0x7FFFu) ? g * s_wout[c * 8 + 2 * e] : 0.f;         // bf16 h > 0  <=>  magnitude bits set (h >= 0)
            const float hi = (hw[e] & 0x7FFF0000u) ? g * s_wout[c * 8 + 2 * e + 1] : 0.f;
            o[e] = pack_bf16x2(lo, hi);
          }
          *reinterpret_cast<uint4*>(s_dz + c * CHUNK_BYTES + row * 16) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        if (ch == 0 && a.g_bout) {
          float sum = g;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          if (lane == 0) atomicAdd(a.g_bout, sum);
        }
      }
      if (ch == 0) {   // side tile row: [d_hi, d_lo, 1, 0, ...]
        const __nv_bfloat16 dh = __float2bfloat16_rn(g);
        const __nv_bfloat16 dl = __float2bfloat16_rn(g - __bfloat162float(dh));
        const uint32_t w0 = (uint32_t)__bfloat16_as_ushort(dh) | ((uint32_t)__bfloat16_as_ushort(dl) << 16);
        *reinterpret_cast<uint4*>(s_side + row * 16) = make_uint4(w0, 0x00003F80u, 0u, 0u);   // 0x3F80 = bf16(1.0)
      }
      fence_proxy_async();
      mbar_arrive(bar_dz);

      for (int j = 0; j < a.n_layers; ++j) {
        const int l = a.l_hi - j;
        if (!(l > 0 || has_lat)) continue;
        mbar_wait(bar_d, ph_d); ph_d ^= 1;
        tc_fence_after();
        if (l > 0) {
          const uint8_t* hprev = s_h + j * 32768;
          uint8_t* out_s = s_dz + ((j & 1) ^ 1) * 32768;
          const bool to_smem = (j + 1 < a.n_layers);
          uint8_t* out_g = a.handoff + (size_t)tile * 32768;
#pragma unroll 1
          for (int cb = 0; cb < 2; ++cb) {
            uint32_t v[32];
            tmem_ld32(t_lane + BW_COL_D + ch * 64 + cb * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int c = ch * 8 + cb * 4 + q4;
              const uint4 hv = *reinterpret_cast<const uint4*>(hprev + c * CHUNK_BYTES + row * 16);
              const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float lo = (hw[e] & 0x7FFFu) ? __uint_as_float(v[q4 * 8 + 2 * e]) : 0.f;
                const float hi = (hw[e] & 0x7FFF0000u) ? __uint_as_float(v[q4 * 8 + 2 * e + 1]) : 0.f;
                o[e] = pack_bf16x2(lo, hi);
              }
              const uint4 q = make_uint4(o[0], o[1], o[2], o[3]);
              if (to_smem) *reinterpret_cast<uint4*>(out_s + c * CHUNK_BYTES + row * 16) = q;
              else         *reinterpret_cast<uint4*>(out_g + c * CHUNK_BYTES + row * 16) = q;
            }
          }
          tc_fence_before();
          if (to_smem) {
            fence_proxy_async();
            mbar_arrive(bar_dz);
          }
        } else {
          // latent gradient: columns [enc_dim, enc_dim + T) of dX0 live at accD columns enc_dim - 8*lat_c0 + t
          if (ch == 0) {
            uint32_t v[16];
            const int phase = valid ? load_phase(a.src, p) : -1;
            for (int c0 = 0; c0 < lat_n; c0 += 16) {
              tmem_ld16(t_lane + BW_COL_D + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int t = c0 + e - (a.enc_dim - 8 * lat_c0);
                if (t >= 0 && t < a.n_latent && phase >= 0 && phase < a.n_phases)
                  atomicAdd(&s_lat[phase * a.n_latent + t], __uint_as_float(v[e]));
              }
            }
          }
          tc_fence_before();
        }
      }
      mbar_arrive(bar_free);
    }
    // ---- all tiles issued: wait for the last MMAs, then flush the TMEM-resident accumulators
    if (n_my > 0) {
      mbar_wait(bar_done, (uint32_t)((n_my - 1) & 1));
      tc_fence_after();
      for (int j = 0; j < a.n_layers; ++j) {
        const int l = a.l_hi - j;
        const int n_in = (l > 0) ? 128 : a.kpad0, k_in = (l > 0) ? 128 : a.in_dim;
        float* gw = a.g_w[j];
        for (int c0 = ch * 64; c0 < n_in && c0 < ch * 64 + 64; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_lane + (j == 0 ? BW_COL_W0 : BW_COL_W1) + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c0 + e < k_in) atomicAdd(gw + (size_t)row * k_in + c0 + e, __uint_as_float(v[e]));
        }
        if (ch == 0) {
          uint32_t v[16];
          tmem_ld16(t_lane + (j == 0 ? BW_COL_B0 : BW_COL_B1), v);
          tmem_ld_wait();
          if (a.g_b[j]) atomicAdd(a.g_b[j] + row, __uint_as_float(v[2]));
        }
      }
      if (a.from_raw && ch == 1) {
        uint32_t v[16];
        tmem_ld16(t_lane + BW_COL_O, v);
        tmem_ld_wait();
        atomicAdd(a.g_wout + row, __uint_as_float(v[0]) + __uint_as_float(v[1]));
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (a.g_lat)
    for (int i = threadIdx.x; i < n_lat_acc; i += blockDim.x)
      if (s_lat[i] != 0.f) atomicAdd(a.g_lat + i, s_lat[i]);
  if (warp == 8) tmem_dealloc(tmem, 512);
}

static size_t bwd_smem_bytes(const nerfca_field_t& f, const TcDims& d, int l_hi, int n_layers) {
  size_t w = 0;
  for (int j = 0; j < n_layers; ++j) {
    const int l = l_hi - j;
    if (l > 0) w += 32768;
    else if (f.n_latent > 0) w += (size_t)d.kpad0 * 256;
  }
  const int l_lo = l_hi - n_layers + 1;
  const int n_lat_acc = (f.n_latent > 0 && l_lo == 0) ? f.n_phases * f.n_latent : 0;
  return w + 4 * 32768 + 4096 + 128 * sizeof(float) + (size_t)((n_lat_acc + 3) & ~3) * sizeof(float) + 8 * sizeof(uint64_t);
}

int tc_field_backward(const nerfca_field_t& f, const nerfca_samples_t& s, const float* d_raw, const void* stash, void* workspace,
                      const nerfca_field_grads_t& gr, cudaStream_t st) {
  const TcDims d = tc_dims(f);
  NERFCA_REQUIRE(f.n_latent == 0 || (size_t)f.n_phases * f.n_latent * sizeof(float) <= 16 * 1024, NERFCA_E_UNSUPPORTED,
                 "tcgen05 backward: latent table too large for the shared-memory accumulator (use precision fp32)");
  NERFCA_REQUIRE(f.n_latent == 0 || (enc_dim_of(f) % 8 + f.n_latent + 15) / 16 * 16 + enc_dim_of(f) / 8 * 8 <= d.kpad0,
                 NERFCA_E_UNSUPPORTED, "tcgen05 backward: latent columns do not fit the padded first layer");
  int rc = pack_params(f, d, workspace, st);   // same packing as the forward (weights may have changed since)
  if (rc) return rc;
  BwdArgs a;
  a.src = make_src(s);
  a.params = (const uint8_t*)workspace;
  a.stash = (const uint8_t*)stash;
  a.handoff = (uint8_t*)workspace + ((tc_param_bytes(d) + 255) & ~(size_t)255);
  a.d_raw = d_raw;
  a.g_wout = gr.weight[d.n_relu]; a.g_bout = gr.bias[d.n_relu]; a.g_lat = gr.latents;
  a.n_tiles = (s.n_points + TILE_M - 1) / TILE_M;
  a.kpad0 = d.kpad0; a.n_relu = d.n_relu; a.in_dim = d.in_dim; a.enc_dim = enc_dim_of(f); a.n_latent = f.n_latent;
  a.n_phases = f.n_phases;
  a.w_bytes = (uint32_t)d.w_bytes; a.tile_stash_bytes = (uint32_t)d.tile_stash_bytes;
  const long long grid = a.n_tiles < sm_count() ? a.n_tiles : sm_count();
  size_t max_smem = 0;
  for (int l_hi = d.n_relu - 1; l_hi >= 0; l_hi -= 2) {
    const size_t sm = bwd_smem_bytes(f, d, l_hi, l_hi >= 1 ? 2 : 1);
    max_smem = sm > max_smem ? sm : max_smem;
  }
  NERFCA_REQUIRE(max_smem <= 227 * 1024, NERFCA_E_UNSUPPORTED, "field does not fit the backward kernel's shared memory");
  NERFCA_CUDA_OK(cudaFuncSetAttribute(tc_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
  for (int l_hi = d.n_relu - 1; l_hi >= 0; l_hi -= 2) {
    a.l_hi = l_hi;
    a.n_layers = l_hi >= 1 ? 2 : 1;
    a.from_raw = (l_hi == d.n_relu - 1);
    for (int j = 0; j < 2; ++j) {
      const int l = l_hi - j;
      a.g_w[j] = (j < a.n_layers) ? gr.weight[l] : nullptr;
      a.g_b[j] = (j < a.n_layers) ? gr.bias[l] : nullptr;
    }
    tc_backward_kernel<<<(unsigned)grid, TC_THREADS, bwd_smem_bytes(f, d, l_hi, a.n_layers), st>>>(a);
    NERFCA_LAUNCH_OK();
  }
  return NERFCA_OK;
}

}  // namespace nerfca
