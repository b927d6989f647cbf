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

size_t tc_workspace_bytes(const nerfca_field_t& f, long long P, int backward) {
  const TcDims d = tc_dims(f);
  (void)P;
  if (!backward) return tc_param_bytes(d);
  return tc_param_bytes(d);
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

int tc_field_backward(const nerfca_field_t&, const nerfca_samples_t&, const float*, const void*, void*,
                      const nerfca_field_grads_t&, cudaStream_t) {
  set_error("tcgen05 backward not built yet");
  return NERFCA_E_UNSUPPORTED;
}

}  // namespace nerfca
