// fp32 path of the two coordinate MLPs (A6 / A7 forward, A11 backward) on the CUDA cores.
//
// This is the precision-reference path (NERFCA_PREC_FP32): true fp32 multiply-accumulate like the reference's
// cuBLAS/MKL sgemm, so it matches the oracle to summation-order noise.  The throughput path is the tcgen05
// kernel in mlp_tc.cu.  Structure: one tiled SGEMM template (128x128x8 block tile, 8x8 register tile) with
// three operand-layout / epilogue combinations -- forward (bias + ReLU), dgrad (ReLU mask) and split-K wgrad
// (atomic accumulate) -- plus small kernels for the 1-wide output layer, bias / latent gradients.
#include "common.cuh"

namespace nerfca {

int launch_encode(const nerfca_field_t& field, const nerfca_samples_t& samples, long long p0, long long np, float* out,
                  cudaStream_t st);

constexpr int BM = 128, BN = 128, BK = 8;
enum { EPI_BIAS_RELU = 0, EPI_MASK = 1, EPI_ATOMIC = 2 };

struct GemmArgs {
  const float* A; const float* B; float* C;
  const float* bias;   // [N] or null (EPI_BIAS_RELU)
  const float* mask;   // same layout as C (EPI_MASK): pass gradient where mask > 0
  long long M, K;      // C is [M, N];  reduction length K
  int N;
  long long lda, ldb, ldc;
  long long k_split;   // reduction elements per blockIdx.z
};

// C[m,n] (+)= sum_k A(m,k) B(k,n)
//   A(m,k) = A_KC ? A[m*lda + k] : A[k*lda + m]      B(k,n) = B_KC ? B[n*ldb + k] : B[k*ldb + n]
template <bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const long long kb = (long long)blockIdx.z * g.k_split;
  const long long ke = (kb + g.k_split < g.K) ? kb + g.k_split : g.K;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (long long k0 = kb; k0 < ke; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int mm, kk;
      if (A_KC) { const int e = t * 4 + i; mm = e >> 3; kk = e & 7; }
      else      { const int e = t + i * 256; mm = e & 127; kk = e >> 7; }
      const long long m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < g.M && k < ke) v = A_KC ? __ldg(g.A + m * g.lda + k) : __ldg(g.A + k * g.lda + m);
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int nn, kk;
      if (B_KC) { const int e = t * 4 + i; nn = e >> 3; kk = e & 7; }
      else      { const int e = t + i * 256; nn = e & 127; kk = e >> 7; }
      const int n = n0 + nn;
      const long long k = k0 + kk;
      float v = 0.f;
      if (n < g.N && k < ke) v = B_KC ? __ldg(g.B + (long long)n * g.ldb + k) : __ldg(g.B + k * g.ldb + n);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      float v = acc[i][j];
      float* c = g.C + m * g.ldc + n;
      if (EPI == EPI_BIAS_RELU) {
        if (g.bias) v += __ldg(g.bias + n);
        *c = fmaxf(v, 0.f);
      } else if (EPI == EPI_MASK) {
        *c = (__ldg(g.mask + m * g.ldc + n) > 0.f) ? v : 0.f;
      } else {
        atomicAdd(c, v);
      }
    }
  }
}

template <bool A_KC, bool B_KC, int EPI>
static int run_gemm(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return NERFCA_OK;
  dim3 grid(div_up(g.M, BM), div_up(g.N, BN), div_up(g.K, g.k_split));
  sgemm_kernel<A_KC, B_KC, EPI><<<grid, 256, 0, st>>>(g);
  NERFCA_LAUNCH_OK();
  return NERFCA_OK;
}

// output layer (hidden -> 1): one warp per sample
__global__ void out_forward_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ b,
                                   long long np, int H, float* __restrict__ raw) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= np) return;
  float acc = 0.f;
  for (int k = lane; k < H; k += 32) acc = fmaf(__ldg(h + row * H + k), __ldg(w + k), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) raw[row] = acc + (b ? __ldg(b) : 0.f);
}

// backward of the output layer: dZ[p,k] = d_raw[p] w[k] 1[h>0];  dW[k] += sum_p d_raw[p] h[p,k];  db += sum_p d_raw[p]
constexpr int ROWS_PER_BLOCK = 128;
__global__ void out_backward_kernel(const float* __restrict__ d_raw, const float* __restrict__ h, const float* __restrict__ w,
                                    long long np, int H, float* __restrict__ dz, float* __restrict__ dw, float* __restrict__ db) {
  const long long r0 = (long long)blockIdx.x * ROWS_PER_BLOCK;
  const long long r1 = (r0 + ROWS_PER_BLOCK < np) ? r0 + ROWS_PER_BLOCK : np;
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    const float wk = __ldg(w + k);
    float aw = 0.f, ab = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const float g = __ldg(d_raw + r), hv = __ldg(h + r * H + k);
      dz[r * H + k] = hv > 0.f ? g * wk : 0.f;
      aw = fmaf(g, hv, aw);
      ab += g;
    }
    atomicAdd(dw + k, aw);
    if (k == 0 && db) atomicAdd(db, ab);
  }
}

// db[n] += sum_p dz[p,n]
__global__ void colsum_kernel(const float* __restrict__ dz, long long np, int H, float* __restrict__ db) {
  const long long r0 = (long long)blockIdx.x * ROWS_PER_BLOCK;
  const long long r1 = (r0 + ROWS_PER_BLOCK < np) ? r0 + ROWS_PER_BLOCK : np;
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    float a = 0.f;
    for (long long r = r0; r < r1; ++r) a += __ldg(dz + r * H + k);
    atomicAdd(db + k, a);
  }
}

// d time_latents[phase[p], t] += sum_n dz0[p,n] W0[n, enc_dim + t]   (scatter-add of Temporal.py:144-147's gather)
__global__ void latent_grad_kernel(const float* __restrict__ dz0, const float* __restrict__ w0, SampleSrc src, int H, int D,
                                   int enc_dim, int T, int n_phases, int use_smem, float* __restrict__ dlat) {
  extern __shared__ float sacc[];  // [n_phases * T] when use_smem
  if (use_smem) {
    for (int i = threadIdx.x; i < n_phases * T; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
  }
  const int per_block = blockDim.x / T;
  const int local = threadIdx.x / T, tt = threadIdx.x - local * T;
  const long long p = (long long)blockIdx.x * per_block + local;
  if (local < per_block && p < src.n_points) {
    float a = 0.f;
    for (int n = 0; n < H; ++n) a = fmaf(__ldg(dz0 + p * H + n), __ldg(w0 + (size_t)n * D + enc_dim + tt), a);
    const int ph = load_phase(src, p);
    if (ph >= 0 && ph < n_phases) atomicAdd(use_smem ? &sacc[ph * T + tt] : dlat + (size_t)ph * T + tt, a);
  }
  if (!use_smem) return;
  __syncthreads();
  for (int i = threadIdx.x; i < n_phases * T; i += blockDim.x)
    if (sacc[i] != 0.f) atomicAdd(dlat + i, sacc[i]);
}

constexpr long long SIMT_CHUNK = 262144;  // samples per pass over the layer stack

size_t simt_stash_bytes(const nerfca_field_t& f, long long P) {
  return (size_t)P * ((size_t)in_dim_of(f) + (size_t)(f.n_hidden + 1) * f.hidden) * sizeof(float);
}
size_t simt_workspace_bytes(const nerfca_field_t& f, long long P, int backward) {
  const long long ch = P < SIMT_CHUNK ? P : SIMT_CHUNK;
  if (backward) return (size_t)ch * f.hidden * 2 * sizeof(float);
  return (size_t)ch * ((size_t)in_dim_of(f) + 2 * (size_t)f.hidden) * sizeof(float);
}

int simt_field_forward(const nerfca_field_t& f, const nerfca_samples_t& s, float* raw_out, void* stash, void* workspace,
                       cudaStream_t st) {
  const long long P = s.n_points;
  const int D = in_dim_of(f), H = f.hidden, L = f.n_hidden + 1;  // L layers with ReLU
  float* st_enc = (float*)stash;
  float* st_h = stash ? st_enc + (size_t)P * D : nullptr;
  for (long long c0 = 0; c0 < P; c0 += SIMT_CHUNK) {
    const long long np = (P - c0 < SIMT_CHUNK) ? P - c0 : SIMT_CHUNK;
    const long long ch = P < SIMT_CHUNK ? P : SIMT_CHUNK;
    float* enc = stash ? st_enc + (size_t)c0 * D : (float*)workspace;
    float* ping[2] = {(float*)workspace + (size_t)ch * D, (float*)workspace + (size_t)ch * D + (size_t)ch * H};
    int rc = launch_encode(f, s, c0, np, enc, st);
    if (rc) return rc;
    const float* in = enc;
    long long ld_in = D;
    int K = D;
    for (int l = 0; l < L; ++l) {
      float* out = stash ? st_h + ((size_t)l * P + c0) * H : ping[l & 1];
      GemmArgs g{};
      g.A = in; g.lda = ld_in; g.B = f.weight[l]; g.ldb = K; g.C = out; g.ldc = H; g.bias = f.bias[l];
      g.M = np; g.N = H; g.K = K; g.k_split = K;
      rc = run_gemm<true, true, EPI_BIAS_RELU>(g, st);
      if (rc) return rc;
      in = out; ld_in = H; K = H;
    }
    out_forward_kernel<<<div_up(np * 32, 256), 256, 0, st>>>(in, f.weight[L], f.bias[L], np, H, raw_out + c0);
    NERFCA_LAUNCH_OK();
  }
  return NERFCA_OK;
}

int simt_field_backward(const nerfca_field_t& f, const nerfca_samples_t& s, const float* d_raw, const void* stash,
                        void* workspace, const nerfca_field_grads_t& gr, cudaStream_t st) {
  const long long P = s.n_points;
  const int D = in_dim_of(f), H = f.hidden, L = f.n_hidden + 1;
  const float* st_enc = (const float*)stash;
  const float* st_h = st_enc + (size_t)P * D;
  const long long ch = P < SIMT_CHUNK ? P : SIMT_CHUNK;
  float* X = (float*)workspace;
  float* Y = X + (size_t)ch * H;
  for (long long c0 = 0; c0 < P; c0 += SIMT_CHUNK) {
    const long long np = (P - c0 < SIMT_CHUNK) ? P - c0 : SIMT_CHUNK;
    const unsigned rb = div_up(np, ROWS_PER_BLOCK);
    const float* h_last = st_h + ((size_t)(L - 1) * P + c0) * H;
    out_backward_kernel<<<rb, 128, 0, st>>>(d_raw + c0, h_last, f.weight[L], np, H, X, gr.weight[L], gr.bias[L]);
    NERFCA_LAUNCH_OK();
    for (int l = L - 1; l >= 0; --l) {
      const float* h_prev = (l > 0) ? st_h + ((size_t)(l - 1) * P + c0) * H : st_enc + (size_t)c0 * D;
      const int Kin = (l > 0) ? H : D;
      {  // wgrad: dW_l[n, k] += sum_p X[p, n] h_prev[p, k]
        GemmArgs g{};
        g.A = X; g.lda = H; g.B = h_prev; g.ldb = Kin; g.C = gr.weight[l]; g.ldc = Kin;
        g.M = H; g.N = Kin; g.K = np; g.k_split = 2048;
        int rc = run_gemm<false, false, EPI_ATOMIC>(g, st);
        if (rc) return rc;
      }
      if (gr.bias[l]) {
        colsum_kernel<<<rb, 128, 0, st>>>(X, np, H, gr.bias[l]);
        NERFCA_LAUNCH_OK();
      }
      if (l > 0) {  // dgrad: Y[p, k] = (sum_n X[p, n] W_l[n, k]) * 1[h_prev[p,k] > 0]
        GemmArgs g{};
        g.A = X; g.lda = H; g.B = f.weight[l]; g.ldb = H; g.C = Y; g.ldc = H; g.mask = h_prev;
        g.M = np; g.N = H; g.K = H; g.k_split = H;
        int rc = run_gemm<true, false, EPI_MASK>(g, st);
        if (rc) return rc;
        float* tmp = X; X = Y; Y = tmp;
      } else if (f.n_latent > 0 && gr.latents) {
        const int T = f.n_latent;
        const int threads = (256 / T) * T;
        const int per_block = threads / T;
        const int use_smem = (size_t)f.n_phases * T * sizeof(float) <= 32 * 1024;  // huge tables (query_time) go direct
        const size_t smem = use_smem ? (size_t)f.n_phases * T * sizeof(float) : 0;
        latent_grad_kernel<<<div_up(np, per_block), threads, smem, st>>>(X, f.weight[0], make_src(s, c0, np), H, D,
                                                                       enc_dim_of(f), T, f.n_phases, use_smem, gr.latents);
        NERFCA_LAUNCH_OK();
      }
    }
  }
  return NERFCA_OK;
}

}  // namespace nerfca
