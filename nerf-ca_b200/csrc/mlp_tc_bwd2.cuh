// Backward of the two coordinate MLPs, second generation: TWO tiles in flight per CTA.  (Included by mlp_tc.cu.)
//
// What changed against tc_bwd_kernel (which stays available, NERFCA_BWD_V1=1): there a CTA worked on ONE tile at a time, so
// the tensor pipe idled during every epilogue (tcgen05.ld -> ReLU mask -> pack -> st.shared, ~1 100 cycles per chain step) and
// the epilogue warps idled during every MMA: ~40 % tensor-pipe activity.  Two tiles did not fit: tensor memory was full with
// the two resident weight-gradient accumulators + one chain accumulator + its TMEM A operand, shared memory with four 32 KB
// buffers per tile.  Here
//   * the chain accumulator (128 TMEM columns) is SHARED by the two tiles in flight: it is only occupied from the start of a
//     dgrad GEMM until the epilogue's tcgen05.ld has drained it (~900 of the ~2 200 cycles between two chain steps of a tile),
//     so the two tiles alternate on it through a ticket lock in shared memory (issuers take a ticket when their operand is
//     ready, epilogue warps release after tcgen05.wait::ld);
//   * every dgrad takes its A operand from shared memory (SS form), no TMEM A region;
//   * a tile needs only TWO 32 KB buffers: gradients overwrite, in place, the activation tile whose ReLU pattern produced
//     them once the weight-gradient GEMM that read the activations has completed (top: H3 -> dZ3, R -> H2; bottom:
//     H1 -> dZ1, dZ2 -> H0 -> dZ0), the second activation tile of a step is loaded into the buffer its predecessor vacates;
//   * the output weight is folded into W4 once per launch (W4' = diag(w_out) W4 in shared memory), so R = d_raw 1[H4 > 0]
//     feeds both dgrad 4 and wgrad 4;
//   * each of the two slots has its own epilogue warps (8), MMA-issuing warp and load warp with plain in-order waits; all
//     hand-offs are mbarriers (one arrival per warp), the weight-gradient accumulators are zero-filled once and every
//     GEMM accumulates.
// Roles, tile order, the dZ2 hand-off ring through L2 and its flags are those of tc_bwd_kernel.
//
// TMEM   top:    ACC [0,128) | WG4 [128,272) | WG3 [272,416)                 (N = 144: column 128 = bias gradient)
//        bottom: ACC [0,128) | WG2 [128,272) | WG1 [272,416) | WG0 [416,512)        (WG2 / WG1: N = 144 as in the top role)
// smem   top:    W3 | W4' | slot s: b1[s] (H3 -> dZ3, + constant-1 block), b2[s] (R -> H2, + constant-1 block) | H4 patterns | d_raw | ...
//        bottom: W1 | W2  | slot s: c1[s] (dZ2 -> H0 -> dZ0, + constant-1 block), c2[s] (H1 -> dZ1 -> X0, + constant-1 block) | W0 latent chunks | ...
//   * bottom role: the rebuilt first-layer input X0 of a tile goes into the tile's OWN c2 buffer once weight gradient 1 has read dZ1
//     from it (there used to be one X0 buffer per CTA, refilled only after the previous tile's weight gradient 0: measured, it paced
//     the whole role, ISS 21->22 = 4 000-5 000 cycles of waiting per tile); the 24 KB this frees hold the constant-1 blocks behind the
//     buffers, so the bias gradients ride in N = 144 weight-gradient GEMMs instead of separate N = 16 GEMMs that re-read dZ.
#pragma once

constexpr int BWD2_THREADS = 24 * 32;   // warps 0-15: epilogue (slot = warp / 8); 16, 17: issuers; 18, 19: loaders; 20-23: X0 producers / publisher
constexpr uint32_t T2_ACC = 0, T2_WG4 = 128, T2_WG3 = 272;
constexpr uint32_t B2_ACC = 0, B2_WG2 = 128, B2_WG1 = 272, B2_WG0 = 416;     // WG2 / WG1 are 144 wide: column 128 = bias gradient

// ---- the shared accumulator's ticket lock ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t atom_add_shared(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
// one issuing lane: take the next ticket, wait until every earlier user's epilogue warps (8 per use) have drained the accumulator
__device__ __forceinline__ void acc_acquire(uint32_t ticket_addr, uint32_t rel_addr) {
  const uint32_t need = 8u * atom_add_shared(ticket_addr, 1u);
  if (ld_acquire_cta_shared(rel_addr) < need) {
    const long long t0 = clock64();
    while (ld_acquire_cta_shared(rel_addr) < need) {
      __nanosleep(40);
      if (clock64() - t0 > 4000000000LL) {
#ifdef NERFCA_TIMELINE_BUILD
        printf("accumulator lock timeout: block %d warp %d need %u have %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), need, ld_acquire_cta_shared(rel_addr));
#endif
        __trap();
      }
    }
  }
  tc_fence_after();
}
// whole epilogue warp, right behind tcgen05.wait::ld of its accumulator columns
__device__ __forceinline__ void acc_release(uint32_t rel_addr, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) red_release_cta_shared_add(rel_addr, 1u);
}
// whole warp: this warp's shared-memory writes are visible to the async proxy, one arrival
__device__ __forceinline__ void warp_publish_smem(uint32_t bar, int lane) {
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// zero `ncols` (multiple of 8) accumulator columns of this warp's 32 TMEM lanes
__device__ __forceinline__ void tmem_zero(uint32_t taddr, int ncols) {
  int c = 0;
  for (; c + 32 <= ncols; c += 32) {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    tmem_st32(taddr + c, z);
  }
  for (; c < ncols; c += 8) tmem_st8(taddr + c, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u);
  tmem_st_wait();
}

// =====================================================================================================================
// top role: output layer, layers 4 and 3
// =====================================================================================================================
constexpr size_t T2_OFF_W3 = 0, T2_OFF_W4 = TILE_BYTES, T2_OFF_BUF = 2 * (size_t)TILE_BYTES;
constexpr size_t T2_OFF_M4 = T2_OFF_BUF + 4 * (size_t)TOP_BUF_STRIDE;       // 2 x 2 KB
constexpr size_t T2_OFF_G = T2_OFF_M4 + 2 * MASK_BYTES;                     // 2 x 128 f32
constexpr size_t T2_OFF_WO = T2_OFF_G + 2 * 128 * 4;                        // 128 f32
constexpr size_t T2_OFF_MISC = T2_OFF_WO + 128 * 4;                         // gbout f32, acc_rel, acc_ticket, pub_cnt[2], slot_ok[2], pad
constexpr size_t T2_OFF_BAR = T2_OFF_MISC + 32;
constexpr int T2_N_BAR = 2 + 2 * 9;
constexpr size_t TOP2_SMEM = T2_OFF_BAR + T2_N_BAR * 8 + 16;

__device__ __forceinline__ void bwd2_top_role(const BwdArgs& a, const BwdNet& nt, const long long worker, const long long n_workers) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* s_w3 = smem + T2_OFF_W3;
  uint8_t* s_w4 = smem + T2_OFF_W4;
  uint8_t* s_buf = smem + T2_OFF_BUF;
  uint8_t* s_m4 = smem + T2_OFF_M4;
  float* s_g = reinterpret_cast<float*>(smem + T2_OFF_G);
  float* s_wo = reinterpret_cast<float*>(smem + T2_OFF_WO);
  float* s_gbout = reinterpret_cast<float*>(smem + T2_OFF_MISC);
  const uint32_t acc_rel = smem_u32(smem + T2_OFF_MISC + 4), acc_ticket = acc_rel + 4, pub_cnt0 = acc_rel + 8, slot_ok0 = acc_rel + 16;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + T2_OFF_BAR);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + T2_N_BAR);
  const uint32_t bar_w = smem_u32(s_bar), bar_setup = bar_w + 8;
  // per slot s (stride 72 B): ready, acc, wg4, wg3, ldm, ldh3, ldh2, mfree, slot
  auto bar_of = [&](int s, int k) { return bar_w + 16u + (uint32_t)(s * 9 + k) * 8u; };
  enum { B_READY = 0, B_ACC, B_WG4, B_WG3, B_LDM, B_LDH3, B_LDH2, B_MFREE, B_SLOT };

  if (warp == 16) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_setup, 16);
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar_of(s, B_READY), 8); mbar_init(bar_of(s, B_ACC), 1); mbar_init(bar_of(s, B_WG4), 1); mbar_init(bar_of(s, B_WG3), 1);
        mbar_init(bar_of(s, B_LDM), 1); mbar_init(bar_of(s, B_LDH3), 1); mbar_init(bar_of(s, B_LDH2), 1);
        mbar_init(bar_of(s, B_MFREE), 8); mbar_init(bar_of(s, B_SLOT), 1);      // (B_SLOT: unused, the ring-slot signal is a counter)
      }
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), 512);
  }
  {
    const float* fb = reinterpret_cast<const float*>(nt.pack + nt.f32_off);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_wo[i] = __ldg(fb + 5 * 128 + i);
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x)      // the constant-1 block behind each tile buffer
      reinterpret_cast<uint4*>(s_buf + (size_t)(i >> 8) * TOP_BUF_STRIDE + TILE_BYTES)[i & 255] =
          ((i & 255) < 128) ? make_uint4(0x00003F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x < 8) reinterpret_cast<uint32_t*>(s_gbout)[threadIdx.x] = 0u;     // gbout, lock words, publication counters
  }
  tc_fence_before();
  fence_proxy_async();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const long long n_my = (a.n_tiles > worker) ? (a.n_tiles - worker + n_workers - 1) / n_workers : 0;
  [[maybe_unused]] int tl_n = 0;     // developer timeline (make TL=1, NERFCA_TIMELINE=top): regions 1/2 = epilogue slot 0/1 (warp 1 / 9, lane 0), 3 = issuer 0, 0 = issuer 1

  // register budget (setmaxnreg; the CTA's pool starts EMPTY: only what warps hand back can be taken): the kernel launches with
  // 80 x 768 = 61 440 registers = 16 epilogue warps x 88 + 8 other warps x 64.  With these budgets ptxas spills 24 bytes in the whole
  // kernel.  Spills are poison here: the flag polls of the hand-off (gpu-scope acquire = CCTL.IVALL) keep invalidating L1, so
  // every reload of a spilled MMA descriptor went to L2 (measured: 3 600 cycles to issue the 8 MMAs of one dgrad at 40 registers).
  if (warp >= 16) reg_dealloc<64>();
  if (warp == 18 || warp == 19) {
    // ================= load warp of slot s =================
    const int s = warp - 18;
    const long long n_s = (n_my + 1 - s) / 2;
    const uint32_t b1 = smem_u32(s_buf) + (uint32_t)(2 * s) * TOP_BUF_STRIDE, b2 = b1 + TOP_BUF_STRIDE;
    const uint32_t m4 = smem_u32(s_m4) + (uint32_t)s * MASK_BYTES;
    float* g_dst = s_g + s * 128;
    if (s == 0 && lane == 0) {
      mbar_expect_tx(bar_w, 2 * TILE_BYTES);
      bulk_g2s(smem_u32(s_w3), nt.pack + nt.w0_bytes + 2 * (size_t)TILE_BYTES, TILE_BYTES, bar_w);
      bulk_g2s(smem_u32(s_w4), nt.pack + nt.w0_bytes + 3 * (size_t)TILE_BYTES, TILE_BYTES, bar_w);
    }
    // H4 pattern + d_raw of the slot's j-th tile -> the small per-slot buffers (whole warp: d_raw by plain loads, the last tile may
    // be ragged and rows past the end get a zero gradient); free once step A of the slot's previous tile has consumed its own
    auto load_pattern = [&](long long j) {
      const long long tile = worker + (s + 2 * j) * n_workers;
      float gv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long p = tile * TILE_M + lane * 4 + e;
        gv[e] = (p < a.src.n_points) ? __ldg(nt.d_raw + p) : 0.f;
      }
      if (j > 0) mbar_wait(bar_of(s, B_MFREE), (uint32_t)((j - 1) & 1));
      *reinterpret_cast<float4*>(g_dst + lane * 4) = make_float4(gv[0], gv[1], gv[2], gv[3]);
      __syncwarp();
      if (lane == 0) {
        mbar_expect_tx(bar_of(s, B_LDM), MASK_BYTES);
        bulk_g2s(m4, nt.stash + (size_t)tile * STASH_STRIDE + STASH_PATTERN4_OFF, MASK_BYTES, bar_of(s, B_LDM));
      }
      __syncwarp();
    };
    if (n_s > 0) load_pattern(0);
    for (long long j = 0; j < n_s; ++j) {
      const long long tile = worker + (s + 2 * j) * n_workers;
      const uint8_t* st = nt.stash + (size_t)tile * STASH_STRIDE;
      const uint32_t pj = (uint32_t)(j & 1);
      if (lane == 0) {
        if (j > 0) mbar_wait(bar_of(s, B_WG3), pj ^ 1);          // weight gradient 3 of the previous tile no longer reads dZ3 (b1) / H2 (b2)
        mbar_expect_tx(bar_of(s, B_LDH3), TILE_BYTES);
        bulk_g2s(b1, st + 3 * (size_t)TILE_BYTES, TILE_BYTES, bar_of(s, B_LDH3));
      }
      __syncwarp();
      if (j + 1 < n_s) load_pattern(j + 1);                       // (waits for step A of tile j: early in the tile)
      if (lane == 0) {
        mbar_wait(bar_of(s, B_WG4), pj);                           // weight gradient 4 no longer reads R: its buffer takes H2
        mbar_expect_tx(bar_of(s, B_LDH2), TILE_BYTES);
        bulk_g2s(b2, st + 2 * (size_t)TILE_BYTES, TILE_BYTES, bar_of(s, B_LDH2));
        if (j + 1 < n_s) {                                         // the slot's next tile: pull its stash pieces into L2 now
          const uint8_t* nx = nt.stash + (size_t)(tile + 2 * n_workers) * STASH_STRIDE;
          bulk_prefetch_l2(nx + 2 * (size_t)TILE_BYTES, TILE_BYTES);
          bulk_prefetch_l2(nx + 3 * (size_t)TILE_BYTES, TILE_BYTES);
        }
        // hand-off slot of this tile: free once the bottom role has copied out the tile that used it `ring` tiles earlier.  Signalled
        // through a monotonic shared-memory counter (slot_ok[s] = tiles whose slot is free), not an mbarrier phase: the epilogue reads
        // it one tile late (deferred dZ2 store) and a counter cannot alias however far the loader runs ahead.
        if (tile >= a.ring) wait_flag_ge(nt.consumed + (tile - a.ring), 1u);
        red_release_cta_shared_add(slot_ok0 + 4u * s, 1u);
      }
      __syncwarp();
    }
  } else if (warp == 16 || warp == 17) {
    // ================= MMA-issuing warp of slot s =================
    const int s = warp - 16;
    const long long n_s = (n_my + 1 - s) / 2;
    const uint32_t w3 = smem_u32(s_w3), w4 = smem_u32(s_w4);
    const uint32_t b1 = smem_u32(s_buf) + (uint32_t)(2 * s) * TOP_BUF_STRIDE, b2 = b1 + TOP_BUF_STRIDE;
    constexpr uint32_t KK = KSTEP_KMAJOR, KM = KSTEP_MNMAJOR;
    constexpr uint32_t id_dgrad = instr_desc(128, 128, 0, 1), id_wgrad = instr_desc(128, 144, 1, 1);
    if (lane == 0) {
      mbar_wait(bar_setup, 0);                                     // W3 / W4' in place, accumulators zeroed
      tc_fence_after();
      for (long long j = 0; j < n_s; ++j) {
        const uint32_t pj = (uint32_t)(j & 1);
        [[maybe_unused]] const int tb = s ? 100 : 3000;
        mbar_wait(bar_of(s, B_READY), 0);                          // step A: R is in b2
        NERFCA_TL(true, tb + 1);
        acc_acquire(acc_ticket, acc_rel);
        NERFCA_TL(true, tb + 2);
        umma_k<8, KK, KM>(tmem + T2_ACC, kmajor(b2), mnmajor(w4), id_dgrad, 0);                  // dH3 = R W4'
        umma_commit(bar_of(s, B_ACC));
        NERFCA_TL(true, tb + 3);
        mbar_wait(bar_of(s, B_LDH3), pj);
        tc_fence_after();
        NERFCA_TL(true, tb + 4);
        umma_k<8, KM, KM>(tmem + T2_WG4, mnmajor(b2), mnmajor(b1), id_wgrad, 1);                 // WG4 += R^T [H3 | 1]
        umma_commit(bar_of(s, B_WG4));
        NERFCA_TL(true, tb + 5);
        mbar_wait(bar_of(s, B_READY), 1);                          // step B: dZ3 is in b1
        NERFCA_TL(true, tb + 11);
        acc_acquire(acc_ticket, acc_rel);
        NERFCA_TL(true, tb + 12);
        umma_k<8, KK, KM>(tmem + T2_ACC, kmajor(b1), mnmajor(w3), id_dgrad, 0);                  // dH2 = dZ3 W3
        umma_commit(bar_of(s, B_ACC));
        NERFCA_TL(true, tb + 13);
        mbar_wait(bar_of(s, B_LDH2), pj);
        tc_fence_after();
        NERFCA_TL(true, tb + 14);
        umma_k<8, KM, KM>(tmem + T2_WG3, mnmajor(b1), mnmajor(b2), id_wgrad, 1);                 // WG3 += dZ3^T [H2 | 1]
        umma_commit(bar_of(s, B_WG3));
        NERFCA_TL(true, tb + 15);
      }
    }
    __syncwarp();
  } else if (warp == 20) {
    // ================= publisher: turns 8 warp arrivals behind a tile's dZ2 stores into the GPU-scope `produced` flag =================
    if (lane == 0) {
      long long done[2] = {0, 0};
      const long long n_sl[2] = {(n_my + 1) / 2, n_my / 2};
      long long left = n_my;
      const long long t0 = clock64();
      long long last = t0;
      while (left > 0) {
        bool any = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (done[s] < n_sl[s] && ld_acquire_cta_shared(pub_cnt0 + 4u * s) >= 8u * (uint32_t)(done[s] + 1)) {
            red_release_gpu_add(nt.produced + (worker + (s + 2 * done[s]) * n_workers), 8u);
            ++done[s];
            --left;
            any = true;
          }
        }
        if (any) last = clock64();
        else {
          __nanosleep(200);
          if (clock64() - last > 4000000000LL) {
#ifdef NERFCA_TIMELINE_BUILD
            printf("publisher timeout: block %d done %lld %lld of %lld\n", (int)blockIdx.x, done[0], done[1], n_my);
#endif
            __trap();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 16) {
    reg_alloc<88>();
    // ================= 2 x 8 epilogue warps: slot = warp / 8, thread = (row, column half) =================
    const int slot = warp >> 3, q = warp & 3, ch = (warp >> 2) & 1;
    const int row = q * 32 + lane;
    const long long n_s = (n_my + 1 - slot) / 2;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t k_acc = t_lane + T2_ACC + ch * 64;
    uint32_t k_rowoff = (uint32_t)(ch * 8) * CHUNK_BYTES + (uint32_t)row * 16u;      // this thread's first chunk inside a tile
    uint32_t b1 = smem_u32(s_buf) + (uint32_t)(2 * slot) * TOP_BUF_STRIDE + k_rowoff, b2 = b1 + TOP_BUF_STRIDE;
    uint32_t k_m4 = smem_u32(s_m4) + (uint32_t)slot * MASK_BYTES + (uint32_t)row * 16u + (uint32_t)ch * 8u;
    pin(k_acc); pin(k_rowoff); pin(b1); pin(b2); pin(k_m4);
    const float* g_src = s_g + slot * 128;
    const uint32_t pub_cnt = pub_cnt0 + 4u * slot;
    // ---- set-up: zero the weight-gradient accumulators (288 columns over the 4 (slot, ch) warp groups), fold w_out into W4
    tmem_zero(t_lane + T2_WG4 + (uint32_t)(slot * 2 + ch) * 72u, 72);
    mbar_wait(bar_w, 0);
    for (int i = threadIdx.x; i < 2048; i += 512) {       // uint4 (chunk c, row n) of the W4 tile: scale row n by w_out[n]
      const float sc = s_wo[i & 127];
      uint4 v = reinterpret_cast<uint4*>(s_w4)[i];
      auto scale2 = [&](uint32_t x) {
        const float lo = __uint_as_float(x << 16), hi = __uint_as_float(x & 0xFFFF0000u);
        return pack_bf16x2(lo * sc, hi * sc);
      };
      v.x = scale2(v.x); v.y = scale2(v.y); v.z = scale2(v.z); v.w = scale2(v.w);
      reinterpret_cast<uint4*>(s_w4)[i] = v;
    }
    tc_fence_before();
    warp_publish_smem(bar_setup, lane);
    uint32_t ring_slot = (uint32_t)((worker + slot * n_workers) % a.ring);
    const uint32_t ring_step = (uint32_t)((2 * n_workers) % a.ring);
    uint32_t ph_acc = 0;
    float gb_sum = 0.f;
    uint32_t dz[32];
    auto store_dz = [&](long long jt) {               // dZ2 of the slot's jt-th tile -> its hand-off slot, then one arrival for the publisher
      if (ld_acquire_cta_shared(slot_ok0 + 4u * slot) < (uint32_t)(jt + 1)) {
        const long long t0 = clock64();
        while (ld_acquire_cta_shared(slot_ok0 + 4u * slot) < (uint32_t)(jt + 1)) {
          __nanosleep(64);
          if (clock64() - t0 > 4000000000LL) __trap();
        }
      }
      uint8_t* dst = nt.handoff + (size_t)ring_slot * TILE_BYTES + k_rowoff;
      ring_slot += ring_step;
      if (ring_slot >= (uint32_t)a.ring) ring_slot -= (uint32_t)a.ring;
#pragma unroll
      for (int c = 0; c < 8; ++c)      // streaming stores straight to L2
        __stcs(reinterpret_cast<uint4*>(dst + c * CHUNK_BYTES), make_uint4(dz[4 * c], dz[4 * c + 1], dz[4 * c + 2], dz[4 * c + 3]));
      __syncwarp();
      if (lane == 0) red_release_cta_shared_add(pub_cnt, 1u);
    };
    for (long long j = 0; j < n_s; ++j) {
      const uint32_t pj = (uint32_t)(j & 1);
      uint32_t va[32], vb[32], w[32];
      [[maybe_unused]] const bool tl_me = (warp & 7) == 1 && lane == 0;
      [[maybe_unused]] const int te = 1000 + 1000 * slot;
      // ---- step A: R = d_raw 1[H4 > 0] -> b2 (A operand of dgrad 4 and of wgrad 4)
      NERFCA_TL(tl_me, te + 0);
      mbar_wait(bar_of(slot, B_LDM), pj);
      NERFCA_TL(tl_me, te + 1);
      if (j > 0) mbar_wait(bar_of(slot, B_WG3), pj ^ 1);         // weight gradient 3 of the previous tile no longer reads H2 in b2
      NERFCA_TL(tl_me, te + 2);
      {
        const float g_cur = g_src[row];
        if (ch == 0) gb_sum += g_cur;
        // pattern word g covers this thread's columns [32 g, 32 g + 32): bit i / 16 + i = columns 2i / 2i + 1.  (0 or 1 in each half) times
        // the 16 bf16 bits of d_raw gives the packed pair without a carry between the halves.
        const uint2 mb = lds_u2(k_m4);
        const uint32_t gbits = pack_bf16x2(g_cur, g_cur) & 0xFFFFu;
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2) {
          w[i2] = ((mb.x >> i2) & 0x00010001u) * gbits;
          w[16 + i2] = ((mb.y >> i2) & 0x00010001u) * gbits;
        }
        sts_row64(b2, w);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_of(slot, B_MFREE)); mbar_arrive(bar_of(slot, B_READY)); }
      NERFCA_TL(tl_me, te + 3);
      if (j > 0) store_dz(j - 1);                      // the previous tile's dZ2 drains while dgrad 4 runs
      NERFCA_TL(tl_me, te + 4);
      // ---- step B: dZ3 = dH3 * 1[H3 > 0] -> b1, over H3 itself once weight gradient 4 has read it
      mbar_wait(bar_of(slot, B_ACC), ph_acc); ph_acc ^= 1;
      tc_fence_after();
      NERFCA_TL(tl_me, te + 10);
      ld_acc64(k_acc, va, vb);
      acc_release(acc_rel, lane);
      NERFCA_TL(tl_me, te + 11);
      mbar_wait(bar_of(slot, B_LDH3), pj);
      NERFCA_TL(tl_me, te + 12);
      masked_grad_pack64(va, vb, b1, w);
      NERFCA_TL(tl_me, te + 13);
      mbar_wait(bar_of(slot, B_WG4), pj);
      NERFCA_TL(tl_me, te + 14);
      sts_row64(b1, w);
      warp_publish_smem(bar_of(slot, B_READY), lane);
      NERFCA_TL(tl_me, te + 15);
      // ---- step C: dZ2 = dH2 * 1[H2 > 0] -> the hand-off ring (L2)
      mbar_wait(bar_of(slot, B_ACC), ph_acc); ph_acc ^= 1;
      tc_fence_after();
      NERFCA_TL(tl_me, te + 20);
      ld_acc64(k_acc, va, vb);
      acc_release(acc_rel, lane);
      NERFCA_TL(tl_me, te + 21);
      mbar_wait(bar_of(slot, B_LDH2), pj);
      NERFCA_TL(tl_me, te + 22);
      // (dz leaves behind step A of the slot's next tile: its 32 KB then drain through the store path while the CTA waits for dgrad 4
      // anyway; issued here they kept the warps in the store queue for ~2 000 cycles before step A could start)
      masked_grad_pack64(va, vb, b2, dz);
      NERFCA_TL(tl_me, te + 23);
    }
    if (n_s > 0) store_dz(n_s - 1);                    // the last tile's dZ2
    // ---- every MMA of both slots has completed: flush the TMEM-resident accumulators
    if (n_s > 0) mbar_wait(bar_of(slot, B_WG3), (uint32_t)((n_s - 1) & 1));
    tc_fence_before();
    named_bar_sync(1, 512);
    tc_fence_after();
    if (ch == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) gb_sum += __shfl_xor_sync(0xffffffffu, gb_sum, o);
      if (lane == 0 && nt.g_b[5]) atomicAdd(s_gbout, gb_sum);
    }
    if (n_my > 0) {
      const float* fb = reinterpret_cast<const float*>(nt.pack + nt.f32_off);
      const float wo_row = s_wo[row];
      if (slot == 0) {
        // dw_out[row] = sum_k W4[row, k] (R^T H3)[row, k] + b4[row] colsum(R)[row]   (H4 = relu(H3 W4^T + b4) is never needed itself)
        const float b4_row = __ldg(fb + 4 * 128 + row);
        float dot = 0.f;
        for (int c0 = ch * 64; c0 < ch * 64 + 64; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_lane + T2_WG4 + c0, v);
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
        flush_wgrad_scaled(t_lane, T2_WG4, nt.g_w[4], row, ch, wo_row);
        if (ch == 0) {
          uint32_t v4[16];
          tmem_ld16(t_lane + T2_WG4 + 128, v4);
          tmem_ld_wait();
          const float cs4 = __uint_as_float(v4[0]);
          if (nt.g_b[4]) atomicAdd(nt.g_b[4] + row, wo_row * cs4);
          dot = fmaf(b4_row, cs4, dot);
        }
        atomicAdd(nt.g_w[5] + row, dot);
      } else {
        flush_wgrad_scaled(t_lane, T2_WG3, nt.g_w[3], row, ch, 1.f);
        if (ch == 0 && nt.g_b[3]) {
          uint32_t v3[16];
          tmem_ld16(t_lane + T2_WG3 + 128, v3);
          tmem_ld_wait();
          atomicAdd(nt.g_b[3] + row, __uint_as_float(v3[0]));
        }
      }
    }
    tc_fence_before();
  }
  // (warps 21-23 of the top role only hand their registers over)
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && nt.g_b[5] && n_my > 0) atomicAdd(nt.g_b[5], s_gbout[0]);
  if (warp == 16) tmem_dealloc(tmem, 512);
}

// =====================================================================================================================
// bottom role: layers 2, 1, 0 (+ latent gradients)
// =====================================================================================================================
constexpr size_t B2_OFF_W1 = 0, B2_OFF_W2 = TILE_BYTES, B2_OFF_BUF = 2 * (size_t)TILE_BYTES;
constexpr size_t B2_OFF_W0LAT = B2_OFF_BUF + 4 * (size_t)TOP_BUF_STRIDE;       // (buffers: 32 KB + a 4 KB constant-1 block each, as in the top role)
constexpr size_t B2_OFF_LAT = B2_OFF_W0LAT + 4096;                          // 8 x 256 f32: one [phase][t] latent-gradient table per (slot, quadrant)
constexpr size_t B2_OFF_TAB = B2_OFF_LAT + 8 * 256 * 4;                         // band weights (32 f32) + latent table (256 f32) for the X0 warps
constexpr size_t B2_OFF_MISC = B2_OFF_TAB + (32 + 256) * 4;                 // acc_rel, acc_ticket, pad
constexpr size_t B2_OFF_BAR = B2_OFF_MISC + 16;
constexpr int B2_N_BAR = 2 + 2 * 9;
constexpr size_t BOT2_SMEM = B2_OFF_BAR + B2_N_BAR * 8 + 16;
constexpr size_t BWD2_SMEM = TOP2_SMEM > BOT2_SMEM ? TOP2_SMEM : BOT2_SMEM;
static_assert(BWD2_SMEM <= 227 * 1024, "backward kernel exceeds the shared memory of an SM");

__device__ __forceinline__ void bwd2_bot_role(const BwdArgs& a, const BwdNet& nt, const long long worker, const long long n_workers) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool onehot = nt.n_latent > 0 && nt.x0.onehot > 0;    // latent gradient through the one-hot columns of X0
  const bool has_lat = nt.n_latent > 0 && !onehot;            // fallback: explicit latent dgrad + scatter by phase
  const int kpad0 = nt.x0.kpad0;
  const int lat_c0 = nt.enc_dim / 8;                                        // first chunk holding latent columns
  const int lat_n = has_lat ? ((nt.enc_dim % 8 + nt.n_latent + 15) / 16) * 16 : 0;   // MMA N covering them
  uint8_t* s_w1 = smem + B2_OFF_W1;
  uint8_t* s_w2 = smem + B2_OFF_W2;
  uint8_t* s_buf = smem + B2_OFF_BUF;
  uint8_t* s_w0lat = smem + B2_OFF_W0LAT;
  float* s_lat = reinterpret_cast<float*>(smem + B2_OFF_LAT);
  const int n_lat_acc = has_lat ? nt.n_phases * nt.n_latent : 0;            // <= 256 checked on the host
  const uint32_t acc_rel = smem_u32(smem + B2_OFF_MISC), acc_ticket = acc_rel + 4;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + B2_OFF_BAR);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + B2_N_BAR);
  const uint32_t bar_w = smem_u32(s_bar), bar_setup = bar_w + 8;
  auto bar_of = [&](int s, int k) { return bar_w + 16u + (uint32_t)(s * 9 + k) * 8u; };
  enum { B_READY = 0, B_ACC, B_WG2, B_WG1, B_WG0, B_LDDZ, B_LDH1, B_LDH0, B_X0 };

  if (warp == 16) {
    if (lane == 0) {
      mbar_init(bar_w, 1);
      mbar_init(bar_setup, 16);
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar_of(s, B_READY), 8); mbar_init(bar_of(s, B_ACC), 1); mbar_init(bar_of(s, B_WG2), 1); mbar_init(bar_of(s, B_WG1), 1);
        mbar_init(bar_of(s, B_WG0), 1); mbar_init(bar_of(s, B_LDDZ), 1); mbar_init(bar_of(s, B_LDH1), 1); mbar_init(bar_of(s, B_LDH0), 1);
        mbar_init(bar_of(s, B_X0), 4);
      }
      mbar_init_fence();
    }
    __syncwarp();
    tmem_alloc(smem_u32(s_tmem), 512);
  }
  // The X0 warps read the band weights and the latent table per row.  From global memory every one of those ~20 loads per row was
  // an L2 round trip: the loaders' flag polls (gpu-scope acquire = CCTL.IVALL) keep the SM's L1 empty.  Measured in-kernel: 6 300
  // cycles to build one X0 tile, with ONE X0 buffer the pacing item of the whole role.  Shared-memory copies, as in the forward.
  float* s_bw = reinterpret_cast<float*>(smem + B2_OFF_TAB);
  float* s_lt = s_bw + 32;
  const bool lt_in_smem = nt.x0.enc.n_phases * nt.x0.enc.n_latent <= 256;
  {
    const EncDesc& e = nt.x0.enc;
    for (int i = threadIdx.x; i < 32; i += blockDim.x) s_bw[i] = (e.band_weight && i < e.n_freq) ? __ldg(e.band_weight + i) : 1.f;
    const int n_lt = e.n_latent > 0 ? e.n_phases * e.n_latent : 0;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lt[i] = (i < n_lt && lt_in_smem) ? __ldg(e.latents + i) : 0.f;
    if (has_lat) for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) s_lat[i] = 0.f;
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x)      // the constant-1 block behind each tile buffer (column 0 == 1 in every row)
      reinterpret_cast<uint4*>(s_buf + (size_t)(i >> 8) * TOP_BUF_STRIDE + TILE_BYTES)[i & 255] =
          ((i & 255) < 128) ? make_uint4(0x00003F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x < 4) reinterpret_cast<uint32_t*>(smem + B2_OFF_MISC)[threadIdx.x] = 0u;
  }
  tc_fence_before();
  fence_proxy_async();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const long long n_my = (a.n_tiles > worker) ? (a.n_tiles - worker + n_workers - 1) / n_workers : 0;
  [[maybe_unused]] int tl_n = 0;     // developer timeline (NERFCA_TIMELINE=bot): regions 1 = epilogue slot 0, 2 = X0 producer, 3 = issuer 0, 0 = loader 0

  if (warp >= 20) {
    // ================= 4 X0 warps: thread = tile row; tiles in the CTA's order, each into its own slot's c2 buffer =================
    reg_dealloc<64>();       // (register budget: see the top role)
    const int row = (warp - 20) * 32 + lane;
    const float* band_w = nt.x0.enc.band_weight ? s_bw : nullptr;
    const float* lat_tab = lt_in_smem ? s_lt : nt.x0.enc.latents;
    RowIn rin;
    {
      const long long p0 = worker * TILE_M + row;
      rin = fetch_row(nt.x0, a.src, p0, n_my > 0 && p0 < a.src.n_points);
    }
    for (long long i = 0; i < n_my; ++i) {
      const RowIn cur = rin;
      const long long pn = (worker + (i + 1) * n_workers) * TILE_M + row;
      rin = fetch_row(nt.x0, a.src, pn, i + 1 < n_my && pn < a.src.n_points);
      NERFCA_TL(warp == 20 && lane == 0, 2001);
      uint8_t* x0_dst = s_buf + (size_t)(2 * (int)(i & 1) + 1) * TOP_BUF_STRIDE;                 // c2 of the tile's slot
      mbar_wait(bar_of((int)(i & 1), B_WG1), (uint32_t)((i >> 1) & 1));                          // weight gradient 1 of THIS tile has read dZ1 from it
      NERFCA_TL(warp == 20 && lane == 0, 2002);
      emit_x0_row(nt.x0, cur, band_w, lat_tab, SmemSink{x0_dst, row}, 0);
      emit_x0_row(nt.x0, cur, band_w, lat_tab, SmemSink{x0_dst, row}, 1);
      warp_publish_smem(bar_of((int)(i & 1), B_X0), lane);
      NERFCA_TL(warp == 20 && lane == 0, 2003);
    }
  } else if (warp >= 16) {
    reg_dealloc<64>();
    if (warp >= 18) {
      // ================= load warp of slot s =================
      const int s = warp - 18;
      const long long n_s = (n_my + 1 - s) / 2;
      const uint32_t c1 = smem_u32(s_buf) + (uint32_t)(2 * s) * TOP_BUF_STRIDE, c2 = c1 + TOP_BUF_STRIDE;
      if (lane == 0) {
        if (s == 0) {
          const uint32_t lat_bytes = has_lat ? (uint32_t)(lat_n / 8) * CHUNK_BYTES : 0u;
          mbar_expect_tx(bar_w, 2 * TILE_BYTES + lat_bytes);
          bulk_g2s(smem_u32(s_w1), nt.pack + nt.w0_bytes, TILE_BYTES, bar_w);
          bulk_g2s(smem_u32(s_w2), nt.pack + nt.w0_bytes + (size_t)TILE_BYTES, TILE_BYTES, bar_w);
          if (has_lat) bulk_g2s(smem_u32(s_w0lat), nt.pack + (size_t)lat_c0 * CHUNK_BYTES, lat_bytes, bar_w);
        }
        uint32_t ring_slot = (uint32_t)((worker + s * n_workers) % a.ring);
        const uint32_t ring_step = (uint32_t)((2 * n_workers) % a.ring);
        for (long long j = 0; j < n_s; ++j) {
          const long long tile = worker + (s + 2 * j) * n_workers;
          const uint8_t* st = nt.stash + (size_t)tile * STASH_STRIDE;
          const uint32_t pj = (uint32_t)(j & 1);
          const uint32_t seen = ld_acquire_gpu(nt.produced + tile);   // requested now, looked at once the buffer is free
          NERFCA_TL(s == 0, 101);
          if (j > 0) mbar_wait(bar_of(s, B_WG0), pj ^ 1);             // weight gradient 0 of the previous tile no longer reads dZ0 (c1) / X0 (c2)
          NERFCA_TL(s == 0, 102);
          if (seen < 8u) wait_flag_ge(nt.produced + tile, 8u);        // all 8 epilogue warps of the top role have written the tile
          NERFCA_TL(s == 0, 103);
          fence_proxy_async_all();
          mbar_expect_tx(bar_of(s, B_LDDZ), TILE_BYTES);
          bulk_g2s(c1, nt.handoff + (size_t)ring_slot * TILE_BYTES, TILE_BYTES, bar_of(s, B_LDDZ));
          ring_slot += ring_step;
          if (ring_slot >= (uint32_t)a.ring) ring_slot -= (uint32_t)a.ring;
          mbar_expect_tx(bar_of(s, B_LDH1), TILE_BYTES);
          bulk_g2s(c2, st + (size_t)TILE_BYTES, TILE_BYTES, bar_of(s, B_LDH1));
          mbar_wait(bar_of(s, B_LDDZ), pj);                           // dZ2 has left its hand-off slot
          NERFCA_TL(s == 0, 104);
          st_release_gpu(nt.consumed + tile, 1u);
          mbar_wait(bar_of(s, B_WG2), pj);                            // weight gradient 2 no longer reads dZ2: its buffer takes H0
          NERFCA_TL(s == 0, 105);
          mbar_expect_tx(bar_of(s, B_LDH0), TILE_BYTES);
          bulk_g2s(c1, st, TILE_BYTES, bar_of(s, B_LDH0));
          if (j + 1 < n_s) {
            const uint8_t* nx = nt.stash + (size_t)(tile + 2 * n_workers) * STASH_STRIDE;
            bulk_prefetch_l2(nx, TILE_BYTES);
            bulk_prefetch_l2(nx + (size_t)TILE_BYTES, TILE_BYTES);
          }
        }
      }
      __syncwarp();
    } else {
      // ================= MMA-issuing warp of slot s =================
      const int s = warp - 16;
      const long long n_s = (n_my + 1 - s) / 2;
      const uint32_t w1 = smem_u32(s_w1), w2 = smem_u32(s_w2), w0lat = smem_u32(s_w0lat);
      const uint32_t c1 = smem_u32(s_buf) + (uint32_t)(2 * s) * TOP_BUF_STRIDE, c2 = c1 + TOP_BUF_STRIDE;
      constexpr uint32_t KK = KSTEP_KMAJOR, KM = KSTEP_MNMAJOR;
      constexpr uint32_t id_dgrad = instr_desc(128, 128, 0, 1), id_wgrad = instr_desc(128, 144, 1, 1);
      const uint32_t id_wg0 = instr_desc(128, kpad0, 1, 1), id_lat = instr_desc(128, lat_n > 0 ? lat_n : 16, 0, 1);
      if (lane == 0) {
        mbar_wait(bar_setup, 0);
        mbar_wait(bar_w, 0);
        tc_fence_after();
        for (long long j = 0; j < n_s; ++j) {
          const uint32_t pj = (uint32_t)(j & 1);
          NERFCA_TL(s == 0, 3001);
          mbar_wait(bar_of(s, B_LDDZ), pj);
          NERFCA_TL(s == 0, 3002);
          acc_acquire(acc_ticket, acc_rel);
          NERFCA_TL(s == 0, 3003);
          umma_k<8, KK, KM>(tmem + B2_ACC, kmajor(c1), mnmajor(w2), id_dgrad, 0);                  // dH1 = dZ2 W2
          umma_commit(bar_of(s, B_ACC));
          NERFCA_TL(s == 0, 3004);
          mbar_wait(bar_of(s, B_LDH1), pj);
          tc_fence_after();
          NERFCA_TL(s == 0, 3005);
          umma_k<8, KM, KM>(tmem + B2_WG2, mnmajor(c1), mnmajor(c2), id_wgrad, 1);                 // WG2 += dZ2^T [H1 | 1]
          umma_commit(bar_of(s, B_WG2));
          NERFCA_TL(s == 0, 3006);
          mbar_wait(bar_of(s, B_READY), 0);                          // step B: dZ1 is in c2
          NERFCA_TL(s == 0, 3011);
          acc_acquire(acc_ticket, acc_rel);
          NERFCA_TL(s == 0, 3012);
          umma_k<8, KK, KM>(tmem + B2_ACC, kmajor(c2), mnmajor(w1), id_dgrad, 0);                  // dH0 = dZ1 W1
          umma_commit(bar_of(s, B_ACC));
          mbar_wait(bar_of(s, B_LDH0), pj);
          tc_fence_after();
          umma_k<8, KM, KM>(tmem + B2_WG1, mnmajor(c2), mnmajor(c1), id_wgrad, 1);                 // WG1 += dZ1^T [H0 | 1]
          umma_commit(bar_of(s, B_WG1));
          NERFCA_TL(s == 0, 3016);
          mbar_wait(bar_of(s, B_READY), 1);                          // step C: dZ0 is in c1
          NERFCA_TL(s == 0, 3021);
          if (has_lat) {
            // latent fallback: issued BEFORE the wait for X0 (it only needs dZ0), so the GEMM and its epilogue run while the X0 warps
            // are still building the tile, and the commit that hands c1 / c2 back to the loader is not held up by the accumulator lock
            acc_acquire(acc_ticket, acc_rel);
            umma_k<8, KK, KM>(tmem + B2_ACC, kmajor(c1), mnmajor(w0lat), id_lat, 0);               // latent columns of dX0
            umma_commit(bar_of(s, B_ACC));
          }
          mbar_wait(bar_of(s, B_X0), pj);
          tc_fence_after();
          NERFCA_TL(s == 0, 3022);
          umma_k<8, KM, KM>(tmem + B2_WG0, mnmajor(c1), mnmajor(c2), id_wg0, 1);                   // WG0 += dZ0^T X0   (X0 sits in c2)
          umma_commit(bar_of(s, B_WG0));                                                           // (covers the latent GEMM's read of c1 too)
          NERFCA_TL(s == 0, 3023);
        }
      }
      __syncwarp();
    }
  } else {
    // ================= 2 x 8 epilogue warps: slot = warp / 8, thread = (row, column half) =================
    reg_alloc<88>();
    const int slot = warp >> 3, q = warp & 3, ch = (warp >> 2) & 1;
    const int row = q * 32 + lane;
    const long long n_s = (n_my + 1 - slot) / 2;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t k_acc = t_lane + B2_ACC + ch * 64;
    uint32_t k_rowoff = (uint32_t)(ch * 8) * CHUNK_BYTES + (uint32_t)row * 16u;
    uint32_t c1 = smem_u32(s_buf) + (uint32_t)(2 * slot) * TOP_BUF_STRIDE + k_rowoff, c2 = c1 + TOP_BUF_STRIDE;
    pin(k_acc); pin(k_rowoff); pin(c1); pin(c2);
    // ---- set-up: zero the weight / bias gradient accumulators (384 columns over the 4 (slot, ch) warp groups)
    tmem_zero(t_lane + B2_WG2 + (uint32_t)(slot * 2 + ch) * 96u, 96);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_setup);
    uint32_t ph_acc = 0;
    for (long long j = 0; j < n_s; ++j) {
      const long long tile = worker + (slot + 2 * j) * n_workers;
      const uint32_t pj = (uint32_t)(j & 1);
      uint32_t va[32], vb[32], w[32];
      [[maybe_unused]] const bool tl_me = warp == 1 && lane == 0;
      // ---- step B: dZ1 = dH1 * 1[H1 > 0] -> c2, over H1 itself once weight gradient 2 has read it
      NERFCA_TL(tl_me, 1001);
      mbar_wait(bar_of(slot, B_ACC), ph_acc); ph_acc ^= 1;
      tc_fence_after();
      NERFCA_TL(tl_me, 1010);
      ld_acc64(k_acc, va, vb);
      acc_release(acc_rel, lane);
      NERFCA_TL(tl_me, 1011);
      mbar_wait(bar_of(slot, B_LDH1), pj);
      masked_grad_pack64(va, vb, c2, w);
      NERFCA_TL(tl_me, 1012);
      mbar_wait(bar_of(slot, B_WG2), pj);
      NERFCA_TL(tl_me, 1013);
      sts_row64(c2, w);
      warp_publish_smem(bar_of(slot, B_READY), lane);
      NERFCA_TL(tl_me, 1014);
      // ---- step C: dZ0 = dH0 * 1[H0 > 0] -> c1, over H0 itself once weight gradient 1 has read it
      int phase = -1;        // latent fallback: the row's phase, requested a whole chain step before it is used (an L2 round trip here:
      if (has_lat) {         // the loaders' flag polls keep invalidating L1)
        const long long p = tile * TILE_M + row;
        if (p < a.src.n_points) phase = load_phase(a.src, p);
      }
      mbar_wait(bar_of(slot, B_ACC), ph_acc); ph_acc ^= 1;
      tc_fence_after();
      ld_acc64(k_acc, va, vb);
      acc_release(acc_rel, lane);
      NERFCA_TL(tl_me, 1021);
      mbar_wait(bar_of(slot, B_LDH0), pj);
      masked_grad_pack64(va, vb, c1, w);
      NERFCA_TL(tl_me, 1022);
      mbar_wait(bar_of(slot, B_WG1), pj);
      NERFCA_TL(tl_me, 1023);
      sts_row64(c1, w);
      warp_publish_smem(bar_of(slot, B_READY), lane);
      NERFCA_TL(tl_me, 1024);
      // ---- latent gradient (fallback): columns [enc_dim, enc_dim + T) of dX0 sit at accumulator columns enc_dim - 8 * lat_c0 + t
      if (has_lat) {
        // The accumulator lock is held only until the latent columns are in registers (the other slot's dgrad is waiting for it);
        // the reduction by phase runs from registers.  Column group ch (16 columns).
        mbar_wait(bar_of(slot, B_ACC), ph_acc); ph_acc ^= 1;
        tc_fence_after();
        uint32_t v[16];
        const int c0 = ch * 16;
        if (c0 < lat_n) {
          tmem_ld16(t_lane + B2_ACC + c0, v);
          tmem_ld_wait();
        }
        acc_release(acc_rel, lane);
        if (c0 < lat_n) {
          // A warp's 32 rows are consecutive samples of one ray, sometimes two (one phase each): for every distinct phase in the warp,
          // sum each latent column over the lanes of that phase by shuffles; lane e keeps column e and adds it to THIS warp's private
          // copy of the [phase][t] table with a plain read-modify-write (the two column groups of a (slot, quadrant) pair own different
          // t).  No atomics: shared-memory float atomics are compare-and-swap loops, and the 32 lanes of a straddling warp contending
          // for two addresses took ~10 000 cycles per tile that the whole slot then waited for (30-phase backward 0.58 -> 0.40 ms, r4g/r4h).
          // The copies are summed once at the end of the launch.
          float* my_lat = s_lat + (slot * 4 + q) * 256;
          const int toff = nt.enc_dim - 8 * lat_c0;
          const bool ok = phase >= 0 && phase < nt.n_phases;
          const int t_me = c0 + lane - toff;
          const bool t_ok = lane < 16 && t_me >= 0 && t_me < nt.n_latent;
          unsigned remaining = __ballot_sync(0xffffffffu, ok);
          while (remaining) {                               // warp-uniform loop: one pass per distinct phase
            const int ph = __shfl_sync(0xffffffffu, phase, __ffs(remaining) - 1);
            const bool in_seg = ok && phase == ph;
            float mine = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int t = c0 + e - toff;
              if (t < 0 || t >= nt.n_latent) continue;      // warp-uniform
              float val = in_seg ? __uint_as_float(v[e]) : 0.f;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
              if (lane == e) mine = val;
            }
            if (t_ok) my_lat[ph * nt.n_latent + t_me] += mine;
            remaining &= ~__ballot_sync(0xffffffffu, in_seg);
          }
          __syncwarp();
        }
      }
    }
    // ---- every MMA of both slots has completed: flush the TMEM-resident accumulators
    if (n_s > 0) mbar_wait(bar_of(slot, B_WG0), (uint32_t)((n_s - 1) & 1));
    tc_fence_before();
    named_bar_sync(1, 512);
    tc_fence_after();
    if (n_my > 0) {
      if (slot == 0) {
        flush_wgrad(t_lane, B2_WG2, nt.g_w[2], row, ch, 128, 128);
        flush_wgrad(t_lane, B2_WG1, nt.g_w[1], row, ch, 128, 128);
        if (ch == 0) {
          uint32_t v[16];
          tmem_ld16(t_lane + B2_WG2 + 128, v);
          tmem_ld_wait();
          if (nt.g_b[2]) atomicAdd(nt.g_b[2] + row, __uint_as_float(v[0]));
          tmem_ld16(t_lane + B2_WG1 + 128, v);
          tmem_ld_wait();
          if (nt.g_b[1]) atomicAdd(nt.g_b[1] + row, __uint_as_float(v[0]));
        }
      } else {
        flush_wgrad(t_lane, B2_WG0, nt.g_w[0], row, ch, kpad0, nt.in_dim);
        if (ch == 1 && nt.g_b[0]) {   // bias 0 = the constant-1 column (index in_dim) of wgrad 0
          const int c0 = nt.in_dim & ~15;
          uint32_t v[16];
          tmem_ld16(t_lane + B2_WG0 + c0, v);
          tmem_ld_wait();
          float val = 0.f;
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c0 + e == nt.in_dim) val = __uint_as_float(v[e]);
          atomicAdd(nt.g_b[0] + row, val);
        }
        if (onehot && nt.g_lat) {
          // S[n][ph] = sum over the samples of phase ph of dZ0[.][n] sits in columns in_dim + 1 + ph of wgrad 0 (all inside the
          // last 16-column group); d latents[ph][t] = sum_n S[n][ph] * W0[n][enc_dim + t].  S goes through shared memory (slot 1's
          // tile buffers are idle now) so that one thread per (ph, t) can run the 128-term dot product.
          float* s_S = reinterpret_cast<float*>(s_buf + 2 * (size_t)TOP_BUF_STRIDE);
          const int cg = kpad0 - 16;
          if (ch == (cg >> 6)) {
            uint32_t v[16];
            tmem_ld16(t_lane + B2_WG0 + cg, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int ph = cg + e - (nt.in_dim + 1);
              if (ph >= 0 && ph < nt.x0.onehot) s_S[ph * 128 + row] = __uint_as_float(v[e]);
            }
          }
          tc_fence_before();
          named_bar_sync(2, 256);
          const int tid = (int)threadIdx.x - 256;
          if (tid < nt.x0.onehot * nt.n_latent) {
            const int ph = tid / nt.n_latent, t = tid - ph * nt.n_latent;
            float g = 0.f;
            for (int n = 0; n < 128; ++n) g = fmaf(s_S[ph * 128 + n], __ldg(nt.w0_f32 + (size_t)n * nt.in_dim + nt.enc_dim + t), g);
            atomicAdd(nt.g_lat + tid, g);
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (nt.g_lat && has_lat)
    for (int i = threadIdx.x; i < n_lat_acc; i += blockDim.x) {
      float g = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) g += s_lat[w8 * 256 + i];
      if (g != 0.f) atomicAdd(nt.g_lat + i, g);
    }
  if (warp == 16) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(BWD2_THREADS, 1) tc_bwd2_kernel(BwdArgs a) {
  const int net_id = blockIdx.x % a.n_nets;
  const int r = blockIdx.x / a.n_nets;
  if (r < a.n_top) bwd2_top_role(a, a.net[net_id], r, a.n_top);
  else bwd2_bot_role(a, a.net[net_id], r - a.n_top, a.n_bot);
}
