// Persistent, warp-specialised bf16 GEMMs for sm_100a:  C[M,N] = epilogue(A[M,K] * W[N,K]^T + bias)
//
// Replaces the nn.Linear calls of the reference hot path (cuBLAS via torch):
//   to_q/to_k/to_v/to_out  reference architecture/transformer_wan.py:60-62,117
//   ffn (FeedForward)      reference architecture/transformer_wan.py:347
//   patch_embedding/proj_out reference architecture/transformer_wan.py:486,537
// with the elementwise tails fused into the epilogue (bias, GELU-tanh, SiLU, gate*y + residual;
// reference transformer_wan.py:336,341,348).
//
// Two kernels, same roles per CTA (256 threads):
//   warp 0        TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, mbarrier ring)
//   warp 1        MMA issuer     (one elected thread, tcgen05.mma kind::f16)
//   warp 2        TMEM allocator (2 accumulator stages)
//   warps 4..7    epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> 16B global stores)
// The accumulator is double buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
//
//   gemm_bf16_kernel<BN>     one CTA per SM, 128 x BN tile, cta_group::1.  Used for small problems.
//   gemm2_bf16_kernel        CTA PAIR (cluster of 2 on one TPC), 256 x 256 tile, cta_group::2: each CTA stages its
//                            own 128 rows of A and HALF of the W tile (128 of the 256 N rows); the pair's tensor
//                            cores read the W halves from both SMs. That cuts shared-memory traffic per MMA from
//                            12 KB to 8 KB per SM — the 1-CTA kernel was shared-memory-bandwidth bound (ncu:
//                            l1tex 83 % busy, tensor pipe 57 %; profiles/r01_gemm_qkv_ncu.txt).
#include "gemm_common.cuh"

namespace fino {

// ================================================================================================
// 1-CTA kernel
// ================================================================================================
template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;
  static constexpr int kBBytes = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;  // 512 or 256 (power of two)
};

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atom
  uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t bars = smem_base + kStages * Cfg::kStageBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], tmem_ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
  uint32_t tmem_ptr_smem = bars + 8u * (2 * kStages + 4);

  // shfl makes the warp index provably warp-uniform for ptxas: role branches become uniform branches and the issue
  // warps keep descriptors / addresses in uniform registers (CUTLASS canonical_warp_idx_sync idiom)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mt, nt;
        tile_coords(tile, p.num_m_tiles, p.num_n_tiles, mt, nt);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_relaxed(empty_bar(stage), phase ^ 1u, 100 + stage);
          uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
          uint32_t b_dst = a_dst + Cfg::kABytes;
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
          tma_load_2d(a_dst, &tmap_a, full_bar(stage), kb * GEMM_BK, mt * GEMM_BM);
          tma_load_2d(b_dst, &tmap_b, full_bar(stage), kb * GEMM_BK, nt * BN);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // whole warp, convergent: one elected lane issues inside the *_w wrappers (see ptx.cuh)
    {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 200 + acc);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, 300 + stage);
          tc_fence_after();
          uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
          uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            uint64_t adesc = make_sdesc_sw128(a_addr + k * 32, 16, 1024);
            uint64_t bdesc = make_sdesc_sw128(b_addr + k * 32, 16, 1024);
            umma_ss_w(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit_w(empty_bar(stage));  // smem slot is free once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit_w(tfull_bar(acc));  // accumulator ready for the epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may touch
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int mt, nt;
      tile_coords(tile, p.num_m_tiles, p.num_n_tiles, mt, nt);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tfull_bar(acc), acc_phase, 400 + acc);
      tc_fence_after();
      const int64_t row = (int64_t)mt * GEMM_BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const float* gate_row = gate_row_ptr(p, row, row_ok);
      const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = nt * BN + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        ResidualChunk rc;
        tmem_ld_32x32b_x32(taddr_row + c * 32, r);
        load_residual_chunk<EPI>(p, row, col0, row_ok, gate_row, rc);
        tmem_wait_ld();
        if (row_ok) epilogue_chunk<EPI>(p, row, col0, gate_row, r, p.bias ? p.bias + col0 : nullptr, rc);
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ================================================================================================
// CTA-pair kernel (cta_group::2). MH = M-halves (128-row accumulators) per CTA:
//   MH = 1: cluster tile 256 x 256, two accumulator stages (epilogue overlaps the next main loop)
//   MH = 2: cluster tile 512 x 256, each CTA owns 256 rows x 256 columns = all 512 TMEM columns, single stage.
//           48 KB of operands per 256x256x64 MACs per SM instead of 32 KB per 128x256x64: the L2->SM fill path
//           (ncu l1tex 83 % busy with MH = 1) stops being the limiter.
// ================================================================================================
template <int MH>
struct Gemm2Cfg {
  static constexpr int BN = 256;                               // N columns per cluster tile
  static constexpr int kStages = (MH == 1) ? 6 : 4;
  static constexpr int kABytes = MH * GEMM_BM * GEMM_BK * 2;   // this CTA's MH*128 rows of A
  static constexpr int kBBytes = (BN / 2) * GEMM_BK * 2;       // this CTA's half (128 rows) of the W tile
  static constexpr int kStageBytes = kABytes + kBBytes;        // 32 KB / 48 KB
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + 8 * 512;  // + per-warp bias staging
  static constexpr int kAccStages = (MH == 1) ? 2 : 1;
  static constexpr int kTmemCols = 512;
  static constexpr int kTileM = 2 * MH * GEMM_BM;              // rows per cluster tile
  static constexpr int kEpiWarps = 4 * MH;                     // one epilogue warp per (M-half, TMEM lane quadrant)
  static constexpr int kThreads = 128 + 32 * kEpiWarps;        // 256 / 384
};

// Work unit -> (cluster tile, K-block range, partial slot). See GemmParams::num_full / splits.
struct GemmUnit {
  int tile, kb0, kb1, part;  // part < 0: whole tile with the fused epilogue
};
__device__ __forceinline__ GemmUnit gemm_unit(const GemmParams& p, int u, int num_kb) {
  GemmUnit w;
  if (u < p.num_full) {
    w.tile = u, w.kb0 = 0, w.kb1 = num_kb, w.part = -1;
  } else {
    const int r = u - p.num_full;
    const int ks = r % p.splits;
    w.tile = p.num_full + r / p.splits;
    w.kb0 = (ks * num_kb) / p.splits;
    w.kb1 = ((ks + 1) * num_kb) / p.splits;
    w.part = r;
  }
  return w;
}

template <int MH, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Cfg<MH>::kThreads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const GemmParams p) {
  using Cfg = Gemm2Cfg<MH>;
  constexpr int kStages = Cfg::kStages;
  constexpr int BN = Cfg::BN;
  constexpr int kAcc = Cfg::kAccStages;
  extern __shared__ uint8_t smem_raw[];
  uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t bars = smem_base + kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };                         // used in the leader CTA only
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };            // per CTA (multicast commit)
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };        // per CTA (multicast commit)
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };   // leader only, 8 warp arrivals
  uint32_t tmem_ptr_smem = bars + 8u * (2 * kStages + 4);

  // shfl makes the warp index provably warp-uniform for ptxas: role branches become uniform branches and the issue
  // warps keep descriptors / addresses in uniform registers (CUTLASS canonical_warp_idx_sync idiom)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;  // cluster tiles
  const int num_units = p.num_full + (num_tiles - p.num_full) * p.splits;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);   // leader's producer arrives once with the byte count of BOTH CTAs
      mbar_init(empty_bar(s), 1);  // one multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * Cfg::kEpiWarps);  // every epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        const GemmUnit wu = gemm_unit(p, u, num_kb);
        int mt, nt;
        tile_coords(wu.tile, p.num_m_tiles, p.num_n_tiles, mt, nt);
        const int row_a = mt * Cfg::kTileM + (int)cta_rank * (MH * GEMM_BM);
        const int row_b = nt * BN + (int)cta_rank * (BN / 2);
        for (int kb = wu.kb0; kb < wu.kb1; ++kb) {
          mbar_wait_relaxed(empty_bar(stage), phase ^ 1u, 100 + stage);
          uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
          uint32_t b_dst = a_dst + Cfg::kABytes;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
          tma_load_2d_2cta(a_dst, &tmap_a, full_bar(stage), kb * GEMM_BK, row_a);  // box {64, MH*128}
          tma_load_2d_2cta(b_dst, &tmap_b, full_bar(stage), kb * GEMM_BK, row_b);  // box {64, 128}
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // whole warp of the leader CTA, convergent: one elected lane issues inside the *_w wrappers (see ptx.cuh)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BM, BN, 0);  // M = 256 across the pair
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = cluster_id; u < num_units; u += num_clusters, ++it) {
        const GemmUnit wu = gemm_unit(p, u, num_kb);
        const int acc = (kAcc == 2) ? (it & 1) : 0;
        const uint32_t acc_phase = (kAcc == 2) ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 200 + acc);
        fence_acq_rel_cluster();  // the arrivals are remote (peer epilogue warps, release.cluster)
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + ((kAcc == 2) ? acc * BN : 0);
        for (int kb = wu.kb0; kb < wu.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase, 300 + stage);
          tc_fence_after();
          uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
          uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            uint64_t bdesc = make_sdesc_sw128(b_addr + k * 32, 16, 1024);
#pragma unroll
            for (int mh = 0; mh < MH; ++mh) {
              uint64_t adesc = make_sdesc_sw128(a_addr + mh * (GEMM_BM * GEMM_BK * 2) + k * 32, 16, 1024);
              umma_ss_2cta_w(tmem_d + mh * BN, adesc, bdesc, idesc, ((kb - wu.kb0) | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit_2cta_w(empty_bar(stage), 0b11);  // frees this stage in BOTH CTAs
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit_2cta_w(tfull_bar(acc), 0b11);  // accumulators ready in both CTAs
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs; warp -> (M-half, TMEM lane quadrant)) =====================
    // Software-pipelined drain: the tcgen05.ld of chunk c+1 is in flight while chunk c is converted and stored, the
    // tile's bias slice is staged in shared memory before the accumulator is ready, and the accumulator is handed
    // back to the MMA warp as soon as the last TMEM read has landed (before the last chunk is written out).
    const int q = warp & 3;
    const int ew = warp - 4;
    const int mh = (MH == 2) ? (ew >> 2) : 0;
    const uint32_t bias_s_u32 = bars + 8u * (2 * kStages + 6) + ew * 512u;  // 512 B per epilogue warp
    const __nv_bfloat16* bias_s = reinterpret_cast<const __nv_bfloat16*>(
        smem_raw + (bias_s_u32 - smem_u32(smem_raw)));
    int it = 0;
    for (int u = cluster_id; u < num_units; u += num_clusters, ++it) {
      const GemmUnit wu = gemm_unit(p, u, num_kb);
      int mt, nt;
      tile_coords(wu.tile, p.num_m_tiles, p.num_n_tiles, mt, nt);
      const int acc = (kAcc == 2) ? (it & 1) : 0;
      const uint32_t acc_phase = (kAcc == 2) ? ((it >> 1) & 1) : (it & 1);
      if (wu.part >= 0) {
        // K-slice of a tail tile: raw fp32 accumulators to the workspace, [part][256 rows][256 columns]
        // (MH == 1 only; gemm_fixup_kernel adds the slices and runs the fused epilogue)
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ((kAcc == 2) ? acc * BN : 0);
        float4* dst = reinterpret_cast<float4*>(p.ws + ((int64_t)wu.part * Cfg::kTileM +
                                                        (int64_t)cta_rank * GEMM_BM + q * 32 + lane) * BN);
        mbar_wait(tfull_bar(acc), acc_phase, 400 + acc);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e)
            dst[c * 8 + e] = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]),
                                         __uint_as_float(r[4 * e + 2]), __uint_as_float(r[4 * e + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
        continue;
      }
      if (p.bias != nullptr) {  // stage bias[nt*256 .. +256) while the main loop is still running
        const int colb = nt * BN + lane * 8;
        uint4 b = make_uint4(0, 0, 0, 0);
        if (colb + 8 <= p.N) b = __ldg(reinterpret_cast<const uint4*>(p.bias + colb));
        else
          for (int j = 0; j < 8; ++j)
            if (colb + j < p.N) reinterpret_cast<__nv_bfloat16*>(&b)[j] = p.bias[colb + j];
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bias_s_u32 + lane * 16), "r"(b.x), "r"(b.y),
                     "r"(b.z), "r"(b.w)
                     : "memory");
        __syncwarp();
      }
      const int64_t row = (int64_t)mt * Cfg::kTileM + (int64_t)cta_rank * (MH * GEMM_BM) + mh * GEMM_BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const float* gate_row = gate_row_ptr(p, row, row_ok);
      const uint32_t taddr_row =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ((kAcc == 2) ? acc * BN : mh * BN);
      const int nchunk = min(BN / 32, (p.N - nt * BN + 31) / 32);  // warp-uniform
      uint32_t buf0[32], buf1[32];
      ResidualChunk rc0, rc1;
      const int colt = nt * BN;
      // the first chunk's residual / gate values do not depend on the accumulator: fetch them while the main loop of
      // this tile is still running (each output element is read and written by its owner thread only, so the
      // in-place form residual == C is safe)
      load_residual_chunk<EPI>(p, row, colt, row_ok, gate_row, rc0);
      mbar_wait(tfull_bar(acc), acc_phase, 400 + acc);
      tc_fence_after();
      auto release_acc = [&]() {  // every TMEM read of this warp is done: hand the accumulator back early
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
      };
      tmem_ld_32x32b_x32(taddr_row, buf0);
      // Two chunks per iteration (static register buffers), not unrolled further: the epilogue body exists twice in
      // the instruction stream instead of eight times (the 8x unrolled version thrashed the instruction cache).
#pragma unroll 1
      for (int c = 0; c < nchunk; c += 2) {
        tmem_wait_ld();  // chunk c has landed in buf0
        if (c + 1 < nchunk) {
          tmem_ld_32x32b_x32(taddr_row + (c + 1) * 32, buf1);
          load_residual_chunk<EPI>(p, row, colt + (c + 1) * 32, row_ok, gate_row, rc1);
        } else {
          release_acc();
        }
        if (row_ok) epilogue_chunk<EPI>(p, row, colt + c * 32, gate_row, buf0, p.bias ? bias_s + c * 32 : nullptr, rc0);
        if (c + 1 < nchunk) {
          tmem_wait_ld();  // chunk c+1 has landed in buf1
          if (c + 2 < nchunk) {
            tmem_ld_32x32b_x32(taddr_row + (c + 2) * 32, buf0);
            load_residual_chunk<EPI>(p, row, colt + (c + 2) * 32, row_ok, gate_row, rc0);
          } else {
            release_acc();
          }
          if (row_ok)
            epilogue_chunk<EPI>(p, row, colt + (c + 1) * 32, gate_row, buf1, p.bias ? bias_s + (c + 1) * 32 : nullptr, rc1);
        }
      }
      __syncwarp();  // bias_s is rewritten for the next tile
    }
  }

  // Neither CTA may exit (or free TMEM) while its peer can still read its shared memory / write its barriers.
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
  }
}

// Adds the K-slices of the tail tiles and applies the fused epilogue. One thread per (row, 8 columns): a warp covers
// one 256-column row segment, so the partial reads (32 B per lane per slice) and the output stores are contiguous.
template <int EPI>
__global__ void __launch_bounds__(256) gemm_fixup_kernel(const GemmParams p) {
  constexpr int TM = 2 * GEMM_BM, BN = 256;
  const int tl = blockIdx.x / (TM / 8);           // tail tile
  const int r = (blockIdx.x % (TM / 8)) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  int mt, nt;
  tile_coords(p.num_full + tl, p.num_m_tiles, p.num_n_tiles, mt, nt);
  const int64_t row = (int64_t)mt * TM + r;
  const int col = nt * BN + lane * 8;
  if (row >= p.M || col >= p.N) return;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  for (int s = 0; s < p.splits; ++s) {
    const float4* src = reinterpret_cast<const float4*>(p.ws + ((int64_t)(tl * p.splits + s) * TM + r) * BN + lane * 8);
    const float4 a = src[0], b = src[1];
    v[0] += a.x, v[1] += a.y, v[2] += a.z, v[3] += a.w;
    v[4] += b.x, v[5] += b.y, v[6] += b.z, v[7] += b.w;
  }
  const float* gate_row = gate_row_ptr(p, row, true);
  float4 g0 = make_float4(1.f, 1.f, 1.f, 1.f), g1 = g0;
  uint4 x = make_uint4(0, 0, 0, 0);
  if (EPI == EPI_GATE_RESIDUAL) {
    x = __ldg(reinterpret_cast<const uint4*>(p.residual + row * p.ldr + col));
    if (gate_row != nullptr) {
      g0 = __ldg(reinterpret_cast<const float4*>(gate_row + col));
      g1 = __ldg(reinterpret_cast<const float4*>(gate_row + col) + 1);
    }
  }
  epilogue_group8<EPI>(p, row, col, v, p.bias ? p.bias + col : nullptr, gate_row != nullptr, g0, g1, x);
}

// Workspace for the K-slices (<= 74 clusters x 256 x 256 fp32 = 19.4 MB; grown on demand), one per (device, stream).
static int gemm_workspace(size_t bytes, cudaStream_t stream, float** out) {
  void* p = nullptr;
  int r = stream_workspace(/*tag=*/1, stream, bytes, &p);
  *out = reinterpret_cast<float*>(p);
  return r;
}

// Split-K plan of the pair kernel (see GemmParams): tiles of 256 x 256 on `clusters` CTA pairs, num_kb K-blocks.
// mode: -1 automatic, 0 never, S >= 2: split every tile of the last (or only) round S ways (test hook).
static int g_gemm_split = -1;
void gemm_set_split(int mode) { g_gemm_split = mode; }
void gemm_plan(int tiles, int num_kb, int clusters, int mode, int* num_full, int* splits) {
  *num_full = tiles;
  *splits = 1;
  if (mode == 0) return;
  const int rem = tiles % clusters;
  if (rem == 0 && mode < 2) return;
  int s;
  if (mode >= 2) {
    s = mode;
  } else {
    s = clusters / rem;
    if (s > 4) s = 4;
  }
  // a slice must be long enough to pay for its exposed partial epilogue and the fix-up launch: measured at M = 3520,
  // N = 3072 (168 tiles, 74 pairs): K = 14336 254 -> 233 us with 3 slices, K = 3072 64 -> 66 us (no gain)
  const int min_kb = mode >= 2 ? 8 : 32;
  if (s > num_kb / min_kb) s = num_kb / min_kb;
  if (s < 2) return;
  *num_full = tiles - (rem == 0 ? clusters < tiles ? clusters : tiles : rem);
  *splits = s;
}

// ================================================================================================
// host side
// ================================================================================================
template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static uint64_t configured = 0;  // one bit per device: the opt-in is a per-device function attribute
  const uint64_t dev_bit = 1ull << (current_device() & 63);
  if (!(configured & dev_bit)) {
    FINO_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
    configured |= dev_bit;
  }
  int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_bf16_kernel<BN, EPI><<<grid, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

template <int MH, int EPI>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<MH>;
  static uint64_t configured = 0;  // one bit per device: the opt-in is a per-device function attribute
  const uint64_t dev_bit = 1ull << (current_device() & 63);
  if (!(configured & dev_bit)) {
    FINO_CHECK_CUDA(cudaFuncSetAttribute(gemm2_bf16_kernel<MH, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
    configured |= dev_bit;
  }
  int tiles = p.num_m_tiles * p.num_n_tiles;
  int units = p.num_full + (tiles - p.num_full) * p.splits;
  int clusters = num_sms() / 2;
  if (units < clusters) clusters = units;
  gemm2_bf16_kernel<MH, EPI><<<2 * clusters, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  FINO_CHECK_CUDA(cudaGetLastError());
  if (p.splits > 1) {
    gemm_fixup_kernel<EPI><<<(tiles - p.num_full) * (Cfg::kTileM / 8), 256, 0, stream>>>(p);
    FINO_CHECK_CUDA(cudaGetLastError());
  }
  return FINO_OK;
}

// The epilogue is a template parameter: one fused tail per kernel instance keeps the epilogue loop small enough for
// the instruction cache (the run-time switch over four inlined tails did not).
template <int EPI>
static int launch_by_shape(int mh, int BN, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                           cudaStream_t stream) {
  if (mh == 2) return launch_gemm2<2, EPI>(ta, tb, p, stream);
  if (mh == 1) return launch_gemm2<1, EPI>(ta, tb, p, stream);
  if (BN == 256) return launch_gemm<256, EPI>(ta, tb, p, stream);
  return launch_gemm<128, EPI>(ta, tb, p, stream);
}

// mode: 0 = auto, 1 = single-CTA kernel, 2 = CTA pair with 256x256 cluster tiles, 3 = CTA pair with 512x256 tiles
static int g_gemm_mode = 0;
void gemm_set_mode(int mode) { g_gemm_mode = mode; }

int gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* c, int64_t ldc,
              int64_t m, int n, int k, int epilogue, int out_fp32, int flags, const void* residual, int64_t ldr,
              const float* gate, int64_t gate_row_stride, const int32_t* row_index, int64_t rows_per_group,
              cudaStream_t stream) {
  FINO_CHECK_ARG(a && w && c, "gemm: null operand pointer");
  FINO_CHECK_ARG(m > 0 && n > 0 && k > 0, "gemm: non-positive shape m=%lld n=%d k=%d", (long long)m, n, k);
  FINO_CHECK_ARG(k % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: k/lda/ldw must be multiples of 8 (16-byte rows)");
  FINO_CHECK_ARG(ldc % 8 == 0, "gemm: ldc must be a multiple of 8");
  FINO_CHECK_ARG(n % 8 == 0, "gemm: n must be a multiple of 8");
  FINO_CHECK_ARG((reinterpret_cast<uintptr_t>(c) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(gate) & 15) == 0,
                 "gemm: c / residual / bias / gate must be 16-byte aligned");
  FINO_CHECK_ARG(epilogue >= EPI_NONE && epilogue <= EPI_GATE_RESIDUAL, "gemm: unknown epilogue %d", epilogue);
  if (epilogue == EPI_GATE_RESIDUAL) {
    FINO_CHECK_ARG(residual != nullptr && ldr % 8 == 0, "gemm: gate-residual epilogue needs a 16B-aligned residual");
    FINO_CHECK_ARG(gate == nullptr || (gate_row_stride % 4 == 0), "gemm: gate row stride must be a multiple of 4");
    FINO_CHECK_ARG(gate == nullptr || row_index != nullptr || rows_per_group > 0,
                   "gemm: gate needs row_index or rows_per_group");
  }
  int mh = 0;  // 0 = single CTA, else M-halves per CTA of the pair kernel
  if (g_gemm_mode == 2) mh = 1;
  else if (g_gemm_mode == 3) mh = 2;
  // auto: the 256x256 pair kernel wins on every large shape measured (profiles/r01_gemm_modes.json); the 512x256
  // variant moves fewer bytes per FLOP but exposes its un-overlapped epilogue and is slower today.
  else if (g_gemm_mode == 0 && n > 128 && m > 256) mh = 1;
  const bool pair = mh != 0;
  const int BN = pair ? 256 : ((n > 128) ? 256 : 128);
  const int BM = pair ? 2 * mh * GEMM_BM : GEMM_BM;
  GemmParams p;
  p.M = m;
  p.N = n;
  p.K = k;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.C = c;
  p.ldc = ldc;
  p.out_fp32 = out_fp32;
  p.epilogue = epilogue;
  p.flags = flags;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.ldr = ldr;
  p.gate = gate;
  p.gate_row_stride = gate_row_stride;
  p.row_index = row_index;
  p.rows_per_group = rows_per_group > 0 ? rows_per_group : (int64_t)1 << 62;
  p.num_m_tiles = (int)((m + BM - 1) / BM);
  p.num_n_tiles = (n + BN - 1) / BN;
  p.num_full = p.num_m_tiles * p.num_n_tiles;
  p.splits = 1;
  p.ws = nullptr;
  if (mh == 1) {  // 256 x 256 pair kernel: split-K of the partly filled last round
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    gemm_plan(tiles, (k + GEMM_BK - 1) / GEMM_BK, num_sms() / 2, g_gemm_split, &p.num_full, &p.splits);
    if (p.splits > 1) {
      const size_t bytes = (size_t)(tiles - p.num_full) * p.splits * 256 * 256 * sizeof(float);
      int r = gemm_workspace(bytes, stream, &p.ws);
      if (r) return r;
    }
  }

  CUtensorMap ta, tb;
  {
    uint64_t dims[2] = {(uint64_t)k, (uint64_t)m};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {GEMM_BK, (uint32_t)(mh == 2 ? 2 * GEMM_BM : GEMM_BM)};
    int r = encode_tmap_bf16(&ta, a, 2, dims, strides, box);
    if (r) return r;
  }
  {
    uint64_t dims[2] = {(uint64_t)k, (uint64_t)n};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {GEMM_BK, (uint32_t)(pair ? BN / 2 : BN)};
    int r = encode_tmap_bf16(&tb, w, 2, dims, strides, box);
    if (r) return r;
  }
  switch (epilogue) {
    case EPI_GELU_TANH: return launch_by_shape<EPI_GELU_TANH>(mh, BN, ta, tb, p, stream);
    case EPI_SILU: return launch_by_shape<EPI_SILU>(mh, BN, ta, tb, p, stream);
    case EPI_GATE_RESIDUAL: return launch_by_shape<EPI_GATE_RESIDUAL>(mh, BN, ta, tb, p, stream);
    default: return launch_by_shape<EPI_NONE>(mh, BN, ta, tb, p, stream);
  }
}

}  // namespace fino
