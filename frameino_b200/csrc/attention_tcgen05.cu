// Non-causal, unmasked flash-attention forward for sm_100a (tcgen05 + TMEM + TMA), head_dim 128 / 64.
//
// Replaces F.scaled_dot_product_attention on the reference hot path:
//   Wan self/cross attention   reference architecture/transformer_wan.py:108-110
//   CogVideoX joint attention  reference architecture/attention_processor.py:2863
//
// Layout contract: Q/K/V/O are token-major, heads contiguous inside a row ([B, N, H*d] views with arbitrary
// row / batch strides, e.g. the three column blocks of a fused QKV GEMM output). TMA gathers the per-head
// [128 x d] tiles straight out of that layout, so no head-major transpose is ever materialised.
//
// CTA = 384 threads, one CTA per SM, 256 query rows per CTA (two 128-row tiles processed ping-pong):
//   warps 0-3   softmax warpgroup for Q tile 0   (thread == query row, TMEM lane == row)
//   warps 4-7   softmax warpgroup for Q tile 1
//   warp  8     TMA producer (Q once, then K/V ring)
//   warp  9     MMA issuer   (S_i = Q_i K_j^T  SS-MMA;  O_i += P_i V_j  TS-MMA with P read from TMEM)
//   warps 10-11 idle (keeps the third warpgroup complete for setmaxnreg)
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,256+d) O1 [384,384+d); P_i (bf16) aliases S_i[0,64).
// MMA issue order per KV tile j:  S0(j), PV1(j-1), S1(j), PV0(j)  -- so the tensor pipe works on one query
// tile while the other tile's softmax runs on the MUFU/FMA pipes.
// head_dim 64 (PSEP): the kernel is bound by the exponentials (256 FLOP per exp instead of 512), and O needs only 64
// columns, so P_i gets its own columns (P0 [320,384), P1 [448,512)) and S_i(j+1) is issued as soon as the softmax
// warps have pulled S_i(j) into registers (s_free barrier) instead of after P_i(j) V_j: the next scores are ready
// before the current exponentials finish and the MUFU pipe never waits on the tensor pipe. Order: S0(j), S1(j),
// PV0(j-1), PV1(j-1); pv_done barriers guard the reuse of P_i and the lazy O rescale.
// Online softmax keeps (m, l) in registers; the O rescale is lazy (only when the running max grows by more
// than 2^8, decided per warp), so the common path never touches O between MMAs.
#include "common.cuh"
#include "ptx.cuh"
#include <type_traits>

namespace fino {

constexpr int ATT_BM = 128;        // query rows per tile (== TMEM lanes)
constexpr int ATT_BN = 128;        // keys per KV tile
constexpr int ATT_THREADS = 384;
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;  // log2 units

constexpr int ATT_MAX_OWNERS = 8;

struct AttnParams {
  // Output rows [g*rows_per_owner, (g+1)*rows_per_owner) live in o[g] (local row = row % rows_per_owner). The plain
  // call uses one owner holding every row; the Ulysses scatter variant passes the peer-mapped O buffers of all ranks.
  __nv_bfloat16* o[ATT_MAX_OWNERS];
  int64_t rows_per_owner;
  int64_t o_row_stride, o_batch_stride;  // elements
  int nq, nk;
  int heads;
  float scale_log2;  // softmax scale * log2(e)
  int num_kv_tiles;
  // Work decomposition (1-D grid). A "tile" is 256 query rows of one (batch, head); tile t = (b*heads + h)*q_tiles + qt.
  // CTAs [0, n_full) each run one whole tile. The remaining tiles (the partial last wave) are split `splits` ways
  // along the KV axis: CTA n_full + r handles KV tiles [s*T/S, (s+1)*T/S) of tile n_full + r/S (s = r%S) and writes an
  // un-normalised fp32 partial (O, m, l) to the workspace; attn_combine_kernel merges them. splits == 1: no partials.
  int q_tiles, n_full, splits;
  int tile_rows;  // query rows per CTA tile: 256 (attn_fwd_kernel) or 512 (attn64x4_fwd_kernel)
  float* ws_o;    // [(tile - n_full)*splits + s][256][HD] fp32
  float2* ws_ml;  // [(tile - n_full)*splits + s][256] (running max in raw score units, row sum)
};

template <int HD>
struct AttnCfg {
  static constexpr int kHalves = HD / 64;                 // 64-column (128 B) swizzle panels per row
  static constexpr int kPanelBytes = 128 * 128;           // 128 rows x 128 B
  static constexpr int kTileBytes = kHalves * kPanelBytes;  // one [128 x HD] bf16 tile
  static constexpr int kKVStages = (HD == 128) ? 2 : 4;
  static constexpr int kQBytes = 2 * kTileBytes;
  static constexpr int kSmemBytes = kQBytes + 2 * kKVStages * kTileBytes + 1024 + 256;
};

// EMU / EMU_B = exponentials per group of 8 (even / odd groups) computed on the FMA/ALU pipes (exp2_emu2) instead of
//         MUFU ex2
// SPLIT = P is published to the MMA warp in two 64-key halves so that P V starts while the second half is still in exp
// PSEP  = P in its own TMEM columns + early S issue (head_dim 64 only, see the header comment)
// PING  = the two softmax warpgroups take turns on the exponential phase (named-barrier token): ptxas paces one warp's
//         MUFU stream at the pipe's own rate (one ex2 per 8 cycles per SM sub-partition), so two warps of a
//         sub-partition inside that phase together only queue up on the MUFU pipe while it then idles during their
//         common load / max / store phases; alternating keeps it busy with one warp while the other does the rest
template <int HD, int EMU, bool SPLIT, bool PSEP, int EMU_B, bool PING>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<HD>;
  static_assert(!PSEP || HD == 64, "separate P columns only fit next to 64-column O accumulators");
  constexpr int KST = Cfg::kKVStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  const uint32_t k_smem = q_smem + Cfg::kQBytes;
  const uint32_t v_smem = k_smem + KST * Cfg::kTileBytes;
  const uint32_t bars = v_smem + KST * Cfg::kTileBytes;
  // barriers: q_full, k_full[KST], k_empty[KST], v_full[KST], v_empty[KST], s_full[2], p_full[2], o_done[2], p_half[2]
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bars + 8u * (1 + KST + s); };
  auto v_full = [&](int s) { return bars + 8u * (1 + 2 * KST + s); };
  auto v_empty = [&](int s) { return bars + 8u * (1 + 3 * KST + s); };
  auto s_full = [&](int i) { return bars + 8u * (1 + 4 * KST + i); };
  auto p_full = [&](int i) { return bars + 8u * (3 + 4 * KST + i); };
  auto o_done = [&](int i) { return bars + 8u * (5 + 4 * KST + i); };
  auto p_half = [&](int i) { return bars + 8u * (7 + 4 * KST + i); };  // first 64 keys of P_i written (SPLIT)
  auto s_free = [&](int i) { return bars + 8u * (9 + 4 * KST + i); };    // S_i pulled into registers (PSEP)
  auto pv_done = [&](int i) { return bars + 8u * (11 + 4 * KST + i); };  // P_i V retired (PSEP)
  const uint32_t tmem_ptr_smem = bars + 8u * (13 + 4 * KST);
  const uint32_t zero_slot = tmem_ptr_smem + 8u;  // holds 0.0f (PING: a load ptxas cannot hoist above the token barrier)

  // shfl makes the warp index provably warp-uniform for ptxas: role branches become uniform branches and the issue
  // warps keep descriptors / addresses in uniform registers (CUTLASS canonical_warp_idx_sync idiom)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // tile / KV-range decode (see AttnParams)
  int tile = blockIdx.x, j0 = 0, T = p.num_kv_tiles;
  const bool partial = (int)blockIdx.x >= p.n_full;
  if (partial) {
    const int r = (int)blockIdx.x - p.n_full;
    const int sp = r % p.splits;
    tile = p.n_full + r / p.splits;
    j0 = (int)(((int64_t)sp * p.num_kv_tiles) / p.splits);
    T = (int)(((int64_t)(sp + 1) * p.num_kv_tiles) / p.splits) - j0;
  }
  const int qt = tile % p.q_tiles;
  const int head = (tile / p.q_tiles) % p.heads;
  const int batch = tile / (p.q_tiles * p.heads);
  const int q0 = qt * (2 * ATT_BM);

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
  }
  if (warp == 9 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KST; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full(i), 1);
      mbar_init(p_full(i), 128);
      mbar_init(p_half(i), 128);
      mbar_init(o_done(i), 1);
      mbar_init(s_free(i), 128);
      mbar_init(pv_done(i), 1);
    }
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(zero_slot), "r"(0u) : "memory");
    fence_barrier_init();
  }
  if (warp == 10) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp >= 8) {
    setmaxnreg_dec<80>();
    if (warp == 8) {
      // ============================ TMA producer ============================
      if (lane == 0) {
        const int c_head = head * HD;
        mbar_arrive_expect_tx(q_full, Cfg::kQBytes);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int hf = 0; hf < Cfg::kHalves; ++hf)
            tma_load_3d(q_smem + i * Cfg::kTileBytes + hf * Cfg::kPanelBytes, &tmap_q, q_full, c_head + hf * 64,
                        q0 + i * ATT_BM, batch);
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < T; ++j) {
          mbar_wait_relaxed(k_empty(stage), phase ^ 1u, 10 + stage);
          mbar_arrive_expect_tx(k_full(stage), Cfg::kTileBytes);
#pragma unroll
          for (int hf = 0; hf < Cfg::kHalves; ++hf)
            tma_load_3d(k_smem + stage * Cfg::kTileBytes + hf * Cfg::kPanelBytes, &tmap_k, k_full(stage),
                        c_head + hf * 64, (j0 + j) * ATT_BN, batch);
          mbar_wait_relaxed(v_empty(stage), phase ^ 1u, 20 + stage);
          mbar_arrive_expect_tx(v_full(stage), Cfg::kTileBytes);
#pragma unroll
          for (int hf = 0; hf < Cfg::kHalves; ++hf)
            tma_load_3d(v_smem + stage * Cfg::kTileBytes + hf * Cfg::kPanelBytes, &tmap_v, v_full(stage),
                        c_head + hf * 64, (j0 + j) * ATT_BN, batch);
          if (++stage == KST) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    } else if (warp == 9) {
      // ============================ MMA issuer ============================
      // The whole warp walks the loop (convergent control flow: descriptors stay in uniform registers); one elected
      // lane issues each tcgen05.mma / commit inside the *_w wrappers.
      {
        constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, 0);  // Q K^T : both K-major
        constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, HD, 1);      // P V   : V is MN-major (d contiguous)
        const uint32_t tmem_s[2] = {tmem_base + 0u, tmem_base + 128u};
        const uint32_t tmem_o[2] = {tmem_base + 256u, tmem_base + 384u};
        const uint32_t tmem_p[2] = {tmem_base + (PSEP ? 320u : 0u), tmem_base + (PSEP ? 448u : 128u)};

        // Descriptors are built once; per KV tile only the stage offset (in 16-byte descriptor units) is added.
        const uint64_t qdesc0 = make_sdesc_sw128(q_smem, 16, 1024);
        const uint64_t kdesc0 = make_sdesc_sw128(k_smem, 16, 1024);
        const uint64_t vdesc0 = make_sdesc_sw128(v_smem, Cfg::kPanelBytes, 1024);  // LBO = panel stride (d 64..127)
        constexpr uint32_t kTileUnits = Cfg::kTileBytes >> 4, kPanelUnits = Cfg::kPanelBytes >> 4;

        auto issue_s = [&](int i, int kstage) {
          const uint64_t qd = qdesc0 + (uint64_t)(i * kTileUnits);
          const uint64_t kd = kdesc0 + (uint64_t)((uint32_t)kstage * kTileUnits);
#pragma unroll
          for (int hf = 0; hf < Cfg::kHalves; ++hf)  // one 64-column panel of d per call, 4 MMAs each
            umma_ss_x4_w(tmem_s[i], qd + hf * kPanelUnits, kd + hf * kPanelUnits, idesc_s, hf != 0 ? 1u : 0u);
        };
        auto issue_pv = [&](int i, int vstage, bool accumulate, int ks0, int ks1) {
          const uint64_t vd = vdesc0 + (uint64_t)((uint32_t)vstage * kTileUnits);
#pragma unroll
          for (int g = ks0 / 4; g < ks1 / 4; ++g)  // 64 keys per call; 16 keys = 16 V rows = 2048 B per MMA
            umma_ts_x4_w(tmem_o[i], tmem_p[i] + g * 32, vd + (uint64_t)(g * 512), idesc_o,
                         (accumulate || g != ks0 / 4) ? 1u : 0u);
        };

        mbar_wait(q_full, 0, 30);
        tc_fence_after();
        int stage = 0;        // K / V stage of tile j
        uint32_t phase = 0;
        int pstage = 0;       // V stage of tile j-1
        if constexpr (PSEP) {
          // P_i V_t for tile t (P published in two halves when SPLIT)
          auto pv_tile = [&](int i, int vstage, int t) {
            if (SPLIT) {
              mbar_wait(p_half(i), (uint32_t)(t & 1), 53 + i);
              tc_fence_after();
              issue_pv(i, vstage, t > 0, 0, 4);
              mbar_wait(p_full(i), (uint32_t)(t & 1), 50 + i);
              tc_fence_after();
              issue_pv(i, vstage, true, 4, 8);
            } else {
              mbar_wait(p_full(i), (uint32_t)(t & 1), 50 + i);
              tc_fence_after();
              issue_pv(i, vstage, t > 0, 0, 8);
            }
            tc_commit_w(pv_done(i));
          };
          uint32_t pphase = 0;
          for (int j = 0; j < T; ++j) {
            mbar_wait(k_full(stage), phase, 40 + stage);
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (j > 0) {  // the softmax warps hold S_i(j-1) in registers: its columns can be overwritten
                mbar_wait(s_free(i), (uint32_t)((j - 1) & 1), 56 + i);
                tc_fence_after();
              }
              issue_s(i, stage);
              tc_commit_w(s_full(i));
            }
            tc_commit_w(k_empty(stage));
            if (j > 0) {
              mbar_wait(v_full(pstage), pphase, 60 + pstage);
              pv_tile(0, pstage, j - 1);
              pv_tile(1, pstage, j - 1);
              tc_commit_w(v_empty(pstage));
            }
            pstage = stage;
            pphase = phase;
            if (++stage == KST) {
              stage = 0;
              phase ^= 1u;
            }
          }
          mbar_wait(v_full(pstage), pphase, 60 + pstage);
          pv_tile(0, pstage, T - 1);
          tc_commit_w(o_done(0));
          pv_tile(1, pstage, T - 1);
          tc_commit_w(v_empty(pstage));
          tc_commit_w(o_done(1));
        } else {
        for (int j = 0; j < T; ++j) {
          mbar_wait(k_full(stage), phase, 40 + stage);
          tc_fence_after();
          issue_s(0, stage);
          tc_commit_w(s_full(0));
          if (j > 0) {
            if (SPLIT) {
              mbar_wait(p_half(1), (uint32_t)((j - 1) & 1), 53);
              tc_fence_after();
              issue_pv(1, pstage, j - 1 > 0, 0, 4);
              mbar_wait(p_full(1), (uint32_t)((j - 1) & 1), 51);
              tc_fence_after();
              issue_pv(1, pstage, true, 4, 8);
            } else {
              mbar_wait(p_full(1), (uint32_t)((j - 1) & 1), 51);
              tc_fence_after();
              issue_pv(1, pstage, j - 1 > 0, 0, 8);
            }
            tc_commit_w(v_empty(pstage));
          }
          issue_s(1, stage);
          tc_commit_w(s_full(1));
          tc_commit_w(k_empty(stage));
          mbar_wait(v_full(stage), phase, 60 + stage);
          if (SPLIT) {
            mbar_wait(p_half(0), (uint32_t)(j & 1), 54);
            tc_fence_after();
            issue_pv(0, stage, j > 0, 0, 4);
            mbar_wait(p_full(0), (uint32_t)(j & 1), 50);
            tc_fence_after();
            issue_pv(0, stage, true, 4, 8);
          } else {
            mbar_wait(p_full(0), (uint32_t)(j & 1), 50);
            tc_fence_after();
            issue_pv(0, stage, j > 0, 0, 8);
          }
          pstage = stage;
          if (++stage == KST) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit_w(o_done(0));
        if (SPLIT) {
          mbar_wait(p_half(1), (uint32_t)((T - 1) & 1), 55);
          tc_fence_after();
          issue_pv(1, pstage, T - 1 > 0, 0, 4);
          mbar_wait(p_full(1), (uint32_t)((T - 1) & 1), 52);
          tc_fence_after();
          issue_pv(1, pstage, true, 4, 8);
        } else {
          mbar_wait(p_full(1), (uint32_t)((T - 1) & 1), 52);
          tc_fence_after();
          issue_pv(1, pstage, T - 1 > 0, 0, 8);
        }
        tc_commit_w(v_empty(pstage));
        tc_commit_w(o_done(1));
        }  // !PSEP
      }
    }
  } else {
    // ============================ softmax warpgroups ============================
    setmaxnreg_inc<208>();
    const int wg = warp >> 2;   // query tile handled by this warpgroup
    const int quad = warp & 3;  // TMEM lane quadrant
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + lane_base + (wg ? 128u : 0u);
    const uint32_t t_o = tmem_base + lane_base + (wg ? 384u : 256u);
    const uint32_t t_p = PSEP ? tmem_base + lane_base + (wg ? 448u : 320u) : t_s;
    const float sl2 = p.scale_log2;

    if (PING && wg == 1) named_bar_arrive(1, 256);  // warpgroup 0 takes the first turn
    int keys_left = p.nk - j0 * ATT_BN;  // keys from this CTA's first KV tile to the end of the sequence
    float m_used = -INFINITY;  // running max (raw score units) the exponentials are referenced to
    float l_sum = 0.f;

    for (int j = 0; j < T; ++j) {
      mbar_wait(s_full(wg), (uint32_t)(j & 1), 70 + wg);
      tc_fence_after();
      uint32_t s[4][32];
      tmem_ld_32x32b_x32(t_s + 0, s[0]);
      tmem_ld_32x32b_x32(t_s + 32, s[1]);
      tmem_ld_32x32b_x32(t_s + 64, s[2]);
      tmem_ld_32x32b_x32(t_s + 96, s[3]);
      tmem_wait_ld();
      if (PSEP) {  // S_i(j) is in registers: the MMA warp may overwrite it with S_i(j+1)
        tc_fence_before();
        mbar_arrive(s_free(wg));
      }

      const int valid = keys_left;  // keys of this tile that exist (>= 1)
      keys_left -= ATT_BN;
      if (valid < ATT_BN) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c * 32 + e >= valid) s[c][e] = 0xff800000u;  // -inf
      }

      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[c][e + 0]));
          mx1 = fmaxf(mx1, __uint_as_float(s[c][e + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[c][e + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(s[c][e + 3]));
        }
      const float tile_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const float m_cand = fmaxf(m_used, tile_max);

      if (j == 0) {
        m_used = m_cand;
      } else {
        const bool need = (m_cand - m_used) * sl2 > ATT_RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {
          // PV_i(j-1) retired before S_i(j) was committed, so O_i is quiescent here (PSEP: wait for it explicitly).
          if (PSEP) {
            mbar_wait(pv_done(wg), (uint32_t)((j - 1) & 1), 72 + wg);
            tc_fence_after();
          }
          const float alpha = need ? ex2_approx((m_used - m_cand) * sl2) : 1.0f;
          if (need) m_used = m_cand;
          l_sum *= alpha;
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(t_o + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st_32x32b_x32(t_o + c * 32, o);
          }
        }
      }

      const float neg_m = -m_used * sl2;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      if (PSEP && j > 0) {  // P_i(j-1) V must have retired before P_i is overwritten
        mbar_wait(pv_done(wg), (uint32_t)((j - 1) & 1), 74 + wg);
        tc_fence_after();
      }
      // The scale-and-subtract of the elements that go through MUFU takes its addend from a shared-memory load placed
      // after the token barrier, so that ptxas keeps the whole MUFU stream inside the turn (it moves plain arithmetic
      // across bar.sync freely); the emulated elements' arithmetic stays free to run ahead of the turn.
      float neg_m_turn = neg_m;
      if (PING) {
        // before the turn: the emulated exponentials (FMA/ALU pipes only), results parked in the score registers.
        // ptxas would sink this arithmetic below the barrier; the barrier id is made to depend on every result (their
        // sign bits, always 0: 2^x > 0) so that it has to be finished first.
        uint32_t dep = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; e += 8)
#pragma unroll
            for (int i = 0; i < (((e >> 3) & 1) ? EMU_B : EMU); i += 2) {
              float x0, x1, p0, p1;
              ffma2(x0, x1, __uint_as_float(s[c][e + i]), __uint_as_float(s[c][e + i + 1]), sl2, sl2, neg_m, neg_m);
              exp2_emu2(p0, p1, x0, x1);
              s[c][e + i] = __float_as_uint(p0);
              s[c][e + i + 1] = __float_as_uint(p1);
              dep |= s[c][e + i] | s[c][e + i + 1];
            }
        named_bar_sync(1 + wg + (int)(dep >> 31), 256);  // my turn on the MUFU pipe
        float z;
        asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(z) : "r"(zero_slot) : "memory");
        neg_m_turn = neg_m + z;
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t pk[32];
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int c = half * 2 + c2;
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            const int emu = ((e >> 3) & 1) ? EMU_B : EMU;
            float x[8], pv[8];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {  // x = s * scale_log2 - m * scale_log2 (packed FFMA2)
              if (PING && i < emu) continue;  // already exponentiated
              const float nm = (i < emu) ? neg_m : neg_m_turn;
              ffma2(x[i], x[i + 1], __uint_as_float(s[c][e + i]), __uint_as_float(s[c][e + i + 1]), sl2, sl2, nm, nm);
            }
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              if (i < emu) {
                if (PING) {
                  pv[i] = __uint_as_float(s[c][e + i]);
                  pv[i + 1] = __uint_as_float(s[c][e + i + 1]);
                } else {
                  exp2_emu2(pv[i], pv[i + 1], x[i], x[i + 1]);
                }
              } else {
                pv[i] = ex2_approx(x[i]);
                pv[i + 1] = ex2_approx(x[i + 1]);
              }
            }
            fadd2(sum0, sum1, sum0, sum1, pv[0], pv[1]);
            fadd2(sum2, sum3, sum2, sum3, pv[2], pv[3]);
            fadd2(sum0, sum1, sum0, sum1, pv[4], pv[5]);
            fadd2(sum2, sum3, sum2, sum3, pv[6], pv[7]);
#pragma unroll
            for (int i = 0; i < 8; i += 2) pk[c2 * 16 + (e + i) / 2] = pack_bf16x2(pv[i], pv[i + 1]);
          }
        }
        // P (bf16, 2 keys per 32-bit column) aliases S_i columns [0,64) / has its own columns (PSEP)
        if (PING && half == 1) named_bar_arrive(1 + (wg ^ 1), 256);  // exponentials issued: the other group's turn
        tmem_st_32x32b_x32(t_p + half * 32, pk);
        if (SPLIT && half == 0) {
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(p_half(wg));
        }
      }
      l_sum += (sum0 + sum1) + (sum2 + sum3);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(p_full(wg));
    }

    if (PING && wg == 0) named_bar_sync(1, 256);  // absorb warpgroup 1's last hand-over (barrier left balanced)
    // ---------------- epilogue: O / l -> bf16 -> shared (swizzled) -> coalesced global rows ----------------
    // Every S_i MMA has retired when o_done(i) fires, so the Q_i tile in shared memory is dead: each warp stages its
    // 32 output rows there (16-byte chunks XOR-swizzled by row, conflict-free both ways) and then writes whole
    // head_dim*2-byte row segments with consecutive lanes — 256-byte (d=128) contiguous stores instead of 32 scattered
    // 16-byte ones, which matters doubly when the destination is a peer GPU's buffer behind NVLink.
    mbar_wait(o_done(wg), 0, 80 + wg);
    tc_fence_after();
    if (partial) {
      // KV-split CTA: un-normalised fp32 partial (O, m, l) to the workspace; attn_combine_kernel finishes the rows.
      // (thread == row, so the stores are 16-byte pieces strided by a row; <= 148 CTAs x 128 KB per launch.)
      const int64_t prow = (int64_t)((int)blockIdx.x - p.n_full) * (2 * ATT_BM) + wg * ATT_BM + row_in_tile;
      float4* dst = reinterpret_cast<float4*>(p.ws_o + prow * HD);
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(t_o + c * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e)
          dst[c * 8 + e] = make_float4(__uint_as_float(o[4 * e]), __uint_as_float(o[4 * e + 1]),
                                       __uint_as_float(o[4 * e + 2]), __uint_as_float(o[4 * e + 3]));
      }
      p.ws_ml[prow] = make_float2(m_used, l_sum);
    } else {
    const float inv_l = 1.0f / l_sum;
    constexpr int CH = HD / 8;            // 16-byte chunks per output row
    constexpr int RPI = 32 / CH;          // rows per store instruction
    const uint32_t stage = q_smem + wg * Cfg::kTileBytes + (uint32_t)(quad * 32) * (HD * 2);
    const uint32_t my_row = stage + (uint32_t)lane * (HD * 2);
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_o + c * 32, o);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t w0 = pack_bf16x2(__uint_as_float(o[8 * e + 0]) * inv_l, __uint_as_float(o[8 * e + 1]) * inv_l);
        const uint32_t w1 = pack_bf16x2(__uint_as_float(o[8 * e + 2]) * inv_l, __uint_as_float(o[8 * e + 3]) * inv_l);
        const uint32_t w2 = pack_bf16x2(__uint_as_float(o[8 * e + 4]) * inv_l, __uint_as_float(o[8 * e + 5]) * inv_l);
        const uint32_t w3 = pack_bf16x2(__uint_as_float(o[8 * e + 6]) * inv_l, __uint_as_float(o[8 * e + 7]) * inv_l);
        const uint32_t chunk = (uint32_t)(c * 4 + e) ^ (uint32_t)(lane & (CH - 1) & 7);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + chunk * 16), "r"(w0), "r"(w1), "r"(w2),
                     "r"(w3)
                     : "memory");
      }
    }
    __syncwarp();
    const int sub = lane / CH;   // row inside the group of RPI rows written by one instruction
    const int ch = lane % CH;    // 16-byte chunk inside the row
#pragma unroll 4
    for (int it = 0; it < 32 / RPI; ++it) {
      const int r = it * RPI + sub;  // row inside this warp's 32
      const int qrow = q0 + wg * ATT_BM + quad * 32 + r;
      uint4 w;
      const uint32_t src = stage + (uint32_t)r * (HD * 2) + (((uint32_t)ch ^ (uint32_t)(r & (CH - 1) & 7)) * 16);
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(src));
      if (qrow < p.nq) {
        const int64_t owner = (int64_t)qrow / p.rows_per_owner;
        const int64_t lrow = (int64_t)qrow - owner * p.rows_per_owner;
        __nv_bfloat16* dst =
            p.o[owner] + (int64_t)batch * p.o_batch_stride + lrow * p.o_row_stride + (int64_t)head * HD + ch * 8;
        *reinterpret_cast<uint4*>(dst) = w;
      }
    }
    }  // !partial
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Merges the KV-split partials of the tiles past n_full: one warp per query row, lane = 4 (d=128) / 2 (d=64) columns.
//   M = max_s m_s;  w_s = 2^((m_s - M) * scale_log2);  O = sum_s w_s O_s / sum_s w_s l_s
template <int HD>
__global__ void __launch_bounds__(256) attn_combine_kernel(const __grid_constant__ AttnParams p, int rows_total) {
  const int gw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= rows_total) return;
  const int tl = gw / p.tile_rows, r = gw % p.tile_rows;
  const int tile = p.n_full + tl;
  const int qt = tile % p.q_tiles;
  const int head = (tile / p.q_tiles) % p.heads;
  const int batch = tile / (p.q_tiles * p.heads);
  const int qrow = qt * p.tile_rows + r;
  if (qrow >= p.nq) return;
  constexpr int CPL = HD / 32;  // columns per lane
  const int S = p.splits;
  float mmax = -INFINITY;
  for (int s = 0; s < S; ++s) mmax = fmaxf(mmax, p.ws_ml[(int64_t)(tl * S + s) * p.tile_rows + r].x);
  float acc[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) acc[i] = 0.f;
  float l = 0.f;
  for (int s = 0; s < S; ++s) {
    const int64_t prow = (int64_t)(tl * S + s) * p.tile_rows + r;
    const float2 ml = p.ws_ml[prow];
    const float w = ex2_approx((ml.x - mmax) * p.scale_log2);
    l += w * ml.y;
    const float* src = p.ws_o + prow * HD + lane * CPL;
    if (CPL == 4) {
      const float4 v = *reinterpret_cast<const float4*>(src);
      acc[0] += w * v.x; acc[1] += w * v.y; acc[2] += w * v.z; acc[CPL - 1] += w * v.w;
    } else {
      const float2 v = *reinterpret_cast<const float2*>(src);
      acc[0] += w * v.x; acc[1] += w * v.y;
    }
  }
  const float inv_l = 1.0f / l;
  const int64_t owner = (int64_t)qrow / p.rows_per_owner;
  const int64_t lrow = (int64_t)qrow - owner * p.rows_per_owner;
  __nv_bfloat16* dst =
      p.o[owner] + (int64_t)batch * p.o_batch_stride + lrow * p.o_row_stride + (int64_t)head * HD + lane * CPL;
  if (CPL == 4) {
    uint2 w2;
    w2.x = pack_bf16x2(acc[0] * inv_l, acc[1] * inv_l);
    w2.y = pack_bf16x2(acc[2] * inv_l, acc[CPL - 1] * inv_l);
    *reinterpret_cast<uint2*>(dst) = w2;
  } else {
    *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(acc[0] * inv_l, acc[1] * inv_l);
  }
}

// Workspace for the KV-split partials (grown on demand; <= 148 CTAs x 256 rows x (HD*4 + 8) B = 19.7 MB in the
// automatic mode), one per (device, stream): consecutive launches on a stream reuse it in stream order.
static int attn_workspace(size_t bytes, cudaStream_t stream, void** out) {
  return stream_workspace(/*tag=*/2, stream, bytes, out);
}

template <int HD, int EMU, bool SPLIT, bool PSEP = false, int EMU_B = EMU, bool PING = false>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                       int batch, cudaStream_t stream) {
  using Cfg = AttnCfg<HD>;
  static uint64_t configured = 0;  // one bit per device: the opt-in is a per-device function attribute
  const uint64_t dev_bit = 1ull << (current_device() & 63);
  if (!(configured & dev_bit)) {
    FINO_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HD, EMU, SPLIT, PSEP, EMU_B, PING>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured |= dev_bit;
  }
  const int n_tiles = p.q_tiles * p.heads * batch;
  const int n_part = (n_tiles - p.n_full) * p.splits;
  attn_fwd_kernel<HD, EMU, SPLIT, PSEP, EMU_B, PING>
      <<<p.n_full + n_part, ATT_THREADS, Cfg::kSmemBytes, stream>>>(tq, tk, tv, p);
  FINO_CHECK_CUDA(cudaGetLastError());
  if (n_part > 0) {
    const int rows_total = (n_tiles - p.n_full) * p.tile_rows;
    attn_combine_kernel<HD><<<(rows_total + 7) / 8, 256, 0, stream>>>(p, rows_total);
    FINO_CHECK_CUDA(cudaGetLastError());
  }
  return FINO_OK;
}

// =====================================================================================================================
// head_dim 64, four query tiles per CTA against 64-key steps (CogVideoX: 48 heads x 64).
//
// At head_dim 64 the kernel is bound by the exponentials, not by the tensor pipe (256 FLOP per exp; measured pipe
// rates in tools/microbench/pipes.cu: one ex2 warp-instruction per 8 cycles per SM sub-partition), and in the two-tile
// kernel above each softmax warp's serial chain per tile (TMEM load, max, exponentials, pack, TMEM store, then the
// P V / next-S round trip through the MMA warp) is far longer than two warps per sub-partition can hide (ncu: MUFU
// pipe 58-62 % busy, profiles/r01_attn_d64_*). This kernel runs FOUR 128-row query tiles per CTA against 64-key steps,
// i.e. four softmax warps per sub-partition, each with a half-length chain, served round-robin by the MMA warp:
//   warps 0-15  softmax warpgroups, one per query tile (thread == query row, TMEM lane == row)
//   warp  16    TMA producer (Q tiles once, then a 6-stage ring of [64 keys x 64] K and V tiles)
//   warp  17    TMEM allocator + MMA issuer: per step j, for each tile i:  O_i += P_i(j-1) V(j-1);  S_i(j) = Q_i K(j)^T
//   warps 18-19 spare (the fifth warpgroup exists so that setmaxnreg can move its registers: the kernel launches at 96
//               registers per thread, the fifth group drops to 32 and the softmax warps rise to 112)
// TMEM (512 columns): tile i owns [128 i, 128 i + 128): S_i fp32 [0,64), P_i (bf16) aliases S_i [0,32), O_i [64,128).
// The softmax step reads the scores in two passes (row maximum over all 64, then the exponentials 32 at a time with
// the second half read from TMEM again): holding 64 scores next to the packed P and the MUFU latency window spilled.
// The ragged last step of the sequence is a separate instantiation of the step body, so the common path has no
// conditional writes to the score registers.
// Measured at the CogVideoX shape (19 126 x 19 126, 48 heads), isolated: 956 TFLOP/s, 980 with the tile pairs started
// half a period apart (STAGGER) (two-tile kernel: 820; with P in
// its own columns + MUFU turn-taking: 885). Dead ends kept out of the tree: three tiles with P in its own columns and
// early S issue (794: three warps per sub-partition hide less than four), the same with a polling MMA scheduler
// instead of the fixed round-robin (527: a dozen mbarrier probes per pass put ~400 cycles into every round trip).
// Work decomposition, partials and the output-owner scatter are those of attn_fwd_kernel with 512-row tiles.
// =====================================================================================================================
constexpr int A64_NT = 4;
constexpr int A64_BN = 64;
constexpr int A64_THREADS = 640;
constexpr int A64_KST = 6;
constexpr int A64_QTILE_BYTES = ATT_BM * 64 * 2;   // 16 KB
constexpr int A64_KVTILE_BYTES = A64_BN * 64 * 2;  // 8 KB
constexpr int A64_SMEM_BYTES = A64_NT * A64_QTILE_BYTES + 2 * A64_KST * A64_KVTILE_BYTES + 1024 + 512;
constexpr int A64_SOFTMAX_WARPS = 4 * A64_NT;
__host__ __device__ constexpr uint32_t a64_col_s(int i) { return 128u * i; }
__host__ __device__ constexpr uint32_t a64_col_p(int i) { return 128u * i; }
__host__ __device__ constexpr uint32_t a64_col_o(int i) { return 128u * i + 64u; }

template <int EMU, int EMU_B, bool STAGGER>
__global__ void __launch_bounds__(A64_THREADS, 1)
attn64x4_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ AttnParams p) {
  constexpr int HD = 64;
  constexpr int KST = A64_KST;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  const uint32_t k_smem = q_smem + A64_NT * A64_QTILE_BYTES;
  const uint32_t v_smem = k_smem + KST * A64_KVTILE_BYTES;
  const uint32_t bars = v_smem + KST * A64_KVTILE_BYTES;
  const uint32_t q_full = bars;
  auto k_full = [&](int s) { return bars + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bars + 8u * (1 + KST + s); };
  auto v_full = [&](int s) { return bars + 8u * (1 + 2 * KST + s); };
  auto v_empty = [&](int s) { return bars + 8u * (1 + 3 * KST + s); };
  auto s_full = [&](int i) { return bars + 8u * (1 + 4 * KST + i); };
  auto p_full = [&](int i) { return bars + 8u * (1 + 4 * KST + A64_NT + i); };
  auto o_done = [&](int i) { return bars + 8u * (1 + 4 * KST + 2 * A64_NT + i); };
  const uint32_t tmem_ptr_smem = bars + 8u * (1 + 4 * KST + 3 * A64_NT);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  int tile = blockIdx.x, j0 = 0, T = p.num_kv_tiles;
  const bool partial = (int)blockIdx.x >= p.n_full;
  if (partial) {
    const int r = (int)blockIdx.x - p.n_full;
    const int sp = r % p.splits;
    tile = p.n_full + r / p.splits;
    j0 = (int)(((int64_t)sp * p.num_kv_tiles) / p.splits);
    T = (int)(((int64_t)(sp + 1) * p.num_kv_tiles) / p.splits) - j0;
  }
  const int qt = tile % p.q_tiles;
  const int head = (tile / p.q_tiles) % p.heads;
  const int batch = tile / (p.q_tiles * p.heads);
  const int q0 = qt * (A64_NT * ATT_BM);

  if (warp == A64_SOFTMAX_WARPS && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
  }
  if (warp == A64_SOFTMAX_WARPS + 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KST; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    for (int i = 0; i < A64_NT; ++i) {
      mbar_init(s_full(i), 1);
      mbar_init(p_full(i), 128);
      mbar_init(o_done(i), 1);
    }
    fence_barrier_init();
  }
  if (warp == A64_SOFTMAX_WARPS + 1) {
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp >= A64_SOFTMAX_WARPS) setmaxnreg_dec<32>();  // 4 warps x 64 registers go to the 16 softmax warps (96 -> 112)
  if (warp == A64_SOFTMAX_WARPS) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      const int c_head = head * HD;
      mbar_arrive_expect_tx(q_full, A64_NT * A64_QTILE_BYTES);
#pragma unroll
      for (int i = 0; i < A64_NT; ++i)
        tma_load_3d(q_smem + i * A64_QTILE_BYTES, &tmap_q, q_full, c_head, q0 + i * ATT_BM, batch);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < T; ++j) {
        mbar_wait_relaxed(k_empty(stage), phase ^ 1u, 10 + stage);
        mbar_arrive_expect_tx(k_full(stage), A64_KVTILE_BYTES);
        tma_load_3d(k_smem + stage * A64_KVTILE_BYTES, &tmap_k, k_full(stage), c_head, (j0 + j) * A64_BN, batch);
        mbar_wait_relaxed(v_empty(stage), phase ^ 1u, 20 + stage);
        mbar_arrive_expect_tx(v_full(stage), A64_KVTILE_BYTES);
        tma_load_3d(v_smem + stage * A64_KVTILE_BYTES, &tmap_v, v_full(stage), c_head, (j0 + j) * A64_BN, batch);
        if (++stage == KST) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == A64_SOFTMAX_WARPS + 1) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, A64_BN, 0);  // Q K^T : both K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, HD, 1);      // P V   : V is MN-major (d contiguous)
    const uint64_t qdesc0 = make_sdesc_sw128(q_smem, 16, 1024);
    const uint64_t kdesc0 = make_sdesc_sw128(k_smem, 16, 1024);
    const uint64_t vdesc0 = make_sdesc_sw128(v_smem, A64_KVTILE_BYTES, 1024);
    constexpr uint32_t kQUnits = A64_QTILE_BYTES >> 4, kKVUnits = A64_KVTILE_BYTES >> 4;

    mbar_wait(q_full, 0, 30);
    mbar_wait(k_full(0), 0, 40);
    tc_fence_after();
#pragma unroll
    for (int i = 0; i < A64_NT; ++i) {
      if (STAGGER && i == A64_NT / 2) {
        // De-phase the two tile pairs: tiles 2,3 get their first scores only when tile 0 has finished its first
        // softmax step, i.e. exactly when tiles 0,1 go idle for their P V -> S round trip. Every tile has the same
        // period, so the offset persists for the whole KV loop: one pair's exponentials fill the MUFU pipe while the
        // other pair waits for the tensor pipe, and the round-robin service order below matches their readiness.
        mbar_wait(p_full(0), 0u, 45);
        tc_fence_after();
      }
      umma_ss_x4_w(tmem_base + a64_col_s(i), qdesc0 + (uint64_t)(i * kQUnits), kdesc0, idesc_s, 0u);
      tc_commit_w(s_full(i));
    }
    tc_commit_w(k_empty(0));
    int stage = 1 % KST;  // K stage of step j
    uint32_t phase = 0;
    int pstage = 0;       // V stage of step j-1
    uint32_t pphase = 0;
    for (int j = 1; j < T; ++j) {
      mbar_wait(k_full(stage), phase, 40 + stage);
      mbar_wait(v_full(pstage), pphase, 60 + pstage);
      tc_fence_after();
      const uint64_t kd = kdesc0 + (uint64_t)((uint32_t)stage * kKVUnits);
      const uint64_t vd = vdesc0 + (uint64_t)((uint32_t)pstage * kKVUnits);
#pragma unroll
      for (int i = 0; i < A64_NT; ++i) {
        mbar_wait(p_full(i), (uint32_t)((j - 1) & 1), 50 + i);
        tc_fence_after();
        umma_ts_x4_w(tmem_base + a64_col_o(i), tmem_base + a64_col_p(i), vd, idesc_o, j - 1 > 0 ? 1u : 0u);
        umma_ss_x4_w(tmem_base + a64_col_s(i), qdesc0 + (uint64_t)(i * kQUnits), kd, idesc_s, 0u);
        tc_commit_w(s_full(i));
      }
      tc_commit_w(k_empty(stage));
      tc_commit_w(v_empty(pstage));
      pstage = stage;
      pphase = phase;
      if (++stage == KST) {
        stage = 0;
        phase ^= 1u;
      }
    }
    mbar_wait(v_full(pstage), pphase, 60 + pstage);
    tc_fence_after();
    {
      const uint64_t vd = vdesc0 + (uint64_t)((uint32_t)pstage * kKVUnits);
#pragma unroll
      for (int i = 0; i < A64_NT; ++i) {
        mbar_wait(p_full(i), (uint32_t)((T - 1) & 1), 50 + i);
        tc_fence_after();
        umma_ts_x4_w(tmem_base + a64_col_o(i), tmem_base + a64_col_p(i), vd, idesc_o, T - 1 > 0 ? 1u : 0u);
        tc_commit_w(o_done(i));
      }
      tc_commit_w(v_empty(pstage));
    }
  } else if (warp < A64_SOFTMAX_WARPS) {
    // ============================ softmax warpgroups ============================
    setmaxnreg_inc<112>();
    const int wg = warp >> 2;   // query tile handled by this warpgroup
    const int quad = warp & 3;  // TMEM lane quadrant
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + lane_base + a64_col_s(wg);
    const uint32_t t_p = tmem_base + lane_base + a64_col_p(wg);
    const uint32_t t_o = tmem_base + lane_base + a64_col_o(wg);
    const float sl2 = p.scale_log2;

    int keys_left = p.nk - j0 * A64_BN;
    float m_used = -INFINITY;
    float l_sum = 0.f;

    // One KV step. Two instantiations: the ragged last step of the sequence (keys past nk get -inf) is a separate copy
    // of the body, so that the common path has no conditional writes to the score registers (a branch around the
    // masking makes ptxas keep the scores in local memory).
    auto step = [&](int j, auto ragged) {
      constexpr bool kRagged = decltype(ragged)::value;
      mbar_wait(s_full(wg), (uint32_t)(j & 1), 70 + wg);
      tc_fence_after();
      // Pass 1: row maximum over the 64 scores. Only the first 32 stay in registers; the second half is read from TMEM
      // again for its exponentials (a second 4 KB tcgen05.ld per warp is cheaper than holding 64 scores next to the
      // packed P, the exponential temporaries and the MUFU latency window in 112 registers — that version spilled).
      uint32_t s0[32];
      float tile_max;
      {
        uint32_t s1[32];
        tmem_ld_32x32b_x32(t_s + 0, s0);
        tmem_ld_32x32b_x32(t_s + 32, s1);
        tmem_wait_ld();
        if constexpr (kRagged) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (e >= keys_left) s0[e] = 0xff800000u;  // -inf
            if (32 + e >= keys_left) s1[e] = 0xff800000u;
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s0[e + 0]), __uint_as_float(s1[e + 0])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s0[e + 1]), __uint_as_float(s1[e + 1])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(s0[e + 2]), __uint_as_float(s1[e + 2])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(s0[e + 3]), __uint_as_float(s1[e + 3])));
        }
        tile_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      }
      const float m_cand = fmaxf(m_used, tile_max);

      if (j == 0) {
        m_used = m_cand;
      } else {
        const bool need = (m_cand - m_used) * sl2 > ATT_RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {
          // P_i V (j-1) was issued before S_i(j) and retired before s_full fired (in-order pipe): O_i is quiescent.
          const float alpha = need ? ex2_approx((m_used - m_cand) * sl2) : 1.0f;
          if (need) m_used = m_cand;
          l_sum *= alpha;
#pragma unroll
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(t_o + c * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st_32x32b_x16(t_o + c * 16, o);
          }
        }
      }

      const float neg_m = -m_used * sl2;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c == 1) {  // pass 2 for the second half: scores [32,64) again
          tmem_ld_32x32b_x32(t_s + 32, s0);
          tmem_wait_ld();
          if constexpr (kRagged) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (32 + e >= keys_left) s0[e] = 0xff800000u;
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          const int emu = ((e >> 3) & 1) ? EMU_B : EMU;
          float x[8], pv[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2)
            ffma2(x[i], x[i + 1], __uint_as_float(s0[e + i]), __uint_as_float(s0[e + i + 1]), sl2, sl2, neg_m, neg_m);
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            if (i < emu) {
              exp2_emu2(pv[i], pv[i + 1], x[i], x[i + 1]);
            } else {
              pv[i] = ex2_approx(x[i]);
              pv[i + 1] = ex2_approx(x[i + 1]);
            }
          }
          fadd2(sum0, sum1, sum0, sum1, pv[0], pv[1]);
          fadd2(sum2, sum3, sum2, sum3, pv[2], pv[3]);
          fadd2(sum0, sum1, sum0, sum1, pv[4], pv[5]);
          fadd2(sum2, sum3, sum2, sum3, pv[6], pv[7]);
#pragma unroll
          for (int i = 0; i < 8; i += 2) pk[(e + i) / 2] = pack_bf16x2(pv[i], pv[i + 1]);
        }
        // P (bf16, 2 keys per 32-bit column) over S_i columns [0,32), 16 columns per half. Half 0 lands on score
        // columns that are already in registers; the scores of half 1 sit in columns [32,64), untouched by it.
        tmem_st_32x32b_x16(t_p + c * 16, pk);
      }
      l_sum += (sum0 + sum1) + (sum2 + sum3);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(p_full(wg));
    };
    for (int j = 0; j < T; ++j) {
      if (keys_left >= A64_BN)
        step(j, std::false_type{});
      else
        step(j, std::true_type{});
      keys_left -= A64_BN;
    }

    // ---------------- epilogue (as attn_fwd_kernel: staged through the dead Q tile, coalesced row stores) -----------
    mbar_wait(o_done(wg), 0, 80 + wg);
    tc_fence_after();
    if (partial) {
      const int64_t prow =
          (int64_t)((int)blockIdx.x - p.n_full) * (A64_NT * ATT_BM) + wg * ATT_BM + row_in_tile;
      float4* dst = reinterpret_cast<float4*>(p.ws_o + prow * HD);
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(t_o + c * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e)
          dst[c * 8 + e] = make_float4(__uint_as_float(o[4 * e]), __uint_as_float(o[4 * e + 1]),
                                       __uint_as_float(o[4 * e + 2]), __uint_as_float(o[4 * e + 3]));
      }
      p.ws_ml[prow] = make_float2(m_used, l_sum);
    } else {
      const float inv_l = 1.0f / l_sum;
      constexpr int CH = HD / 8;    // 16-byte chunks per output row
      constexpr int RPI = 32 / CH;  // rows per store instruction
      const uint32_t stage = q_smem + wg * A64_QTILE_BYTES + (uint32_t)(quad * 32) * (HD * 2);
      const uint32_t my_row = stage + (uint32_t)lane * (HD * 2);
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(t_o + c * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t w0 = pack_bf16x2(__uint_as_float(o[8 * e + 0]) * inv_l, __uint_as_float(o[8 * e + 1]) * inv_l);
          const uint32_t w1 = pack_bf16x2(__uint_as_float(o[8 * e + 2]) * inv_l, __uint_as_float(o[8 * e + 3]) * inv_l);
          const uint32_t w2 = pack_bf16x2(__uint_as_float(o[8 * e + 4]) * inv_l, __uint_as_float(o[8 * e + 5]) * inv_l);
          const uint32_t w3 = pack_bf16x2(__uint_as_float(o[8 * e + 6]) * inv_l, __uint_as_float(o[8 * e + 7]) * inv_l);
          const uint32_t chunk = (uint32_t)(c * 4 + e) ^ (uint32_t)(lane & (CH - 1) & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + chunk * 16), "r"(w0), "r"(w1),
                       "r"(w2), "r"(w3)
                       : "memory");
        }
      }
      __syncwarp();
      const int sub = lane / CH;
      const int ch = lane % CH;
#pragma unroll 4
      for (int it = 0; it < 32 / RPI; ++it) {
        const int r = it * RPI + sub;
        const int qrow = q0 + wg * ATT_BM + quad * 32 + r;
        uint4 w;
        const uint32_t src = stage + (uint32_t)r * (HD * 2) + (((uint32_t)ch ^ (uint32_t)(r & (CH - 1) & 7)) * 16);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(src));
        if (qrow < p.nq) {
          const int64_t owner = (int64_t)qrow / p.rows_per_owner;
          const int64_t lrow = (int64_t)qrow - owner * p.rows_per_owner;
          __nv_bfloat16* dst =
              p.o[owner] + (int64_t)batch * p.o_batch_stride + lrow * p.o_row_stride + (int64_t)head * HD + ch * 8;
          *reinterpret_cast<uint4*>(dst) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == A64_SOFTMAX_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int EMU, int EMU_B, bool STAGGER = true>
static int launch_attn64x4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                           int batch, cudaStream_t stream) {
  static uint64_t configured = 0;  // one bit per device: the opt-in is a per-device function attribute
  const uint64_t dev_bit = 1ull << (current_device() & 63);
  if (!(configured & dev_bit)) {
    FINO_CHECK_CUDA(cudaFuncSetAttribute(attn64x4_fwd_kernel<EMU, EMU_B, STAGGER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         A64_SMEM_BYTES));
    configured |= dev_bit;
  }
  const int n_tiles = p.q_tiles * p.heads * batch;
  const int n_part = (n_tiles - p.n_full) * p.splits;
  attn64x4_fwd_kernel<EMU, EMU_B, STAGGER><<<p.n_full + n_part, A64_THREADS, A64_SMEM_BYTES, stream>>>(tq, tk, tv, p);
  FINO_CHECK_CUDA(cudaGetLastError());
  if (n_part > 0) {
    const int rows_total = (n_tiles - p.n_full) * p.tile_rows;
    attn_combine_kernel<64><<<(rows_total + 7) / 8, 256, 0, stream>>>(p, rows_total);
    FINO_CHECK_CUDA(cudaGetLastError());
  }
  return FINO_OK;
}

// Tuning hook (fino_attention_set_variant): 0 = default. Variants differ only in scheduling / which pipe computes
// the exponentials; results agree to bf16 rounding.
static int g_attn_variant = 0;
void attention_set_variant(int v) { g_attn_variant = v; }

// KV split of the partial last wave (fino_attention_set_split): -1 = automatic (default), 0 = never, S >= 2 = split
// EVERY tile S ways (test hook: exercises the partial + combine path at any shape).
static int g_attn_split = -1;
void attention_set_split(int s) { g_attn_split = s; }

// Chooses (n_full, splits) for n_tiles tiles of T KV tiles each on `sms` SMs. A grid of n_tiles equal CTAs runs in
// ceil(n_tiles / sms) waves; when the last wave is only partly full (rem tiles), splitting each of its tiles
// S = floor(sms / rem) ways turns that wave into 1/S of a wave plus a small merge — at 8-way Ulysses (3 heads,
// 330 tiles) the kernel goes from 3 waves to 2.25.
void attention_plan(int n_tiles, int T, int sms, int mode, int* n_full, int* splits) {
  *n_full = n_tiles;
  *splits = 1;
  if (mode == 0) return;
  if (mode >= 2) {
    const int s = mode < T ? mode : T;
    if (s >= 2) {
      *n_full = 0;
      *splits = s;
    }
    return;
  }
  const int rem = n_tiles % sms;
  if (rem == 0) return;
  int s = sms / rem;
  if (s > 8) s = 8;
  if (s > T / 8) s = T / 8;  // keep at least 8 KV tiles per partial CTA
  if (s < 2) return;
  *n_full = n_tiles - rem;
  *splits = s;
}

template <int HD>
static int dispatch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                         int batch, cudaStream_t stream) {
  switch (g_attn_variant) {
    case 1: return launch_attn<HD, 0, false>(tq, tk, tv, p, batch, stream);  // all MUFU, P published once
    case 2: return launch_attn<HD, 0, true>(tq, tk, tv, p, batch, stream);   // all MUFU, split P
    case 3: return launch_attn<HD, 2, false>(tq, tk, tv, p, batch, stream);  // 2/8 emulated
    case 4: return launch_attn<HD, 2, true>(tq, tk, tv, p, batch, stream);
    case 5: return launch_attn<HD, 4, true>(tq, tk, tv, p, batch, stream);   // 4/8 emulated
    default: break;
  }
  if constexpr (HD == 64) {  // separate P columns + early S issue (variants 6..9; 0 = the measured best of them)
    switch (g_attn_variant) {
      case 6: return launch_attn<HD, 2, true, true, 2>(tq, tk, tv, p, batch, stream);   // 2/8 emulated
      case 7: return launch_attn<HD, 2, true, true, 4>(tq, tk, tv, p, batch, stream);   // 3/8
      case 8: return launch_attn<HD, 4, true, true, 4>(tq, tk, tv, p, batch, stream);   // 4/8
      case 9: return launch_attn<HD, 2, false, true, 4>(tq, tk, tv, p, batch, stream);  // 3/8, P published once
      case 10: return launch_attn<HD, 2, true, true, 2, true>(tq, tk, tv, p, batch, stream);   // 2/8 + turns
      case 11: return launch_attn<HD, 2, false, true, 2, true>(tq, tk, tv, p, batch, stream);  // 2/8 + turns, no split
      case 12: return launch_attn<HD, 0, false, true, 0, true>(tq, tk, tv, p, batch, stream);  // all MUFU + turns
      case 13: return launch_attn<HD, 2, false, true, 4, true>(tq, tk, tv, p, batch, stream);  // 3/8 + turns
      default: return launch_attn<HD, 2, true, true, 2, true>(tq, tk, tv, p, batch, stream);  // = variant 10
    }
  }
  return launch_attn<HD, 2, true>(tq, tk, tv, p, batch, stream);
}

// o_owners[num_owners]: output buffers; rows [g*rows_per_owner, (g+1)*rows_per_owner) go to owner g (see AttnParams).
int attention_fwd_owners(const void* q, const void* k, const void* v, void* const* o_owners, int num_owners,
                         int64_t rows_per_owner, int batch, int heads, int64_t nq, int64_t nk, int head_dim,
                         int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride, int64_t o_row_stride,
                         int64_t q_batch_stride, int64_t k_batch_stride, int64_t v_batch_stride,
                         int64_t o_batch_stride, float scale, cudaStream_t stream) {
  FINO_CHECK_ARG(q && k && v && o_owners, "attention: null pointer");
  FINO_CHECK_ARG(num_owners >= 1 && num_owners <= ATT_MAX_OWNERS, "attention: 1..%d output owners", ATT_MAX_OWNERS);
  FINO_CHECK_ARG(rows_per_owner > 0 && rows_per_owner * num_owners >= nq, "attention: owners do not cover nq rows");
  for (int g = 0; g < num_owners; ++g)
    FINO_CHECK_ARG(o_owners[g] != nullptr && (reinterpret_cast<uintptr_t>(o_owners[g]) & 15) == 0,
                   "attention: output pointer %d null or not 16-byte aligned", g);
  FINO_CHECK_ARG(head_dim == 128 || head_dim == 64, "attention: head_dim %d unsupported (64 or 128)", head_dim);
  FINO_CHECK_ARG(batch > 0 && heads > 0 && nq > 0 && nk > 0, "attention: non-positive shape");
  FINO_CHECK_ARG(q_row_stride % 8 == 0 && k_row_stride % 8 == 0 && v_row_stride % 8 == 0 && o_row_stride % 8 == 0,
                 "attention: row strides must be multiples of 8 elements");
  FINO_CHECK_ARG(q_batch_stride % 8 == 0 && k_batch_stride % 8 == 0 && v_batch_stride % 8 == 0 &&
                     o_batch_stride % 8 == 0,
                 "attention: batch strides must be multiples of 8 elements");

  CUtensorMap tq, tk, tv;
  const uint64_t inner = (uint64_t)heads * head_dim;
  // head_dim 64 runs the four-tile kernel (512 query rows per CTA, 64-key steps) unless a variant of the two-tile
  // kernel is selected through the tuning hook
  const bool x4 = head_dim == 64 && (g_attn_variant == 0 || g_attn_variant >= 14);
  const int kv_rows = x4 ? A64_BN : ATT_BN;
  const int tile_rows = x4 ? A64_NT * ATT_BM : 2 * ATT_BM;
  auto enc = [&](CUtensorMap* tm, const void* base, int64_t n, int64_t rs, int64_t bs, uint32_t box_rows) {
    uint64_t dims[3] = {inner, (uint64_t)n, (uint64_t)batch};
    // a batch stride of 0 is not encodable; with batch == 1 any legal value works
    uint64_t strides[2] = {(uint64_t)rs * 2, (uint64_t)(batch > 1 ? bs : rs * n) * 2};
    uint32_t box[3] = {64, box_rows, 1};
    return encode_tmap_bf16(tm, base, 3, dims, strides, box);
  };
  int r;
  if ((r = enc(&tq, q, nq, q_row_stride, q_batch_stride, ATT_BM))) return r;
  if ((r = enc(&tk, k, nk, k_row_stride, k_batch_stride, (uint32_t)kv_rows))) return r;
  if ((r = enc(&tv, v, nk, v_row_stride, v_batch_stride, (uint32_t)kv_rows))) return r;

  AttnParams p;
  for (int g = 0; g < ATT_MAX_OWNERS; ++g)
    p.o[g] = reinterpret_cast<__nv_bfloat16*>(o_owners[g < num_owners ? g : 0]);
  p.rows_per_owner = rows_per_owner;
  p.o_row_stride = o_row_stride;
  p.o_batch_stride = o_batch_stride;
  p.nq = (int)nq;
  p.nk = (int)nk;
  p.heads = heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.num_kv_tiles = (int)((nk + kv_rows - 1) / kv_rows);
  p.q_tiles = (int)((nq + tile_rows - 1) / tile_rows);
  p.tile_rows = tile_rows;
  FINO_CHECK_ARG((int64_t)p.q_tiles * heads * batch < (int64_t)1 << 30, "attention: too many query tiles");
  const int n_tiles = p.q_tiles * heads * batch;
  attention_plan(n_tiles, p.num_kv_tiles, num_sms(), g_attn_split, &p.n_full, &p.splits);
  p.ws_o = nullptr;
  p.ws_ml = nullptr;
  if (p.splits > 1) {
    const size_t prow = (size_t)(n_tiles - p.n_full) * p.splits * tile_rows;
    void* ws = nullptr;
    if ((r = attn_workspace(prow * ((size_t)head_dim * 4 + 8), stream, &ws))) return r;
    p.ws_o = reinterpret_cast<float*>(ws);
    p.ws_ml = reinterpret_cast<float2*>(p.ws_o + prow * head_dim);
  }
  if (head_dim == 128) return dispatch_attn<128>(tq, tk, tv, p, batch, stream);
  if (x4) {
    switch (g_attn_variant) {
      case 14: return launch_attn64x4<2, 2, false>(tq, tk, tv, p, batch, stream);  // tile pairs in phase
      case 15: return launch_attn64x4<2, 4>(tq, tk, tv, p, batch, stream);  // 3/8 emulated
      default: return launch_attn64x4<2, 2>(tq, tk, tv, p, batch, stream);  // 2/8 emulated
    }
  }
  return dispatch_attn<64>(tq, tk, tv, p, batch, stream);
}

int attention_fwd(const void* q, const void* k, const void* v, void* o, int batch, int heads, int64_t nq, int64_t nk,
                  int head_dim, int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride, int64_t o_row_stride,
                  int64_t q_batch_stride, int64_t k_batch_stride, int64_t v_batch_stride, int64_t o_batch_stride,
                  float scale, cudaStream_t stream) {
  void* owners[1] = {o};
  return attention_fwd_owners(q, k, v, owners, 1, nq, batch, heads, nq, nk, head_dim, q_row_stride, k_row_stride,
                              v_row_stride, o_row_stride, q_batch_stride, k_batch_stride, v_batch_stride,
                              o_batch_stride, scale, stream);
}

}  // namespace fino
