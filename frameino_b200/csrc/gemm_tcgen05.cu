// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * W[N,K]^T + bias)
//
// Replaces the nn.Linear calls of the reference hot path (cuBLAS via torch):
//   to_q/to_k/to_v/to_out  reference architecture/transformer_wan.py:60-62,117
//   ffn (FeedForward)      reference architecture/transformer_wan.py:347
//   patch_embedding/proj_out reference architecture/transformer_wan.py:486,537
// with the elementwise tails fused into the epilogue (bias, GELU-tanh, SiLU, gate*y + residual;
// reference transformer_wan.py:336,341,348).
//
// Design (B200): one CTA per SM, 256 threads.
//   warp 0        TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles, 4-stage mbarrier ring)
//   warp 1        MMA issuer     (one elected thread, tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16)
//   warp 2        TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4..7    epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> 16B global stores)
// The accumulator is double buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
#include "common.cuh"
#include "ptx.cuh"

namespace fino {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int GEMM_THREADS = 256;

enum GemmEpilogue : int {
  EPI_NONE = 0,           // C = acc (+bias)
  EPI_GELU_TANH = 1,      // C = gelu_tanh(bf16(acc+bias))
  EPI_SILU = 2,           // C = silu(bf16(acc+bias))
  EPI_GATE_RESIDUAL = 3,  // C = residual + bf16(acc+bias) * gate[row_index[row]]   (gate optional => 1)
};
enum GemmFlags : int {
  GEMM_FLAG_ROUND_PRODUCT = 1,  // round gate*y to bf16 before the residual add (CogVideoX bf16 flow)
};

struct GemmParams {
  int64_t M;
  int N, K;
  const __nv_bfloat16* bias;  // [N] or null
  void* C;
  int64_t ldc;
  int out_fp32;
  int epilogue;
  int flags;
  const __nv_bfloat16* residual;
  int64_t ldr;
  const float* gate;  // fp32 [R, gate_row_stride], column = output column
  int64_t gate_row_stride;
  const int32_t* row_index;  // [M] or null => row / rows_per_group
  int64_t rows_per_group;
  int num_m_tiles, num_n_tiles;
};

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;
  static constexpr int kBBytes = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;  // 512 or 256 (power of two)
};

__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int& mt, int& nt) {
  // groups of 8 M-tiles; inside a group N is the slow axis so that the 8 CTAs sharing a W tile run together
  constexpr int GM = 8;
  int tiles_per_group = GM * num_n;
  int g = tile / tiles_per_group;
  int first_m = g * GM;
  int gm = min(GM, num_m - first_m);
  int r = tile - g * tiles_per_group;
  nt = r / gm;
  mt = first_m + (r - nt * gm);
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715x^3)))  (torch GELU(approximate="tanh"))
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(inner));
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <int BN>
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

  const int warp = threadIdx.x >> 5;
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
          mbar_wait(empty_bar(stage), phase ^ 1u, 100 + stage);
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
    if (lane == 0) {
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
            umma_ss(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(empty_bar(stage));  // smem slot is free once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit(tfull_bar(acc));  // accumulator ready for the epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may touch
    int it = 0;
    const __nv_bfloat16* __restrict__ bias = p.bias;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int mt, nt;
      tile_coords(tile, p.num_m_tiles, p.num_n_tiles, mt, nt);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tfull_bar(acc), acc_phase, 400 + acc);
      tc_fence_after();
      const int64_t row = (int64_t)mt * GEMM_BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const float* gate_row = nullptr;
      if (p.epilogue == EPI_GATE_RESIDUAL && p.gate != nullptr && row_ok) {
        int64_t gi = p.row_index ? (int64_t)p.row_index[row] : (row / p.rows_per_group);
        gate_row = p.gate + gi * p.gate_row_stride;
      }
      const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = nt * BN + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr_row + c * 32, r);
        tmem_wait_ld();
        if (!row_ok) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const int ncol = min(32, p.N - col0);
        if (bias != nullptr) {
          if (ncol == 32) {
            const uint4* bp = reinterpret_cast<const uint4*>(bias + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 b = __ldg(bp + j);
              v[8 * j + 0] += bf16_lo_to_f32(b.x);
              v[8 * j + 1] += bf16_hi_to_f32(b.x);
              v[8 * j + 2] += bf16_lo_to_f32(b.y);
              v[8 * j + 3] += bf16_hi_to_f32(b.y);
              v[8 * j + 4] += bf16_lo_to_f32(b.z);
              v[8 * j + 5] += bf16_hi_to_f32(b.z);
              v[8 * j + 6] += bf16_lo_to_f32(b.w);
              v[8 * j + 7] += bf16_hi_to_f32(b.w);
            }
          } else {
            for (int j = 0; j < ncol; ++j) v[j] += __bfloat162float(bias[col0 + j]);
          }
        }
        if (p.epilogue == EPI_GELU_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(round_bf16(v[j]));
        } else if (p.epilogue == EPI_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu_f(round_bf16(v[j]));
        } else if (p.epilogue == EPI_GATE_RESIDUAL) {
          const __nv_bfloat16* rp = p.residual + row * p.ldr + col0;
          if (ncol == 32) {
            float g[32];
            if (gate_row != nullptr) {
              const float4* gp = reinterpret_cast<const float4*>(gate_row + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 t = __ldg(gp + j);
                g[4 * j + 0] = t.x;
                g[4 * j + 1] = t.y;
                g[4 * j + 2] = t.z;
                g[4 * j + 3] = t.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) g[j] = 1.0f;
            }
            const uint4* rp4 = reinterpret_cast<const uint4*>(rp);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 x = __ldg(rp4 + j);
              float xr[8] = {bf16_lo_to_f32(x.x), bf16_hi_to_f32(x.x), bf16_lo_to_f32(x.y), bf16_hi_to_f32(x.y),
                             bf16_lo_to_f32(x.z), bf16_hi_to_f32(x.z), bf16_lo_to_f32(x.w), bf16_hi_to_f32(x.w)};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float y = round_bf16(v[8 * j + e]) * g[8 * j + e];
                if (p.flags & GEMM_FLAG_ROUND_PRODUCT) y = round_bf16(y);
                v[8 * j + e] = xr[e] + y;
              }
            }
          } else {
            for (int j = 0; j < ncol; ++j) {
              float g = gate_row ? gate_row[col0 + j] : 1.0f;
              float y = round_bf16(v[j]) * g;
              if (p.flags & GEMM_FLAG_ROUND_PRODUCT) y = round_bf16(y);
              v[j] = __bfloat162float(rp[j]) + y;
            }
          }
        }
        if (p.out_fp32) {
          float* cp = reinterpret_cast<float*>(p.C) + row * p.ldc + col0;
          if (ncol == 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(cp)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            for (int j = 0; j < ncol; ++j) cp[j] = v[j];
          }
        } else {
          __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col0;
          if (ncol == 32) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
              o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              reinterpret_cast<uint4*>(cp)[j] = o;
            }
          } else {
            for (int j = 0; j < ncol; ++j) cp[j] = __float2bfloat16_rn(v[j]);
          }
        }
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

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    FINO_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
    configured = true;
  }
  int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_bf16_kernel<BN><<<grid, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

int gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* c, int64_t ldc,
              int64_t m, int n, int k, int epilogue, int out_fp32, int flags, const void* residual, int64_t ldr,
              const float* gate, int64_t gate_row_stride, const int32_t* row_index, int64_t rows_per_group,
              cudaStream_t stream) {
  FINO_CHECK_ARG(a && w && c, "gemm: null operand pointer");
  FINO_CHECK_ARG(m > 0 && n > 0 && k > 0, "gemm: non-positive shape m=%lld n=%d k=%d", (long long)m, n, k);
  FINO_CHECK_ARG(k % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: k/lda/ldw must be multiples of 8 (16-byte rows)");
  FINO_CHECK_ARG(ldc % 8 == 0, "gemm: ldc must be a multiple of 8");
  FINO_CHECK_ARG(n % 8 == 0, "gemm: n must be a multiple of 8");
  FINO_CHECK_ARG(epilogue >= EPI_NONE && epilogue <= EPI_GATE_RESIDUAL, "gemm: unknown epilogue %d", epilogue);
  if (epilogue == EPI_GATE_RESIDUAL) {
    FINO_CHECK_ARG(residual != nullptr && ldr % 8 == 0, "gemm: gate-residual epilogue needs a 16B-aligned residual");
    FINO_CHECK_ARG(gate == nullptr || (gate_row_stride % 4 == 0), "gemm: gate row stride must be a multiple of 4");
    FINO_CHECK_ARG(gate == nullptr || row_index != nullptr || rows_per_group > 0,
                   "gemm: gate needs row_index or rows_per_group");
  }
  const int BN = (n > 128) ? 256 : 128;
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
  p.num_m_tiles = (int)((m + GEMM_BM - 1) / GEMM_BM);
  p.num_n_tiles = (n + BN - 1) / BN;

  CUtensorMap ta, tb;
  {
    uint64_t dims[2] = {(uint64_t)k, (uint64_t)m};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int r = encode_tmap_bf16(&ta, a, 2, dims, strides, box);
    if (r) return r;
  }
  {
    uint64_t dims[2] = {(uint64_t)k, (uint64_t)n};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {GEMM_BK, (uint32_t)BN};
    int r = encode_tmap_bf16(&tb, w, 2, dims, strides, box);
    if (r) return r;
  }
  if (BN == 256) return launch_gemm<256>(ta, tb, p, stream);
  return launch_gemm<128>(ta, tb, p, stream);
}

}  // namespace fino
