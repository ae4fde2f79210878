// Shared pieces of the tcgen05 GEMM kernels: parameters, tile rasterisation and the fused epilogue.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace fino {

constexpr int GEMM_BM = 128;  // accumulator rows per CTA (== TMEM lanes)
constexpr int GEMM_BK = 64;   // 64 bf16 = 128 B = one swizzle-128B row
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
  // Split-K of the partly filled last round of the persistent CTA-pair kernel (gemm2). Work units [0, num_full) are
  // whole tiles; unit num_full + r is K-slice r % splits of tile num_full + r / splits, whose raw fp32 accumulators
  // go to ws[r][256][256]; gemm_fixup_kernel adds the slices and applies the epilogue. splits == 1: no partial units.
  int num_full, splits;
  float* ws;
};

// groups of GM M-tiles; inside a group N is the slow axis so that the CTAs sharing a W tile run together
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int& mt, int& nt) {
  constexpr int GM = 8;
  int tiles_per_group = GM * num_n;
  int g = tile / tiles_per_group;
  int first_m = g * GM;
  int gm = min(GM, num_m - first_m);
  int r = tile - g * tiles_per_group;
  nt = r / gm;
  mt = first_m + (r - nt * gm);
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 0.5*x*(1+tanh(u)), u = sqrt(2/pi)*(x+0.044715x^3)  (torch GELU(approximate="tanh")), written as x*sigmoid(2u) =
// x / (1 + 2^(-2u*log2e)): two MUFU ops (ex2, rcp), absolute error ~1e-7 * |x| — far below the bf16 rounding that follows.
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (x + k1 * x * x * x);
  return x * rcp_approx(1.0f + ex2_approx(-2.885390081777927f * u));
}
__device__ __forceinline__ float silu_f(float x) { return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ const float* gate_row_ptr(const GemmParams& p, int64_t row, bool row_ok) {
  if (p.epilogue == EPI_GATE_RESIDUAL && p.gate != nullptr && row_ok) {
    int64_t gi = p.row_index ? (int64_t)p.row_index[row] : (row / p.rows_per_group);
    return p.gate + gi * p.gate_row_stride;
  }
  return nullptr;
}

// Residual and gate values of one 32-column chunk of one row, fetched one chunk ahead of their use
// (EPI_GATE_RESIDUAL only). The gate loads used to sit inside epilogue_chunk, right in front of the multiply that
// consumes them: ncu showed the epilogue warps parked on that long-scoreboard stall eight times per tile
// (profiles/r01_gemm_gate_res_ncu.txt), which made the gate-residual epilogue longer than a K = 3072 main loop.
struct ResidualChunk {
  uint4 v[4];
  float4 g[8];
};
template <int EPI>
__device__ __forceinline__ void load_residual_chunk(const GemmParams& p, int64_t row, int col0, bool row_ok,
                                                    const float* gate_row, ResidualChunk& rc) {
  if (EPI != EPI_GATE_RESIDUAL) return;
  const int ngrp = row_ok ? min(4, (p.N - col0) >> 3) : 0;
  const uint4* rp4 = reinterpret_cast<const uint4*>(p.residual + row * p.ldr + col0);
#pragma unroll
  for (int g = 0; g < 4; ++g)
    if (g < ngrp) rc.v[g] = __ldg(rp4 + g);
  if (gate_row != nullptr) {
    const float4* gp4 = reinterpret_cast<const float4*>(gate_row + col0);
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (g < ngrp) {
        rc.g[2 * g] = __ldg(gp4 + 2 * g);
        rc.g[2 * g + 1] = __ldg(gp4 + 2 * g + 1);
      }
  }
}

// Eight consecutive columns of one output row: bias, activation / gated residual, convert, one 16-byte store (two for
// fp32 output). v = accumulators; bias8 = the 8 bias values (global or shared memory) or null; g0/g1 = gate values
// (used when has_gate); x = the 8 residual values (EPI_GATE_RESIDUAL).
template <int EPI>
__device__ __forceinline__ void epilogue_group8(const GemmParams& p, int64_t row, int col, float (&v)[8],
                                                const __nv_bfloat16* bias8, bool has_gate, const float4& g0,
                                                const float4& g1, const uint4& x) {
  if (bias8 != nullptr) {
    const uint4 b = *reinterpret_cast<const uint4*>(bias8);
    v[0] += bf16_lo_to_f32(b.x);
    v[1] += bf16_hi_to_f32(b.x);
    v[2] += bf16_lo_to_f32(b.y);
    v[3] += bf16_hi_to_f32(b.y);
    v[4] += bf16_lo_to_f32(b.z);
    v[5] += bf16_hi_to_f32(b.z);
    v[6] += bf16_lo_to_f32(b.w);
    v[7] += bf16_hi_to_f32(b.w);
  }
  if (EPI == EPI_GELU_TANH) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = gelu_tanh_f(round_bf16(v[e]));
  } else if (EPI == EPI_SILU) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = silu_f(round_bf16(v[e]));
  } else if (EPI == EPI_GATE_RESIDUAL) {
    float gt[8];
    if (has_gate) {
      gt[0] = g0.x, gt[1] = g0.y, gt[2] = g0.z, gt[3] = g0.w;
      gt[4] = g1.x, gt[5] = g1.y, gt[6] = g1.z, gt[7] = g1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) gt[e] = 1.0f;
    }
    const float xr[8] = {bf16_lo_to_f32(x.x), bf16_hi_to_f32(x.x), bf16_lo_to_f32(x.y), bf16_hi_to_f32(x.y),
                         bf16_lo_to_f32(x.z), bf16_hi_to_f32(x.z), bf16_lo_to_f32(x.w), bf16_hi_to_f32(x.w)};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float y = round_bf16(v[e]) * gt[e];
      if (p.flags & GEMM_FLAG_ROUND_PRODUCT) y = round_bf16(y);
      v[e] = xr[e] + y;
    }
  }
  if (p.out_fp32) {
    float4* cp = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + row * p.ldc + col);
    cp[0] = make_float4(v[0], v[1], v[2], v[3]);
    cp[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.ldc + col) = o;
  }
}

// One 32-column chunk of one accumulator row. N is a multiple of 8 (checked by the launcher), so a ragged chunk is
// handled as whole 8-column groups: everything stays in registers with static indices. `bias32`: the 32 bias values
// of this chunk (global or shared memory), or null.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int64_t row, int col0, const float* gate_row,
                                               const uint32_t (&r)[32], const __nv_bfloat16* bias32,
                                               const ResidualChunk& rc) {
  const int ngrp = min(4, (p.N - col0) >> 3);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g < ngrp) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * g + e]);
      epilogue_group8<EPI>(p, row, col0 + 8 * g, v, bias32 ? bias32 + 8 * g : nullptr, gate_row != nullptr,
                           rc.g[2 * g], rc.g[2 * g + 1], rc.v[g]);
    }
  }
}

}  // namespace fino
