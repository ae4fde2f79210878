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

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715x^3)))  (torch GELU(approximate="tanh"))
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(inner));
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ const float* gate_row_ptr(const GemmParams& p, int64_t row, bool row_ok) {
  if (p.epilogue == EPI_GATE_RESIDUAL && p.gate != nullptr && row_ok) {
    int64_t gi = p.row_index ? (int64_t)p.row_index[row] : (row / p.rows_per_group);
    return p.gate + gi * p.gate_row_stride;
  }
  return nullptr;
}

// One 32-column chunk of one accumulator row: bias, activation / gated residual, convert, 16-byte stores.
// `bias32`: pointer to the 32 bias values of this chunk (global or shared memory), or null.
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int64_t row, int col0, const float* gate_row,
                                               const uint32_t (&r)[32], const __nv_bfloat16* bias32) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  const int ncol = min(32, p.N - col0);
  if (bias32 != nullptr) {
    if (ncol == 32) {
      const uint4* bp = reinterpret_cast<const uint4*>(bias32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 b = bp[j];
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
      for (int j = 0; j < ncol; ++j) v[j] += __bfloat162float(bias32[j]);
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

}  // namespace fino
