// HBM-bound row kernels of the denoise step: one warp owns one token row, the row lives in registers,
// reductions are warp shuffles, every global access is a 128-bit vector.
//
//   ln_modulate      FP32LayerNorm (+affine) + AdaLN modulate   reference transformer_wan.py:334,339,344-346,536
//                    CogVideoXLayerNormZero / AdaLayerNorm body  reference cogvideox_transformer_3d.py:134,150,541
//   gate_residual    x + y*gate                                  reference transformer_wan.py:336,341,348
//   qk_norm_rope     RMSNorm across heads + 3-D RoPE (Wan)       reference transformer_wan.py:64-90
//                    per-head LayerNorm + RoPE on video tokens   reference attention_processor.py:2848-2860
#include "common.cuh"
#include "ptx.cuh"

namespace fino {

constexpr int ROW_WARPS = 8;  // rows per CTA

enum LnFlags : int {
  LN_FLAG_BF16_STEPS = 1,  // round after LN, after (1+scale), after the product and after the add (bf16 module flow)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo_to_f32(u.x);
  f[1] = bf16_hi_to_f32(u.x);
  f[2] = bf16_lo_to_f32(u.y);
  f[3] = bf16_hi_to_f32(u.y);
  f[4] = bf16_lo_to_f32(u.z);
  f[5] = bf16_hi_to_f32(u.z);
  f[6] = bf16_lo_to_f32(u.w);
  f[7] = bf16_hi_to_f32(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void ld8f(const float* p, float* f) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (+affine) (+modulate)
// ------------------------------------------------------------------------------------------------
template <int CPL>  // 16-byte chunks per lane
__global__ void __launch_bounds__(ROW_WARPS * 32)
ln_modulate_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t rows, int dim,
                   int64_t x_stride, int64_t out_stride, float eps, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ shift, const float* __restrict__ scale,
                   int64_t mod_row_stride, const int32_t* __restrict__ row_index, int64_t rows_per_group, int flags) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nchunks = dim >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * x_stride);
  float v[CPL][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = i * 32 + lane;
    if (c < nchunks) {
      uint4 u = ld_stream(xr + c);
      unpack8(u, v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[i][e];
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
    }
  }
  const float mean = warp_sum(sum) / (float)dim;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = i * 32 + lane;
    if (c < nchunks) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float d = v[i][e] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)dim + eps);

  const float* sh = nullptr;
  const float* sc = nullptr;
  if (shift != nullptr) {
    const int64_t g = row_index ? (int64_t)row_index[row] : row / rows_per_group;
    sh = shift + g * mod_row_stride;
    sc = scale + g * mod_row_stride;
  }
  const bool steps = (flags & LN_FLAG_BF16_STEPS) != 0;
  uint4* orow = reinterpret_cast<uint4*>(out + row * out_stride);
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = i * 32 + lane;
    if (c < nchunks) {
      float y[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = (v[i][e] - mean) * rstd;
      if (gamma != nullptr) {
        float g[8], b[8];
        ld8f(gamma + c * 8, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] *= g[e];
        if (beta != nullptr) {
          ld8f(beta + c * 8, b);
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] += b[e];
        }
      }
      if (sh != nullptr) {
        float s8[8], h8[8];
        ld8f(sc + c * 8, s8);
        ld8f(sh + c * 8, h8);
        if (steps) {
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = rbf(rbf(y[e]) * rbf(1.0f + s8[e])) + h8[e];
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = y[e] * (1.0f + s8[e]) + h8[e];
        }
      }
      orow[c] = pack8(y);
    }
  }
}

// Block-per-row variant for wide rows: thread c owns 16-byte chunk c of every row the CTA processes, so the LayerNorm
// affine and the AdaLN shift / scale values for its 8 columns stay in REGISTERS across rows and are re-read only when
// the modulation row changes (with the warp-per-row kernel every row pulled 24 KB of fp32 modulation through L1 for
// 12 KB of activation traffic, and L1 bandwidth, not HBM, set the pace: profiles/r01_rows_ncu.txt).
constexpr int LN_ROWS_PER_BLOCK = 16;
static bool g_ln_block_kernel = false;  // measured 2.4 TB/s vs 3.6 TB/s for the warp-per-row kernel (r01)
static bool g_qk_block_kernel = true;
void rows_set_variant(int ln_block, int qk_block) {
  g_ln_block_kernel = ln_block != 0;
  g_qk_block_kernel = qk_block != 0;
}

__global__ void __launch_bounds__(1024)
ln_modulate_block_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t rows, int dim,
                         int64_t x_stride, int64_t out_stride, float eps, const float* __restrict__ gamma,
                         const float* __restrict__ beta, const float* __restrict__ shift,
                         const float* __restrict__ scale, int64_t mod_row_stride,
                         const int32_t* __restrict__ row_index, int64_t rows_per_group, int flags) {
  __shared__ float red[2][32];
  const int c = threadIdx.x;
  const int lane = c & 31, warp = c >> 5;
  const int nwarps = blockDim.x >> 5;
  const int nchunks = dim >> 3;
  const bool active = c < nchunks;
  const bool steps = (flags & LN_FLAG_BF16_STEPS) != 0;
  const float inv_dim = 1.0f / (float)dim;
  float g8[8], b8[8], sc8[8], sh8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    g8[e] = 1.f;
    b8[e] = 0.f;
    sc8[e] = 0.f;
    sh8[e] = 0.f;
  }
  if (active && gamma != nullptr) ld8f(gamma + c * 8, g8);
  if (active && beta != nullptr) ld8f(beta + c * 8, b8);
  int64_t cached_g = -1;
  const int64_t r0 = (int64_t)blockIdx.x * LN_ROWS_PER_BLOCK;
  const int64_t r1 = min(rows, r0 + (int64_t)LN_ROWS_PER_BLOCK);
  uint4 cur = make_uint4(0, 0, 0, 0);
  if (active && r0 < r1) cur = ld_stream(reinterpret_cast<const uint4*>(x + r0 * x_stride) + c);
  for (int64_t row = r0; row < r1; ++row) {
    float v[8];
    unpack8(cur, v);
    if (active && row + 1 < r1) cur = ld_stream(reinterpret_cast<const uint4*>(x + (row + 1) * x_stride) + c);  // prefetch
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[e];
    s = warp_sum(s);
    if (lane == 0) red[0][warp] = s;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < nwarps; ++w) tot += red[0][w];
    const float mean = tot * inv_dim;
    float sq = 0.f;
    if (active) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[e] - mean;
        sq += d * d;
      }
    }
    sq = warp_sum(sq);
    if (lane == 0) red[1][warp] = sq;
    __syncthreads();
    float tsq = 0.f;
    for (int w = 0; w < nwarps; ++w) tsq += red[1][w];
    const float rstd = rsqrtf(tsq * inv_dim + eps);
    if (shift != nullptr) {
      const int64_t g = row_index ? (int64_t)row_index[row] : row / rows_per_group;
      if (g != cached_g) {  // block-uniform
        cached_g = g;
        if (active) {
          ld8f(scale + g * mod_row_stride + c * 8, sc8);
          ld8f(shift + g * mod_row_stride + c * 8, sh8);
        }
      }
    }
    if (active) {
      float y[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = (v[e] - mean) * rstd;
      if (gamma != nullptr) {
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = y[e] * g8[e] + b8[e];
      }
      if (shift != nullptr) {
        if (steps) {
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = rbf(rbf(y[e]) * rbf(1.0f + sc8[e])) + sh8[e];
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = y[e] * (1.0f + sc8[e]) + sh8[e];
        }
      }
      reinterpret_cast<uint4*>(out + row * out_stride)[c] = pack8(y);
    }
  }
}

int ln_modulate_tma(const void* x, void* out, int64_t rows, int dim, int64_t x_stride, int64_t out_stride, float eps,
                    const float* gamma, const float* beta, const float* shift, const float* scale,
                    int64_t mod_row_stride, const int32_t* row_index, int64_t rows_per_group, int flags,
                    cudaStream_t stream);
// TMA-staged persistent row kernels (rows_tma.cu). Off by default: measured slower than the warp-per-row kernels
// (in-step LN 260 us vs 120 us, q/k 331 us vs 267 us at 28160 x 3072, r01) - the block-wide reductions put three
// barrier phases on every ring stage. Kept behind fino_rows_set_tma for the next iteration (per-warp rings).
static bool g_rows_tma = false;
void rows_set_tma(int on) { g_rows_tma = on != 0; }

int ln_modulate(const void* x, void* out, int64_t rows, int dim, int64_t x_stride, int64_t out_stride, float eps,
                const float* gamma, const float* beta, const float* shift, const float* scale, int64_t mod_row_stride,
                const int32_t* row_index, int64_t rows_per_group, int flags, cudaStream_t stream) {
  FINO_CHECK_ARG(x && out, "ln_modulate: null pointer");
  FINO_CHECK_ARG(rows > 0 && dim > 0 && dim % 8 == 0, "ln_modulate: dim %d must be a positive multiple of 8", dim);
  FINO_CHECK_ARG(x_stride % 8 == 0 && out_stride % 8 == 0, "ln_modulate: row strides must be multiples of 8");
  FINO_CHECK_ARG((shift == nullptr) == (scale == nullptr), "ln_modulate: shift and scale go together");
  FINO_CHECK_ARG(shift == nullptr || mod_row_stride % 4 == 0, "ln_modulate: modulation row stride must be 16B aligned");
  FINO_CHECK_ARG(shift == nullptr || row_index != nullptr || rows_per_group > 0,
                 "ln_modulate: need row_index or rows_per_group");
  FINO_CHECK_ARG(dim <= 32 * 8 * 32, "ln_modulate: dim %d too large (max 8192)", dim);
  if (g_rows_tma && dim >= 1024 && dim <= 4096 && rows >= 256 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
    return ln_modulate_tma(x, out, rows, dim, x_stride, out_stride, eps, gamma, beta, shift, scale, mod_row_stride,
                           row_index, rows_per_group, flags, stream);
  const int cpl = (dim / 8 + 31) / 32;
  dim3 grid((unsigned)((rows + ROW_WARPS - 1) / ROW_WARPS));
  if (rows_per_group <= 0) rows_per_group = (int64_t)1 << 62;
  if (g_ln_block_kernel && dim >= 1024 && rows >= 64) {  // experimental block-per-row kernel (slower today)
    const int threads = ((dim / 8 + 31) / 32) * 32;
    dim3 bgrid((unsigned)((rows + LN_ROWS_PER_BLOCK - 1) / LN_ROWS_PER_BLOCK));
    ln_modulate_block_kernel<<<bgrid, threads, 0, stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)out, rows, dim, x_stride, out_stride, eps, gamma, beta, shift, scale,
        mod_row_stride, row_index, rows_per_group, flags);
    FINO_CHECK_CUDA(cudaGetLastError());
    return FINO_OK;
  }
#define LAUNCH_LN(C)                                                                                                 \
  ln_modulate_kernel<C><<<grid, ROW_WARPS * 32, 0, stream>>>(                                                        \
      (const __nv_bfloat16*)x, (__nv_bfloat16*)out, rows, dim, x_stride, out_stride, eps, gamma, beta, shift, scale, \
      mod_row_stride, row_index, rows_per_group, flags)
  if (cpl <= 1) LAUNCH_LN(1);
  else if (cpl <= 2) LAUNCH_LN(2);
  else if (cpl <= 4) LAUNCH_LN(4);
  else if (cpl <= 8) LAUNCH_LN(8);
  else if (cpl <= 12) LAUNCH_LN(12);
  else if (cpl <= 16) LAUNCH_LN(16);
  else LAUNCH_LN(32);
#undef LAUNCH_LN
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// out = x + y * gate   (gate fp32 per modulation row, optional)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gate_residual_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                     __nv_bfloat16* __restrict__ out, int64_t rows, int dim, int64_t x_stride, int64_t y_stride,
                     int64_t out_stride, const float* __restrict__ gate, int64_t mod_row_stride,
                     const int32_t* __restrict__ row_index, int64_t rows_per_group, int round_product) {
  const int nchunks = dim >> 3;
  const int64_t total = rows * nchunks;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / nchunks;
    const int c = (int)(idx - row * nchunks);
    float xv[8], yv[8];
    unpack8(ld_stream(reinterpret_cast<const uint4*>(x + row * x_stride) + c), xv);
    unpack8(ld_stream(reinterpret_cast<const uint4*>(y + row * y_stride) + c), yv);
    if (gate != nullptr) {
      const int64_t g = row_index ? (int64_t)row_index[row] : row / rows_per_group;
      float gv[8];
      ld8f(gate + g * mod_row_stride + c * 8, gv);
#pragma unroll
      for (int e = 0; e < 8; ++e) yv[e] *= gv[e];
      if (round_product) {
#pragma unroll
        for (int e = 0; e < 8; ++e) yv[e] = rbf(yv[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) xv[e] += yv[e];
    reinterpret_cast<uint4*>(out + row * out_stride)[c] = pack8(xv);
  }
}

int gate_residual(const void* x, const void* y, void* out, int64_t rows, int dim, int64_t x_stride, int64_t y_stride,
                  int64_t out_stride, const float* gate, int64_t mod_row_stride, const int32_t* row_index,
                  int64_t rows_per_group, int round_product, cudaStream_t stream) {
  FINO_CHECK_ARG(x && y && out, "gate_residual: null pointer");
  FINO_CHECK_ARG(rows > 0 && dim > 0 && dim % 8 == 0, "gate_residual: dim must be a positive multiple of 8");
  FINO_CHECK_ARG(x_stride % 8 == 0 && y_stride % 8 == 0 && out_stride % 8 == 0, "gate_residual: strides % 8");
  FINO_CHECK_ARG(gate == nullptr || row_index != nullptr || rows_per_group > 0,
                 "gate_residual: need row_index or rows_per_group");
  if (rows_per_group <= 0) rows_per_group = (int64_t)1 << 62;
  const int64_t total = rows * (dim / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  gate_residual_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, (__nv_bfloat16*)out, rows, dim, x_stride, y_stride, out_stride,
      gate, mod_row_stride, row_index, rows_per_group, round_product);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// q/k normalisation + rotary embedding, in place
// ------------------------------------------------------------------------------------------------
enum QkNormMode : int { QK_RMS_ACROSS_HEADS = 0, QK_LAYERNORM_PER_HEAD = 1 };
enum RopeMode : int { ROPE_NONE = 0, ROPE_WAN = 1, ROPE_COGVIDEOX = 2 };

struct QkTensor {
  __nv_bfloat16* ptr;
  int64_t rows;
  int64_t row_stride;
  const __nv_bfloat16* weight;
  const __nv_bfloat16* bias;
  int rope;  // apply rope to this tensor
};

struct QkParams {
  QkTensor t[2];
  int heads, head_dim;
  int norm_mode;
  float eps;
  int rope_mode;
  const float* cos;  // [rope_rows, head_dim] fp32 (full width, as produced by the reference rope modules)
  const float* sin;
  int64_t seq_len;    // rows per batch element
  int64_t rope_skip;  // rows [0, rope_skip) of every sequence are not rotated (CogVideoX text tokens)
  int64_t blocks0;    // CTAs assigned to tensor 0
};

template <int CPL>
__global__ void __launch_bounds__(ROW_WARPS * 32) qk_norm_rope_kernel(const QkParams p) {
  const int lane = threadIdx.x & 31;
  const int which = (int64_t)blockIdx.x >= p.blocks0 ? 1 : 0;
  const QkTensor& t = p.t[which];
  const int64_t blk = which ? (int64_t)blockIdx.x - p.blocks0 : (int64_t)blockIdx.x;
  const int64_t row = blk * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= t.rows) return;
  const int dim = p.heads * p.head_dim;
  const int nchunks = dim >> 3;
  uint4* xr = reinterpret_cast<uint4*>(t.ptr + row * t.row_stride);
  float v[CPL][8];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = i * 32 + lane;
    if (c < nchunks) {
      unpack8(xr[c], v[i]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
    }
  }

  if (p.norm_mode == QK_RMS_ACROSS_HEADS) {
    // RMSNorm(D): fp32 mean square -> x*rsqrt -> cast to weight dtype (bf16) -> * weight (bf16 product)
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) sq += v[i][e] * v[i][e];
    const float rstd = rsqrtf(warp_sum(sq) / (float)dim + p.eps);
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = i * 32 + lane;
      if (c < nchunks) {
        float w[8];
        if (t.weight != nullptr) {
          unpack8(__ldg(reinterpret_cast<const uint4*>(t.weight) + c), w);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[i][e] = rbf(rbf(v[i][e] * rstd) * w[e]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[i][e] = rbf(v[i][e] * rstd);
        }
      }
    }
  } else {
    // LayerNorm(head_dim) per head; a head occupies head_dim/8 consecutive lanes (power of two <= 32)
    const int lanes_per_head = p.head_dim >> 3;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = i * 32 + lane;
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
      for (int o = lanes_per_head >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / (float)p.head_dim;
      float sq = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float d = v[i][e] - mean;
        sq += d * d;
      }
      for (int o = lanes_per_head >> 1; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq / (float)p.head_dim + p.eps);
      if (c < nchunks) {
        const int hc = c % lanes_per_head;  // chunk inside the head
        float w[8], b[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          w[e] = 1.f;
          b[e] = 0.f;
        }
        if (t.weight != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(t.weight) + hc), w);
        if (t.bias != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(t.bias) + hc), b);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i][e] = rbf((v[i][e] - mean) * rstd * w[e] + b[e]);
      }
    }
  }

  const int64_t s_in_seq = row % p.seq_len;
  const bool do_rope = t.rope && p.rope_mode != ROPE_NONE && s_in_seq >= p.rope_skip;
  const float* cos_row = do_rope ? p.cos + (s_in_seq - p.rope_skip) * p.head_dim : nullptr;
  const float* sin_row = do_rope ? p.sin + (s_in_seq - p.rope_skip) * p.head_dim : nullptr;
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = i * 32 + lane;
    if (c < nchunks) {
      if (do_rope) {
        const int off = (c * 8) % p.head_dim;
        float cs[8], sn[8];
        ld8f(cos_row + off, cs);
        ld8f(sin_row + off, sn);
        float o[8];
        if (p.rope_mode == ROPE_WAN) {
          // cos = freqs_cos[..., 0::2], sin = freqs_sin[..., 1::2]   (transformer_wan.py:83-87)
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const float x1 = v[i][e], x2 = v[i][e + 1];
            o[e] = __fsub_rn(__fmul_rn(x1, cs[e]), __fmul_rn(x2, sn[e + 1]));
            o[e + 1] = __fadd_rn(__fmul_rn(x1, sn[e + 1]), __fmul_rn(x2, cs[e]));
          }
        } else {
          // x*cos + rotate(x)*sin with rotate = (-x_odd, x_even)   (embeddings.py:1247-1256)
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const float xe = v[i][e], xo = v[i][e + 1];
            o[e] = __fadd_rn(__fmul_rn(xe, cs[e]), __fmul_rn(-xo, sn[e]));
            o[e + 1] = __fadd_rn(__fmul_rn(xo, cs[e + 1]), __fmul_rn(xe, sn[e + 1]));
          }
        }
        xr[c] = pack8(o);
      } else {
        xr[c] = pack8(v[i]);
      }
    }
  }
}

// Block-per-token variant of the RMS-across-heads path (Wan self-attention): thread c owns 16-byte chunk c of the q
// row AND of the k row of each token the CTA processes. The RMSNorm weights of its 8 columns live in registers for
// the whole CTA, the cos/sin values of the token are fetched once and used for both q and k, and the two
// sum-of-squares reductions share one __syncthreads per token.
constexpr int QK_TOKENS_PER_BLOCK = 8;

__global__ void __launch_bounds__(512) qk_rms_rope_block_kernel(const QkParams p) {
  __shared__ float red[2][2][32];
  const int c = threadIdx.x;
  const int lane = c & 31, warp = c >> 5;
  const int nwarps = blockDim.x >> 5;
  const int dim = p.heads * p.head_dim;
  const int nchunks = dim >> 3;
  const bool active = c < nchunks;
  const float inv_dim = 1.0f / (float)dim;
  const QkTensor& tq = p.t[0];
  const QkTensor& tk = p.t[1];
  float wq[8], wk[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) wq[e] = wk[e] = 1.f;
  if (active && tq.weight != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(tq.weight) + c), wq);
  if (active && tk.weight != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(tk.weight) + c), wk);
  const int off = (c * 8) % p.head_dim;
  const int64_t t0 = (int64_t)blockIdx.x * QK_TOKENS_PER_BLOCK;
  const int64_t t1 = min(tq.rows, t0 + (int64_t)QK_TOKENS_PER_BLOCK);
  uint4 uq = make_uint4(0, 0, 0, 0), uk = make_uint4(0, 0, 0, 0);
  if (active && t0 < t1) {
    uq = reinterpret_cast<const uint4*>(tq.ptr + t0 * tq.row_stride)[c];
    uk = reinterpret_cast<const uint4*>(tk.ptr + t0 * tk.row_stride)[c];
  }
  int it = 0;
  for (int64_t tok = t0; tok < t1; ++tok, ++it) {
    float q[8], k[8];
    unpack8(uq, q);
    unpack8(uk, k);
    uint4* qrow = reinterpret_cast<uint4*>(tq.ptr + tok * tq.row_stride);
    uint4* krow = reinterpret_cast<uint4*>(tk.ptr + tok * tk.row_stride);
    if (active && tok + 1 < t1) {  // prefetch the next token's rows
      uq = reinterpret_cast<const uint4*>(tq.ptr + (tok + 1) * tq.row_stride)[c];
      uk = reinterpret_cast<const uint4*>(tk.ptr + (tok + 1) * tk.row_stride)[c];
    }
    const int64_t s_in_seq = tok % p.seq_len;
    const bool do_rope = p.rope_mode != ROPE_NONE && s_in_seq >= p.rope_skip;
    float cs[8], sn[8];
    if (do_rope && active) {
      ld8f(p.cos + (s_in_seq - p.rope_skip) * p.head_dim + off, cs);
      ld8f(p.sin + (s_in_seq - p.rope_skip) * p.head_dim + off, sn);
    }
    float sq = 0.f, sk = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sq += q[e] * q[e];
      sk += k[e] * k[e];
    }
    sq = warp_sum(sq);
    sk = warp_sum(sk);
    if (lane == 0) {
      red[it & 1][0][warp] = sq;
      red[it & 1][1][warp] = sk;
    }
    __syncthreads();
    float tq2 = 0.f, tk2 = 0.f;
    for (int w = 0; w < nwarps; ++w) {
      tq2 += red[it & 1][0][w];
      tk2 += red[it & 1][1][w];
    }
    const float rq = rsqrtf(tq2 * inv_dim + p.eps);
    const float rk = rsqrtf(tk2 * inv_dim + p.eps);
    if (active) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        q[e] = tq.weight ? rbf(rbf(q[e] * rq) * wq[e]) : rbf(q[e] * rq);
        k[e] = tk.weight ? rbf(rbf(k[e] * rk) * wk[e]) : rbf(k[e] * rk);
      }
      if (do_rope) {
        float oq[8], ok[8];
        if (p.rope_mode == ROPE_WAN) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            oq[e] = __fsub_rn(__fmul_rn(q[e], cs[e]), __fmul_rn(q[e + 1], sn[e + 1]));
            oq[e + 1] = __fadd_rn(__fmul_rn(q[e], sn[e + 1]), __fmul_rn(q[e + 1], cs[e]));
            ok[e] = __fsub_rn(__fmul_rn(k[e], cs[e]), __fmul_rn(k[e + 1], sn[e + 1]));
            ok[e + 1] = __fadd_rn(__fmul_rn(k[e], sn[e + 1]), __fmul_rn(k[e + 1], cs[e]));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            oq[e] = __fadd_rn(__fmul_rn(q[e], cs[e]), __fmul_rn(-q[e + 1], sn[e]));
            oq[e + 1] = __fadd_rn(__fmul_rn(q[e + 1], cs[e + 1]), __fmul_rn(q[e], sn[e + 1]));
            ok[e] = __fadd_rn(__fmul_rn(k[e], cs[e]), __fmul_rn(-k[e + 1], sn[e]));
            ok[e + 1] = __fadd_rn(__fmul_rn(k[e + 1], cs[e + 1]), __fmul_rn(k[e], sn[e + 1]));
          }
        }
        qrow[c] = pack8(oq);
        krow[c] = pack8(ok);
      } else {
        qrow[c] = pack8(q);
        krow[c] = pack8(k);
      }
    }
  }
}

bool qk_tma_eligible(int64_t rows, int heads, int head_dim);
int qk_rms_rope_tma(void* q, int64_t q_stride, void* k, int64_t k_stride, int64_t rows, const void* wq, const void* wk,
                    int heads, int head_dim, float eps, const float* cos, const float* sin, int64_t seq_len,
                    cudaStream_t stream);

int qk_norm_rope(void* x0, int64_t rows0, int64_t stride0, const void* w0, const void* b0, int rope0, void* x1,
                 int64_t rows1, int64_t stride1, const void* w1, const void* b1, int rope1, int heads, int head_dim,
                 int norm_mode, float eps, int rope_mode, const float* cos, const float* sin, int64_t seq_len,
                 int64_t rope_skip, cudaStream_t stream) {
  FINO_CHECK_ARG(x0 != nullptr && rows0 > 0, "qk_norm_rope: first tensor missing");
  FINO_CHECK_ARG(heads > 0 && head_dim >= 8 && head_dim % 8 == 0, "qk_norm_rope: bad heads/head_dim");
  FINO_CHECK_ARG(stride0 % 8 == 0 && (x1 == nullptr || stride1 % 8 == 0), "qk_norm_rope: strides % 8");
  FINO_CHECK_ARG(norm_mode == QK_RMS_ACROSS_HEADS || norm_mode == QK_LAYERNORM_PER_HEAD, "qk_norm_rope: norm_mode");
  if (norm_mode == QK_LAYERNORM_PER_HEAD) {
    const int lph = head_dim / 8;
    FINO_CHECK_ARG(lph <= 32 && (lph & (lph - 1)) == 0, "qk_norm_rope: per-head LayerNorm needs head_dim in {8..256} pow2");
  }
  FINO_CHECK_ARG(rope_mode >= ROPE_NONE && rope_mode <= ROPE_COGVIDEOX, "qk_norm_rope: rope_mode");
  const bool any_rope = rope_mode != ROPE_NONE && (rope0 || (x1 && rope1));
  FINO_CHECK_ARG(!any_rope || (cos && sin && seq_len > 0), "qk_norm_rope: rope tables / seq_len missing");
  const int dim = heads * head_dim;
  FINO_CHECK_ARG(dim <= 8192, "qk_norm_rope: heads*head_dim too large");
  // Wan self-attention (q and k, same rows, both rotated or neither) and cross-attention q (alone, no RoPE):
  // TMA-staged persistent kernel (rows_tma.cu)
  if (g_rows_tma && norm_mode == QK_RMS_ACROSS_HEADS && qk_tma_eligible(rows0, heads, head_dim) && rope_skip == 0 &&
      b0 == nullptr && b1 == nullptr && (rope_mode == ROPE_NONE || rope_mode == ROPE_WAN) &&
      (x1 == nullptr ? !any_rope : (rows1 == rows0 && (!any_rope || (rope0 && rope1)))) &&
      ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(cos) |
        reinterpret_cast<uintptr_t>(sin)) & 15) == 0 && (!any_rope || head_dim % 4 == 0))
    return qk_rms_rope_tma(x0, stride0, x1, stride1, rows0, w0, w1, heads, head_dim, eps, any_rope ? cos : nullptr,
                           any_rope ? sin : nullptr, seq_len, stream);
  // cross-attention: q has N rows, k the 512 text rows, no RoPE -> the wide q through the TMA kernel, k on its own
  if (g_rows_tma && norm_mode == QK_RMS_ACROSS_HEADS && x1 != nullptr && rows1 != rows0 && !any_rope &&
      b0 == nullptr && b1 == nullptr && qk_tma_eligible(rows0, heads, head_dim) &&
      (reinterpret_cast<uintptr_t>(x0) & 15) == 0) {
    int r = qk_rms_rope_tma(x0, stride0, nullptr, 0, rows0, w0, nullptr, heads, head_dim, eps, nullptr, nullptr, 0, stream);
    if (r != FINO_OK) return r;
    return qk_norm_rope(x1, rows1, stride1, w1, nullptr, 0, nullptr, 0, 0, nullptr, nullptr, 0, heads, head_dim, norm_mode,
                        eps, ROPE_NONE, nullptr, nullptr, 0, 0, stream);
  }
  QkParams p;
  p.t[0] = {(__nv_bfloat16*)x0, rows0, stride0, (const __nv_bfloat16*)w0, (const __nv_bfloat16*)b0, rope0};
  p.t[1] = {(__nv_bfloat16*)x1, x1 ? rows1 : 0, stride1, (const __nv_bfloat16*)w1, (const __nv_bfloat16*)b1, rope1};
  p.heads = heads;
  p.head_dim = head_dim;
  p.norm_mode = norm_mode;
  p.eps = eps;
  p.rope_mode = any_rope ? rope_mode : ROPE_NONE;
  p.cos = cos;
  p.sin = sin;
  p.seq_len = seq_len > 0 ? seq_len : (int64_t)1 << 62;
  p.rope_skip = rope_skip;
  p.blocks0 = (rows0 + ROW_WARPS - 1) / ROW_WARPS;
  const int64_t blocks1 = x1 ? (rows1 + ROW_WARPS - 1) / ROW_WARPS : 0;
  if (g_qk_block_kernel && norm_mode == QK_RMS_ACROSS_HEADS && x1 != nullptr && rows0 == rows1 && rows0 >= 64 && dim >= 1024 && dim <= 4096 &&
      (p.rope_mode == ROPE_NONE || (rope0 && rope1))) {
    const int threads = ((dim / 8 + 31) / 32) * 32;
    dim3 bgrid((unsigned)((rows0 + QK_TOKENS_PER_BLOCK - 1) / QK_TOKENS_PER_BLOCK));
    qk_rms_rope_block_kernel<<<bgrid, threads, 0, stream>>>(p);
    FINO_CHECK_CUDA(cudaGetLastError());
    return FINO_OK;
  }
  const int cpl = (dim / 8 + 31) / 32;
  dim3 grid((unsigned)(p.blocks0 + blocks1));
#define LAUNCH_QK(C) qk_norm_rope_kernel<C><<<grid, ROW_WARPS * 32, 0, stream>>>(p)
  if (cpl <= 1) LAUNCH_QK(1);
  else if (cpl <= 2) LAUNCH_QK(2);
  else if (cpl <= 4) LAUNCH_QK(4);
  else if (cpl <= 8) LAUNCH_QK(8);
  else if (cpl <= 12) LAUNCH_QK(12);
  else if (cpl <= 16) LAUNCH_QK(16);
  else LAUNCH_QK(32);
#undef LAUNCH_QK
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

}  // namespace fino
