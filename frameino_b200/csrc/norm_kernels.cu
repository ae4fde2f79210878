// HBM-bound row kernels of the denoise step: one warp owns one token row, the row lives in registers,
// reductions are warp shuffles, every global access is a 128-bit vector.
//
//   ln_modulate      FP32LayerNorm (+affine) + AdaLN modulate   reference transformer_wan.py:334,339,344-346,536
//                    CogVideoXLayerNormZero / AdaLayerNorm body  reference cogvideox_transformer_3d.py:134,150,541
//   gate_residual    x + y*gate                                  reference transformer_wan.py:336,341,348
//   qk_norm_rope     RMSNorm across heads + 3-D RoPE (Wan)       reference transformer_wan.py:64-90
//                    per-head LayerNorm + RoPE on video tokens   reference attention_processor.py:2848-2860
#include <algorithm>
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

// Warp-per-row kernel, second layout (variant 3). ncu on the kernel above (profiles/r01_rows_ncu.txt): L1 63 % busy,
// 16 warps per SM. Two causes, both addressed here:
//   * a lane owning 8 consecutive columns reads its fp32 shift / scale values as two float4 32 bytes apart, so every
//     warp-wide LDG.128 of the modulation tables spans 1 KB and costs 8 L1 wavefronts for 512 useful bytes. Here a
//     lane owns 4 consecutive columns per group of 128: the bf16 row moves as 8-byte pieces (256 contiguous bytes per
//     instruction) and the fp32 tables as 16-byte pieces (512 contiguous bytes) — every wavefront is full;
//   * the row was held as 96 fp32 registers (128 registers per thread, 2 CTAs per SM). Here it stays packed (bf16x2,
//     2 registers per group) and is unpacked where it is used, which fits 3 CTAs = 24 warps per SM.
template <int GPL, bool AFFINE, bool MOD, bool STEPS>  // GPL 4-column groups per lane: dim == GPL * 128 exactly
__global__ void __launch_bounds__(ROW_WARPS * 32, (GPL <= 24 ? 2 : 1))
ln_modulate_kernel2(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t rows,
                    int64_t x_stride, int64_t out_stride, float eps, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ shift, const float* __restrict__ scale,
                    int64_t mod_row_stride, const int32_t* __restrict__ row_index, int64_t rows_per_group) {
  constexpr int dim = GPL * 128;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint2* xr = reinterpret_cast<const uint2*>(x + row * x_stride) + lane;
  uint2 v[GPL];
#pragma unroll
  for (int i = 0; i < GPL; ++i)
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v[i].x), "=r"(v[i].y) : "l"(xr + i * 32));
  int64_t grp = 0;
  if (MOD) grp = row_index ? (int64_t)__ldg(row_index + row) : row / rows_per_group;
  // all fp32 arithmetic below is issued as packed f32x2 operations (add/fma.rn.f32x2): half the instruction count of
  // the scalar form — the scalar kernel was issue/latency bound (its time scales with the SM clock, not with HBM)
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < GPL; ++i) {
    fadd2(s0, s1, s0, s1, bf16_lo_to_f32(v[i].x), bf16_hi_to_f32(v[i].x));
    fadd2(s2, s3, s2, s3, bf16_lo_to_f32(v[i].y), bf16_hi_to_f32(v[i].y));
  }
  const float mean = warp_sum((s0 + s1) + (s2 + s3)) * (1.0f / (float)dim);
  const float nmean = -mean;
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
  for (int i = 0; i < GPL; ++i) {
    float d0, d1, d2, d3;
    fadd2(d0, d1, bf16_lo_to_f32(v[i].x), bf16_hi_to_f32(v[i].x), nmean, nmean);
    fadd2(d2, d3, bf16_lo_to_f32(v[i].y), bf16_hi_to_f32(v[i].y), nmean, nmean);
    ffma2(q0, q1, d0, d1, d0, d1, q0, q1);
    ffma2(q2, q3, d2, d3, d2, d3, q2, q3);
  }
  const float rstd = rsqrtf(warp_sum((q0 + q1) + (q2 + q3)) * (1.0f / (float)dim) + eps);
  const float c0 = nmean * rstd;  // (x - mean) * rstd == fma(x, rstd, -mean * rstd) up to one fp32 rounding

  const float4* sh4 = MOD ? reinterpret_cast<const float4*>(shift + grp * mod_row_stride) + lane : nullptr;
  const float4* sc4 = MOD ? reinterpret_cast<const float4*>(scale + grp * mod_row_stride) + lane : nullptr;
  const float4* ga4 = reinterpret_cast<const float4*>(gamma) + lane;
  const float4* be4 = reinterpret_cast<const float4*>(beta) + lane;
  uint2* orow = reinterpret_cast<uint2*>(out + row * out_stride) + lane;
#pragma unroll
  for (int i = 0; i < GPL; ++i) {
    float y0, y1, y2, y3;
    ffma2(y0, y1, bf16_lo_to_f32(v[i].x), bf16_hi_to_f32(v[i].x), rstd, rstd, c0, c0);
    ffma2(y2, y3, bf16_lo_to_f32(v[i].y), bf16_hi_to_f32(v[i].y), rstd, rstd, c0, c0);
    if (AFFINE) {
      const float4 a = __ldg(ga4 + i * 32);
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (beta != nullptr) b = __ldg(be4 + i * 32);
      ffma2(y0, y1, y0, y1, a.x, a.y, b.x, b.y);
      ffma2(y2, y3, y2, y3, a.z, a.w, b.z, b.w);
    }
    if (MOD) {
      const float4 s4 = __ldg(sc4 + i * 32);
      const float4 h4 = __ldg(sh4 + i * 32);
      float t0, t1, t2, t3;
      fadd2(t0, t1, s4.x, s4.y, 1.0f, 1.0f);
      fadd2(t2, t3, s4.z, s4.w, 1.0f, 1.0f);
      if (STEPS) {
        y0 = rbf(rbf(y0) * rbf(t0)) + h4.x;
        y1 = rbf(rbf(y1) * rbf(t1)) + h4.y;
        y2 = rbf(rbf(y2) * rbf(t2)) + h4.z;
        y3 = rbf(rbf(y3) * rbf(t3)) + h4.w;
      } else {
        ffma2(y0, y1, y0, y1, t0, t1, h4.x, h4.y);
        ffma2(y2, y3, y2, y3, t2, t3, h4.z, h4.w);
      }
    }
    orow[i * 32] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
  }
}

template <int GPL>
static void launch_ln2(dim3 grid, cudaStream_t stream, bool aff, bool mod, bool steps, const void* x, void* out,
                       int64_t rows, int64_t x_stride, int64_t out_stride, float eps, const float* gamma,
                       const float* beta, const float* shift, const float* scale, int64_t mod_row_stride,
                       const int32_t* row_index, int64_t rows_per_group) {
#define LN2(A, M, S)                                                                                             \
  ln_modulate_kernel2<GPL, A, M, S><<<grid, ROW_WARPS * 32, 0, stream>>>(                                        \
      (const __nv_bfloat16*)x, (__nv_bfloat16*)out, rows, x_stride, out_stride, eps, gamma, beta, shift, scale,  \
      mod_row_stride, row_index, rows_per_group)
  if (!mod) {
    if (aff) LN2(true, false, false);
    else LN2(false, false, false);
  } else if (steps) {
    if (aff) LN2(true, true, true);
    else LN2(false, true, true);
  } else {
    if (aff) LN2(true, true, false);
    else LN2(false, true, false);
  }
#undef LN2
}

// Block-per-row variant for wide rows: thread c owns 16-byte chunk c of every row the CTA processes, so the LayerNorm
// affine and the AdaLN shift / scale values for its 8 columns stay in REGISTERS across rows and are re-read only when
// the modulation row changes (with the warp-per-row kernel every row pulled 24 KB of fp32 modulation through L1 for
// 12 KB of activation traffic, and L1 bandwidth, not HBM, set the pace: profiles/r01_rows_ncu.txt).
constexpr int LN_ROWS_PER_BLOCK = 16;
// LayerNorm variant: 0 = warp per row, 1 = block per row one row at a time (2.4 TB/s vs 3.6 TB/s for the warp kernel,
// r01), 2 = batched block-per-row kernel (ln_rows_kernel; 1.3-1.6 TB/s, kept as a measured dead end), 3 = packed
// warp-per-row kernel (ln_modulate_kernel2; 5.3-5.9 TB/s, the default for dim 1024/2048/3072/4096/5120)
static int g_ln_block_kernel = 3;
// q/k norm variant: 0 = warp per row, 1 = block per token, 2 = packed warp per row (qk_rms_rope_kernel2)
static int g_qk_block_kernel = 2;
void rows_set_variant(int ln_block, int qk_block) {
  g_ln_block_kernel = ln_block;
  g_qk_block_kernel = qk_block;
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

// Batched block-per-row kernel (the default for wide rows). Same ownership as above — thread c owns 16-byte chunk c of
// every row, so the affine / modulation values of its 8 columns live in registers — but built around the two things
// that held the one-row-at-a-time version at 2.4 TB/s:
//   * bytes in flight: R rows are processed per iteration and the NEXT R rows are already being fetched (64 B per
//     thread outstanding; 2 CTAs x 384 threads per SM = 48 KB per SM, enough to cover HBM latency at the read rate a
//     read+write stream needs);
//   * barriers: one __syncthreads per R rows. Each warp reduces its own 256 elements to (sum, M2 about the warp mean)
//     with shuffles and the CTA combines the per-warp pairs with Chan's parallel-variance formula
//     (M2 = sum M2_w + sum n_w (mean_w - mean)^2) — as accurate as the two-pass form, one block-wide exchange.
// The reduction scratch is double buffered, so iteration i+1 may write while a slow warp still reads iteration i.
constexpr int LN3_R = 4;
constexpr int LN3_MAX_THREADS = 384;  // dim <= 3072

template <bool AFFINE, bool MOD>
__global__ void __launch_bounds__(LN3_MAX_THREADS, 2)
ln_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t rows, int dim,
               int64_t x_stride, int64_t out_stride, float eps, const float* __restrict__ gamma,
               const float* __restrict__ beta, const float* __restrict__ shift, const float* __restrict__ scale,
               int64_t mod_row_stride, const int32_t* __restrict__ row_index, int64_t rows_per_group, int flags,
               int rows_per_cta) {
  constexpr int R = LN3_R;
  __shared__ __align__(16) float2 red[2][R][LN3_MAX_THREADS / 32];
  const int c = threadIdx.x;
  const int lane = c & 31, warp = c >> 5;
  const int nwarps = blockDim.x >> 5;
  const int nchunks = dim >> 3;
  const bool active = c < nchunks;
  const bool steps = (flags & LN_FLAG_BF16_STEPS) != 0;
  const float inv_dim = 1.0f / (float)dim;
  // elements held by this warp (the last warp of a row may be partial)
  const int n_mine = 8 * max(0, min(32, nchunks - 32 * warp));
  const float inv_n_mine = n_mine > 0 ? 1.0f / (float)n_mine : 0.f;
  const float n_last = (float)(8 * (nchunks - 32 * (nwarps - 1)));  // elements of the row's last warp
  const float inv_n_last = 1.0f / n_last;

  float g8[8], b8[8], sc8[8], sh8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    g8[e] = 1.f;
    b8[e] = 0.f;
    sc8[e] = 1.f;  // holds 1 + scale
    sh8[e] = 0.f;
  }
  if (AFFINE && active) {
    ld8f(gamma + c * 8, g8);
    if (beta != nullptr) ld8f(beta + c * 8, b8);
  }
  int64_t cached_g = -1;

  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(rows, r0 + (int64_t)rows_per_cta);
  uint4 nxt[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    nxt[r] = make_uint4(0, 0, 0, 0);
    if (active && r0 + r < r1) nxt[r] = ld_stream(reinterpret_cast<const uint4*>(x + (r0 + r) * x_stride) + c);
  }
  int it = 0;
  for (int64_t base = r0; base < r1; base += R, ++it) {
    uint4 cur[R];
#pragma unroll
    for (int r = 0; r < R; ++r) cur[r] = nxt[r];
#pragma unroll
    for (int r = 0; r < R; ++r) {  // prefetch the next batch
      nxt[r] = make_uint4(0, 0, 0, 0);
      const int64_t row = base + R + r;
      if (active && row < r1) nxt[r] = ld_stream(reinterpret_cast<const uint4*>(x + row * x_stride) + c);
    }
    int gi[R];
    if (MOD) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t row = min(base + r, r1 - 1);
        gi[r] = row_index ? __ldg(row_index + row) : (int)(row / rows_per_group);
      }
    }
    // per-warp (sum, M2)
    float s[R], m2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float v[8];
      unpack8(cur[r], v);
      float a = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) a += v[e];
      s[r] = a;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) s[r] = warp_sum(s[r]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float v[8];
      unpack8(cur[r], v);
      const float mw = s[r] * inv_n_mine;
      float a = 0.f;
      if (active) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = v[e] - mw;
          a += d * d;
        }
      }
      m2[r] = a;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) m2[r] = warp_sum(m2[r]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) red[it & 1][r][warp] = make_float2(s[r], m2[r]);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = base + r;
      if (row >= r1) break;  // block-uniform
      // every thread combines the per-warp pairs itself (broadcast shared-memory reads, no second barrier)
      constexpr int MAXW = LN3_MAX_THREADS / 32;
      const float4* rp = reinterpret_cast<const float4*>(&red[it & 1][r][0]);
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < MAXW; w += 2) {
        const float4 t = rp[w >> 1];
        if (w < nwarps) tot += t.x;
        if (w + 1 < nwarps) tot += t.z;
      }
      const float mean = tot * inv_dim;
      float m2t = 0.f;
#pragma unroll
      for (int w = 0; w < MAXW; w += 2) {  // second read of the pairs instead of 24 live registers
        const float4 t = rp[w >> 1];
        if (w < nwarps) {
          const bool last = w == nwarps - 1;
          const float dm = t.x * (last ? inv_n_last : (1.0f / 256.0f)) - mean;
          m2t += t.y + (last ? n_last : 256.0f) * dm * dm;
        }
        if (w + 1 < nwarps) {
          const bool last = w + 1 == nwarps - 1;
          const float dm = t.z * (last ? inv_n_last : (1.0f / 256.0f)) - mean;
          m2t += t.w + (last ? n_last : 256.0f) * dm * dm;
        }
      }
      const float rstd = rsqrtf(m2t * inv_dim + eps);
      if (MOD) {
        if ((int64_t)gi[r] != cached_g) {  // block-uniform
          cached_g = gi[r];
          if (active) {
            ld8f(scale + cached_g * mod_row_stride + c * 8, sc8);
            ld8f(shift + cached_g * mod_row_stride + c * 8, sh8);
#pragma unroll
            for (int e = 0; e < 8; ++e) sc8[e] = 1.0f + sc8[e];
          }
        }
      }
      if (active) {
        float v[8], y[8];
        unpack8(cur[r], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = (v[e] - mean) * rstd;
        if (AFFINE) {
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = y[e] * g8[e] + b8[e];
        }
        if (MOD) {
          if (steps) {
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = rbf(rbf(y[e]) * rbf(sc8[e])) + sh8[e];
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = y[e] * sc8[e] + sh8[e];
          }
        }
        reinterpret_cast<uint4*>(out + row * out_stride)[c] = pack8(y);
      }
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
  if (g_ln_block_kernel == 2 && dim >= 1024 && dim <= 8 * LN3_MAX_THREADS && rows >= 64) {
    // ~2 CTAs per SM, each walking a contiguous run of rows (a multiple of the batch size R)
    const int threads = ((dim / 8 + 31) / 32) * 32;
    const int64_t want_ctas = (int64_t)num_sms() * 2;
    int64_t per = (rows + want_ctas - 1) / want_ctas;
    per = (per + LN3_R - 1) / LN3_R * LN3_R;
    const unsigned ctas = (unsigned)((rows + per - 1) / per);
#define LAUNCH_LN3(A, M)                                                                                          \
  ln_rows_kernel<A, M><<<ctas, threads, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, rows, dim,     \
                                                     x_stride, out_stride, eps, gamma, beta, shift, scale,        \
                                                     mod_row_stride, row_index, rows_per_group, flags, (int)per)
    const bool aff = gamma != nullptr, mod = shift != nullptr;
    if (aff && mod) LAUNCH_LN3(true, true);
    else if (aff) LAUNCH_LN3(true, false);
    else if (mod) LAUNCH_LN3(false, true);
    else LAUNCH_LN3(false, false);
#undef LAUNCH_LN3
    FINO_CHECK_CUDA(cudaGetLastError());
    return FINO_OK;
  }
  if (g_ln_block_kernel == 1 && dim >= 1024 && rows >= 64) {  // experimental block-per-row kernel (slower today)
    const int threads = ((dim / 8 + 31) / 32) * 32;
    dim3 bgrid((unsigned)((rows + LN_ROWS_PER_BLOCK - 1) / LN_ROWS_PER_BLOCK));
    ln_modulate_block_kernel<<<bgrid, threads, 0, stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)out, rows, dim, x_stride, out_stride, eps, gamma, beta, shift, scale,
        mod_row_stride, row_index, rows_per_group, flags);
    FINO_CHECK_CUDA(cudaGetLastError());
    return FINO_OK;
  }
  if (g_ln_block_kernel == 3 && (dim == 1024 || dim == 2048 || dim == 3072 || dim == 4096 || dim == 5120)) {
    const bool aff = gamma != nullptr, mod = shift != nullptr, steps = (flags & LN_FLAG_BF16_STEPS) != 0;
#define LN2_ARGS grid, stream, aff, mod, steps, x, out, rows, x_stride, out_stride, eps, gamma, beta, shift, scale, \
                 mod_row_stride, row_index, rows_per_group
    switch (dim / 128) {
      case 8: launch_ln2<8>(LN2_ARGS); break;
      case 16: launch_ln2<16>(LN2_ARGS); break;
      case 24: launch_ln2<24>(LN2_ARGS); break;
      case 32: launch_ln2<32>(LN2_ARGS); break;
      default: launch_ln2<40>(LN2_ARGS); break;
    }
#undef LN2_ARGS
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

// Per-head LayerNorm(64) + CogVideoX RoPE (attention_processor.py:2848-2860, embeddings.py:1247-1256), dim = CPL * 256.
// Warp per row; lane l owns the 16-byte chunk c = i*32 + l of every 256-column group i, so a 64-wide head is the 8
// consecutive lanes [8g, 8g+8) of ONE group and a lane's 8 columns sit at the same position (l % 8) of 12 different
// heads: the norm gain/bias and the token's cos/sin values are loaded once per row, and the per-head statistics are
// 3-step butterflies over 8 lanes that run side by side for all CPL heads of the lane (two rounds of 3 shuffle steps,
// each step CPL independent shuffles — the generic kernel above walks the heads one after the other, 72 dependent
// shuffles per row). The row stays packed (bf16) in registers, 16 warps per SM. q rows on even CTAs, k rows on odd
// ones, so the two rows of a token run side by side and share the cos/sin lines.
template <int CPL, int G>  // G = 256-column groups in flight per pass (CPL % G == 0)
__global__ void __launch_bounds__(ROW_WARPS * 32, (G <= 2 ? 4 : G <= 6 ? 3 : 2)) qk_ln64_rope_kernel(const QkParams p) {
  const int lane = threadIdx.x & 31;
  const int which = blockIdx.x & 1;
  const QkTensor& t = p.t[which];
  const int64_t row = (int64_t)(blockIdx.x >> 1) * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= t.rows) return;
  uint4* xr = reinterpret_cast<uint4*>(t.ptr + row * t.row_stride) + lane;
  const int hc = lane & 7;  // chunk inside the head
  const int64_t s_in_seq = row % p.seq_len;
  const bool do_rope = t.rope && p.rope_mode == ROPE_COGVIDEOX && s_in_seq >= p.rope_skip;
  uint4 wq = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);  // bf16 ones
  uint4 bq = make_uint4(0u, 0u, 0u, 0u);
  if (t.weight != nullptr) wq = __ldg(reinterpret_cast<const uint4*>(t.weight) + hc);
  if (t.bias != nullptr) bq = __ldg(reinterpret_cast<const uint4*>(t.bias) + hc);
  float cs[8], sn[8];
  if (do_rope) {
    ld8f(p.cos + (s_in_seq - p.rope_skip) * 64 + hc * 8, cs);
    ld8f(p.sin + (s_in_seq - p.rope_skip) * 64 + hc * 8, sn);
  }
#pragma unroll 1
  for (int g0 = 0; g0 < CPL; g0 += G) {
    uint4 raw[G];
#pragma unroll
    for (int i = 0; i < G; ++i) raw[i] = xr[(g0 + i) * 32];
    float mean[G], rstd[G];
#pragma unroll
    for (int i = 0; i < G; ++i) {
      float v[8];
      unpack8(raw[i], v);
      mean[i] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < G; ++i) mean[i] += __shfl_xor_sync(0xffffffffu, mean[i], o);
#pragma unroll
    for (int i = 0; i < G; ++i) {
      mean[i] *= (1.0f / 64.0f);
      float v[8];
      unpack8(raw[i], v);
      float sq = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[e] - mean[i];
        sq = fmaf(d, d, sq);
      }
      rstd[i] = sq;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < G; ++i) rstd[i] += __shfl_xor_sync(0xffffffffu, rstd[i], o);
    float w[8], b[8];
    unpack8(wq, w);
    unpack8(bq, b);
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const float r = rsqrtf(rstd[i] * (1.0f / 64.0f) + p.eps);
      float v[8];
      unpack8(raw[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = rbf((v[e] - mean[i]) * r * w[e] + b[e]);
      if (do_rope) {
        // x*cos + rotate(x)*sin with rotate = (-x_odd, x_even), separate roundings as the reference's elementwise ops
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float xe = v[e], xo = v[e + 1];
          o[e] = __fadd_rn(__fmul_rn(xe, cs[e]), __fmul_rn(-xo, sn[e]));
          o[e + 1] = __fadd_rn(__fmul_rn(xo, cs[e + 1]), __fmul_rn(xe, sn[e + 1]));
        }
        xr[(g0 + i) * 32] = pack8(o);
      } else {
        xr[(g0 + i) * 32] = pack8(v);
      }
    }
  }
}

// Warp-per-row RMSNorm-across-heads (+ Wan RoPE) in the 4-column-group layout with packed f32x2 arithmetic (see
// ln_modulate_kernel2): dim == GPL * 128. With head_dim == 128 a lane's 4 columns sit at the same position of every
// head, so ONE float4 of cos and one of sin per token serve all heads of the row. Even CTAs take tensor 0 (q), odd
// CTAs tensor 1 (k), so the q and k rows of a token run side by side and share the cos/sin lines in L2.
// Cast points as in qk_norm_rope_kernel: fp32 rsqrt -> bf16 -> x weight (bf16 product) -> fp32 rotate (separate,
// unfused multiplies and adds, like the reference's elementwise ops) -> bf16.
template <int GPL, bool ROPE>
__global__ void __launch_bounds__(ROW_WARPS * 32, (GPL <= 24 ? 2 : 1)) qk_rms_rope_kernel2(const QkParams p) {
  constexpr int dim = GPL * 128;
  const int lane = threadIdx.x & 31;
  const int which = blockIdx.x & 1;
  const QkTensor& t = p.t[which];
  const int64_t row = (int64_t)(blockIdx.x >> 1) * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= t.rows) return;
  uint2* xr = reinterpret_cast<uint2*>(t.ptr + row * t.row_stride) + lane;
  uint2 v[GPL];
#pragma unroll
  for (int i = 0; i < GPL; ++i) v[i] = xr[i * 32];
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), sn = cs;
  bool do_rope = false;
  if (ROPE) {
    const int64_t s_in_seq = row % p.seq_len;
    do_rope = t.rope && s_in_seq >= p.rope_skip;
    if (do_rope) {
      cs = __ldg(reinterpret_cast<const float4*>(p.cos + (s_in_seq - p.rope_skip) * 128) + lane);
      sn = __ldg(reinterpret_cast<const float4*>(p.sin + (s_in_seq - p.rope_skip) * 128) + lane);
    }
  }
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
  for (int i = 0; i < GPL; ++i) {
    const float a0 = bf16_lo_to_f32(v[i].x), a1 = bf16_hi_to_f32(v[i].x);
    const float a2 = bf16_lo_to_f32(v[i].y), a3 = bf16_hi_to_f32(v[i].y);
    ffma2(q0, q1, a0, a1, a0, a1, q0, q1);
    ffma2(q2, q3, a2, a3, a2, a3, q2, q3);
  }
  const float rstd = rsqrtf(warp_sum((q0 + q1) + (q2 + q3)) * (1.0f / (float)dim) + p.eps);
  const uint2* wr = reinterpret_cast<const uint2*>(t.weight) + lane;
  // rotation coefficients of this lane's two pairs: (x1, x2) -> (x1 c - x2 s, x2 c + x1 s), c = cos[even], s = sin[odd]
  const float c01 = cs.x, s01 = sn.y, c23 = cs.z, s23 = sn.w;
#pragma unroll
  for (int i = 0; i < GPL; ++i) {
    float a0, a1, a2, a3;
    fmul2(a0, a1, bf16_lo_to_f32(v[i].x), bf16_hi_to_f32(v[i].x), rstd, rstd);
    fmul2(a2, a3, bf16_lo_to_f32(v[i].y), bf16_hi_to_f32(v[i].y), rstd, rstd);
    uint32_t n01 = pack_bf16x2(a0, a1), n23 = pack_bf16x2(a2, a3);  // .to(bf16)
    if (t.weight != nullptr) {
      const uint2 w = __ldg(wr + i * 32);
      fmul2(a0, a1, bf16_lo_to_f32(n01), bf16_hi_to_f32(n01), bf16_lo_to_f32(w.x), bf16_hi_to_f32(w.x));
      fmul2(a2, a3, bf16_lo_to_f32(n23), bf16_hi_to_f32(n23), bf16_lo_to_f32(w.y), bf16_hi_to_f32(w.y));
      n01 = pack_bf16x2(a0, a1);  // bf16 * bf16 product rounded to bf16
      n23 = pack_bf16x2(a2, a3);
    }
    if (ROPE && do_rope) {
      const float x0 = bf16_lo_to_f32(n01), x1 = bf16_hi_to_f32(n01);
      const float x2 = bf16_lo_to_f32(n23), x3 = bf16_hi_to_f32(n23);
      float pa, pb, ra, rb, o0, o1, o2, o3;
      fmul2(pa, pb, x0, x1, c01, c01);    // (x1 c, x2 c)
      fmul2(ra, rb, x1, x0, -s01, s01);   // (-(x2 s), x1 s): the sign folds into the exact product
      fadd2(o0, o1, pa, pb, ra, rb);
      fmul2(pa, pb, x2, x3, c23, c23);
      fmul2(ra, rb, x3, x2, -s23, s23);
      fadd2(o2, o3, pa, pb, ra, rb);
      n01 = pack_bf16x2(o0, o1);
      n23 = pack_bf16x2(o2, o3);
    }
    xr[i * 32] = make_uint2(n01, n23);
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
  // CogVideoX: per-head LayerNorm over 64-wide heads (+ RoPE on the video tokens), rows of CPL * 256 columns
  if (g_qk_block_kernel != 0 && norm_mode == QK_LAYERNORM_PER_HEAD && head_dim == 64 && (dim == 3072 || dim == 256) &&
      (p.rope_mode == ROPE_NONE || p.rope_mode == ROPE_COGVIDEOX)) {
    const int64_t nb = std::max(p.blocks0, blocks1);
    dim3 g2((unsigned)(2 * nb));
    if (x1 == nullptr) p.t[1].rows = 0;
    if (dim == 3072) {
      if (g_qk_block_kernel == 3) qk_ln64_rope_kernel<12, 12><<<g2, ROW_WARPS * 32, 0, stream>>>(p);
      else if (g_qk_block_kernel == 4) qk_ln64_rope_kernel<12, 6><<<g2, ROW_WARPS * 32, 0, stream>>>(p);
      else if (g_qk_block_kernel == 5) qk_ln64_rope_kernel<12, 3><<<g2, ROW_WARPS * 32, 0, stream>>>(p);
      else if (g_qk_block_kernel == 6) qk_ln64_rope_kernel<12, 2><<<g2, ROW_WARPS * 32, 0, stream>>>(p);
      else qk_ln64_rope_kernel<12, 4><<<g2, ROW_WARPS * 32, 0, stream>>>(p);  // measured best: 90.7 us, 5.3 TB/s
    } else {
      qk_ln64_rope_kernel<1, 1><<<g2, ROW_WARPS * 32, 0, stream>>>(p);
    }
    FINO_CHECK_CUDA(cudaGetLastError());
    return FINO_OK;
  }
  // packed warp-per-row kernel: RMS across heads, dim 3072 / 5120, head_dim 128 when rotating, no bias, row strides
  // that keep the 8-byte pieces aligned (always true: strides are multiples of 8 elements)
  if (g_qk_block_kernel == 2 && norm_mode == QK_RMS_ACROSS_HEADS && (dim == 3072 || dim == 5120) && b0 == nullptr &&
      b1 == nullptr && rope_skip == 0 && (p.rope_mode == ROPE_NONE || (p.rope_mode == ROPE_WAN && head_dim == 128))) {
    const int64_t nb = std::max(p.blocks0, blocks1);
    dim3 g2((unsigned)(2 * nb));
    if (x1 == nullptr) p.t[1].rows = 0;
    const bool rope = p.rope_mode == ROPE_WAN;
    if (dim == 3072) {
      if (rope) qk_rms_rope_kernel2<24, true><<<g2, ROW_WARPS * 32, 0, stream>>>(p);
      else qk_rms_rope_kernel2<24, false><<<g2, ROW_WARPS * 32, 0, stream>>>(p);
    } else {
      if (rope) qk_rms_rope_kernel2<40, true><<<g2, ROW_WARPS * 32, 0, stream>>>(p);
      else qk_rms_rope_kernel2<40, false><<<g2, ROW_WARPS * 32, 0, stream>>>(p);
    }
    FINO_CHECK_CUDA(cudaGetLastError());
    return FINO_OK;
  }
  if (g_qk_block_kernel >= 1 && norm_mode == QK_RMS_ACROSS_HEADS && x1 != nullptr && rows0 == rows1 && rows0 >= 64 && dim >= 1024 && dim <= 4096 &&
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
