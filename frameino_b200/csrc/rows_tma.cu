// HBM-bound row kernels, TMA-staged: persistent CTAs (one per SM) stream blocks of token rows through a shared-memory
// ring with bulk async copies (cp.async.bulk + mbarrier), so the bytes in flight per SM are set by the ring size
// (~190 KB) instead of by registers and occupancy; results go back to HBM with bulk stores from the same buffers.
//
//   ln_rows_tma_kernel     FP32LayerNorm (+affine) (+AdaLN modulate)    reference transformer_wan.py:334,339,344-346,536
//   qk_rows_tma_kernel     RMSNorm-across-heads + 3-D RoPE of q and k    reference transformer_wan.py:64-90
//                          (in place, or with the stores scattered to the Ulysses peers: peer_kernels.cu)
//
// Inside a CTA thread c owns the 16-byte chunk c (8 columns) of every row: per-column constants (LayerNorm affine,
// AdaLN shift/scale of the current timestep row, RMSNorm weights) live in registers for the whole kernel, rows are
// processed RB at a time so that the two block-wide reductions cost two barriers per RB rows.
// The warp-per-row kernels in norm_kernels.cu remain for narrow rows (tiny test configs) and odd layouts.
#include "common.cuh"
#include "ptx.cuh"

namespace fino {

namespace {

__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
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
__device__ __forceinline__ void ld8f(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int RB = 8;            // rows per ring stage
constexpr int MAX_WARPS = 16;    // consumer warps (dim <= 4096: 512 threads x 8 columns)
constexpr int kRingBudget = 196608;  // bytes of shared memory for the ring

// Sums RB per-thread partials over the CTA's consumer threads: warp shuffle, one shared-memory exchange, one named
// barrier. `red` is [RB][MAX_WARPS] floats; every thread returns with the RB totals in tot[].
__device__ __forceinline__ void block_sum_rb(float (&part)[RB], float (&tot)[RB], float* red, int warp, int lane,
                                             int nwarps, int nthreads) {
#pragma unroll
  for (int r = 0; r < RB; ++r) part[r] = warp_sum(part[r]);
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < RB; ++r) red[r * MAX_WARPS + warp] = part[r];
  }
  named_bar_sync(1, nthreads);
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += red[r * MAX_WARPS + w];
    tot[r] = t;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// LayerNorm (+affine) (+modulate)
// ------------------------------------------------------------------------------------------------
struct LnTmaParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
  int64_t rows, x_stride, out_stride;
  int dim;
  float eps;
  const float* gamma;
  const float* beta;
  const float* shift;
  const float* scale;
  int64_t mod_row_stride;
  const int32_t* row_index;
  int64_t rows_per_group;
  int bf16_steps;
  int stages;
};

__global__ void __launch_bounds__(MAX_WARPS * 32 + 32, 1) ln_rows_tma_kernel(const __grid_constant__ LnTmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t ring = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t row_bytes = (uint32_t)p.dim * 2u;
  const uint32_t stage_bytes = RB * row_bytes;
  const uint32_t bars = ring + (uint32_t)p.stages * stage_bytes;  // full[stages], empty[stages]
  float* red = reinterpret_cast<float*>(smem_raw + (bars + 16u * p.stages - smem_u32(smem_raw)));  // 2 x [RB][MAX_WARPS]
  const int nthreads = blockDim.x - 32;  // consumers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nwarps = nthreads >> 5;
  const int64_t nblocks = (p.rows + RB - 1) / RB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bars + 8u * s, 1);
      mbar_init(bars + 8u * (p.stages + s), 1);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == nwarps) {
    // ===================== producer warp =====================
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        mbar_wait_relaxed(bars + 8u * (p.stages + s), phase ^ 1u, 500 + s);
        const int64_t r0 = blk * RB;
        const int nr = (int)min((int64_t)RB, p.rows - r0);
        mbar_arrive_expect_tx(bars + 8u * s, (uint32_t)nr * row_bytes);
        const uint32_t dst = ring + (uint32_t)s * stage_bytes;
        if (p.x_stride == p.dim) {
          bulk_load(dst, p.x + r0 * p.x_stride, (uint32_t)nr * row_bytes, bars + 8u * s);
        } else {
          for (int r = 0; r < nr; ++r) bulk_load(dst + r * row_bytes, p.x + (r0 + r) * p.x_stride, row_bytes, bars + 8u * s);
        }
        if (++s == p.stages) {
          s = 0;
          phase ^= 1u;
        }
      }
    }
    return;
  }

  // ===================== consumers: thread c owns columns [8c, 8c+8) =====================
  const int c = threadIdx.x;
  const bool active = c < (p.dim >> 3);
  const float inv_dim = 1.0f / (float)p.dim;
  float g8[8], b8[8], sc8[8], sh8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) g8[e] = 1.f, b8[e] = 0.f, sc8[e] = 0.f, sh8[e] = 0.f;
  if (active && p.gamma != nullptr) ld8f(p.gamma + c * 8, g8);
  if (active && p.beta != nullptr) ld8f(p.beta + c * 8, b8);
  int64_t cached_g = -1;
  int s = 0, prev_s = -1;
  uint32_t phase = 0;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t r0 = blk * RB;
    const int nr = (int)min((int64_t)RB, p.rows - r0);
    const uint32_t base = ring + (uint32_t)s * stage_bytes + (uint32_t)c * 16u;
    mbar_wait(bars + 8u * s, phase, 510 + s);
    uint4 raw[RB];  // the rows stay packed (4 registers each) and are unpacked where used: 3x a few ALU ops
    float part[RB], tot[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      part[r] = 0.f;
      raw[r] = make_uint4(0, 0, 0, 0);
      if (active && r < nr) {
        raw[r] = lds128(base + r * row_bytes);
        float v[8];
        unpack8(raw[r], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) part[r] += v[e];
      }
    }
    block_sum_rb(part, tot, red, warp, lane, nwarps, nthreads);
    float mean[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      mean[r] = tot[r] * inv_dim;
      part[r] = 0.f;
      if (active) {
        float v[8];
        unpack8(raw[r], v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = v[e] - mean[r];
          part[r] += d * d;
        }
      }
    }
    block_sum_rb(part, tot, red + RB * MAX_WARPS, warp, lane, nwarps, nthreads);
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      if (r < nr) {  // block-uniform
        const float rstd = rsqrtf(tot[r] * inv_dim + p.eps);
        if (p.shift != nullptr) {
          const int64_t row = r0 + r;
          const int64_t g = p.row_index ? (int64_t)__ldg(p.row_index + row) : row / p.rows_per_group;
          if (g != cached_g) {  // block-uniform: the modulation row changes a handful of times per forward
            cached_g = g;
            if (active) {
              ld8f(p.scale + g * p.mod_row_stride + c * 8, sc8);
              ld8f(p.shift + g * p.mod_row_stride + c * 8, sh8);
            }
          }
        }
        if (active) {
          float y[8];
          unpack8(raw[r], y);
#pragma unroll
          for (int e = 0; e < 8; ++e) y[e] = (y[e] - mean[r]) * rstd;
          if (p.gamma != nullptr) {
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = y[e] * g8[e] + b8[e];
          }
          if (p.shift != nullptr) {
            if (p.bf16_steps) {
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = rbf(rbf(y[e]) * rbf(1.0f + sc8[e])) + sh8[e];
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = y[e] * (1.0f + sc8[e]) + sh8[e];
            }
          }
          sts128(base + r * row_bytes, pack8(y));
        }
      }
    }
    fence_proxy_async_smem();      // generic-proxy writes -> visible to the bulk (async proxy) store
    named_bar_sync(1, nthreads);   // every chunk of the stage is written
    if (threadIdx.x == 0) {
      const uint32_t src = ring + (uint32_t)s * stage_bytes;
      if (p.out_stride == p.dim) {
        bulk_store(p.out + r0 * p.out_stride, src, (uint32_t)nr * row_bytes);
      } else {
        for (int r = 0; r < nr; ++r) bulk_store(p.out + (r0 + r) * p.out_stride, src + r * row_bytes, row_bytes);
      }
      bulk_commit();
      if (prev_s >= 0) {  // the previous stage's store has finished READING shared memory: hand the stage back
        bulk_wait_read<1>();
        mbar_arrive(bars + 8u * (p.stages + prev_s));
      }
    }
    prev_s = s;
    if (++s == p.stages) {
      s = 0;
      phase ^= 1u;
    }
  }
  if (threadIdx.x == 0) bulk_wait_read<0>();  // shared memory must outlive the last store's reads
}

int ln_modulate_tma(const void* x, void* out, int64_t rows, int dim, int64_t x_stride, int64_t out_stride, float eps,
                    const float* gamma, const float* beta, const float* shift, const float* scale,
                    int64_t mod_row_stride, const int32_t* row_index, int64_t rows_per_group, int flags,
                    cudaStream_t stream) {
  LnTmaParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.rows = rows;
  p.x_stride = x_stride;
  p.out_stride = out_stride;
  p.dim = dim;
  p.eps = eps;
  p.gamma = gamma;
  p.beta = beta;
  p.shift = shift;
  p.scale = scale;
  p.mod_row_stride = mod_row_stride;
  p.row_index = row_index;
  p.rows_per_group = rows_per_group > 0 ? rows_per_group : (int64_t)1 << 62;
  p.bf16_steps = flags & 1;
  const int stage_bytes = RB * dim * 2;
  int stages = kRingBudget / stage_bytes;
  if (stages > 4) stages = 4;
  FINO_CHECK_ARG(stages >= 2, "ln_modulate_tma: dim %d too wide for a 2-stage ring", dim);
  p.stages = stages;
  const int smem = stages * stage_bytes + 16 * stages + 2 * RB * MAX_WARPS * 4 + 256;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    FINO_CHECK_CUDA(cudaFuncSetAttribute(ln_rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  const int threads = ((dim / 8 + 31) / 32) * 32 + 32;
  const int64_t nblocks = (rows + RB - 1) / RB;
  int grid = num_sms();
  if (nblocks < grid) grid = (int)nblocks;
  ln_rows_tma_kernel<<<grid, threads, smem, stream>>>(p);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// q/k RMSNorm across heads + Wan RoPE (in place, or scattered to the Ulysses peers)
// ------------------------------------------------------------------------------------------------
constexpr int QT = RB / 2;  // tokens per ring stage: QT q rows + QT k rows = RB sum-of-squares reductions

struct QkTmaParams {
  __nv_bfloat16* q;   // [rows, q_stride]
  __nv_bfloat16* k;   // [rows, k_stride] or null (cross-attention q-only norm)
  const __nv_bfloat16* v;  // scatter only
  int64_t rows, q_stride, k_stride, v_stride;
  const __nv_bfloat16* wq;
  const __nv_bfloat16* wk;
  int heads, head_dim;
  float eps;
  const float* cos;  // [seq_len, head_dim] fp32 or null (no RoPE)
  const float* sin;
  int64_t seq_len;
  int stages;
  // scatter (peer_kernels.cu: fino_qkv_norm_rope_scatter)
  __nv_bfloat16* dst[8];
  int world, rank, inner;
  int64_t rows_per_rank, dst_row_stride;
};

template <bool SCATTER>
__global__ void __launch_bounds__(MAX_WARPS * 32 + 32, 1) qk_rows_tma_kernel(const __grid_constant__ QkTmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t ring = (smem_u32(smem_raw) + 127u) & ~127u;
  const int dim = p.heads * p.head_dim;
  const uint32_t row_bytes = (uint32_t)dim * 2u;
  const uint32_t tab_bytes = (uint32_t)p.head_dim * 4u;
  const bool has_k = p.k != nullptr;
  const bool rope = p.cos != nullptr;
  // stage layout: q rows | k rows | cos rows | sin rows
  const uint32_t k_off = QT * row_bytes;
  const uint32_t cos_off = k_off + (has_k ? QT * row_bytes : 0u);
  const uint32_t sin_off = cos_off + (rope ? QT * tab_bytes : 0u);
  const uint32_t stage_bytes = (sin_off + (rope ? QT * tab_bytes : 0u) + 127u) & ~127u;
  const uint32_t bars = ring + (uint32_t)p.stages * stage_bytes;
  float* red = reinterpret_cast<float*>(smem_raw + (bars + 16u * p.stages - smem_u32(smem_raw)));
  const int nthreads = blockDim.x - 32;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nwarps = nthreads >> 5;
  const int64_t nblocks = (p.rows + QT - 1) / QT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bars + 8u * s, 1);
      mbar_init(bars + 8u * (p.stages + s), 1);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == nwarps) {
    // ===================== producer warp =====================
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      const uint32_t per_token = row_bytes * (has_k ? 2u : 1u) + (rope ? 2u * tab_bytes : 0u);
      for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        mbar_wait_relaxed(bars + 8u * (p.stages + s), phase ^ 1u, 520 + s);
        const int64_t t0 = blk * QT;
        const int nt = (int)min((int64_t)QT, p.rows - t0);
        const uint32_t full = bars + 8u * s;
        mbar_arrive_expect_tx(full, (uint32_t)nt * per_token);
        const uint32_t dst = ring + (uint32_t)s * stage_bytes;
        for (int t = 0; t < nt; ++t) {
          const int64_t tok = t0 + t;
          bulk_load(dst + t * row_bytes, p.q + tok * p.q_stride, row_bytes, full);
          if (has_k) bulk_load(dst + k_off + t * row_bytes, p.k + tok * p.k_stride, row_bytes, full);
          if (rope) {
            const int64_t trow = tok % p.seq_len;
            bulk_load(dst + cos_off + t * tab_bytes, p.cos + trow * p.head_dim, tab_bytes, full);
            bulk_load(dst + sin_off + t * tab_bytes, p.sin + trow * p.head_dim, tab_bytes, full);
          }
        }
        if (++s == p.stages) {
          s = 0;
          phase ^= 1u;
        }
      }
    }
    return;
  }

  // ===================== consumers: thread c owns columns [8c, 8c+8) of q, k (and v) =====================
  const int c = threadIdx.x;
  const bool active = c < (dim >> 3);
  const float inv_dim = 1.0f / (float)dim;
  float wq[8], wk[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) wq[e] = wk[e] = 1.f;
  if (active && p.wq != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(p.wq) + c), wq);
  if (active && has_k && p.wk != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(p.wk) + c), wk);
  const int col = c * 8;
  const uint32_t tab_col = (uint32_t)(col % p.head_dim) * 4u;
  // scatter destination of this thread's chunk: rank g owns head group g
  const int g = (SCATTER && active) ? col / p.inner : 0;
  __nv_bfloat16* dst_base = nullptr;
  int ichunks = 0;
  if (SCATTER) {
    dst_base = p.dst[g] + (int64_t)p.rank * p.rows_per_rank * p.dst_row_stride + (col - g * p.inner);
    ichunks = p.inner >> 3;
  }
  int s = 0, prev_s = -1;
  uint32_t phase = 0;
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int64_t t0 = blk * QT;
    const int nt = (int)min((int64_t)QT, p.rows - t0);
    const uint32_t stage = ring + (uint32_t)s * stage_bytes;
    uint4 uv[QT];
    if (SCATTER) {  // v rides along un-normalised: fetched straight from global while the ring stage lands
#pragma unroll
      for (int t = 0; t < QT; ++t)
        if (active && t < nt) uv[t] = __ldg(reinterpret_cast<const uint4*>(p.v + (t0 + t) * p.v_stride) + c);
    }
    mbar_wait(bars + 8u * s, phase, 530 + s);
    uint4 rq[QT], rk[QT];
    float part[RB], tot[RB];
#pragma unroll
    for (int t = 0; t < QT; ++t) {
      part[t] = part[QT + t] = 0.f;
      rq[t] = rk[t] = make_uint4(0, 0, 0, 0);
      if (active && t < nt) {
        float f[8];
        rq[t] = lds128(stage + t * row_bytes + c * 16);
        unpack8(rq[t], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) part[t] += f[e] * f[e];
        if (has_k) {
          rk[t] = lds128(stage + k_off + t * row_bytes + c * 16);
          unpack8(rk[t], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) part[QT + t] += f[e] * f[e];
        }
      }
    }
    block_sum_rb(part, tot, red, warp, lane, nwarps, nthreads);
#pragma unroll
    for (int t = 0; t < QT; ++t) {
      if (t < nt && active) {
        const float rq_std = rsqrtf(tot[t] * inv_dim + p.eps);
        const float rk_std = rsqrtf(tot[QT + t] * inv_dim + p.eps);
        float q[8], k[8];
        unpack8(rq[t], q);
        unpack8(rk[t], k);
        // RMSNorm (upstream diffusers): fp32 x*rsqrt -> bf16 -> * weight in bf16
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          q[e] = p.wq ? rbf(rbf(q[e] * rq_std) * wq[e]) : rbf(q[e] * rq_std);
          k[e] = p.wk ? rbf(rbf(k[e] * rk_std) * wk[e]) : rbf(k[e] * rk_std);
        }
        if (rope) {
          float cs[8], sn[8];
          const uint4 c0 = lds128(stage + cos_off + t * tab_bytes + tab_col);
          const uint4 c1 = lds128(stage + cos_off + t * tab_bytes + tab_col + 16);
          const uint4 s0 = lds128(stage + sin_off + t * tab_bytes + tab_col);
          const uint4 s1 = lds128(stage + sin_off + t * tab_bytes + tab_col + 16);
          cs[0] = __uint_as_float(c0.x), cs[1] = __uint_as_float(c0.y), cs[2] = __uint_as_float(c0.z);
          cs[3] = __uint_as_float(c0.w), cs[4] = __uint_as_float(c1.x), cs[5] = __uint_as_float(c1.y);
          cs[6] = __uint_as_float(c1.z), cs[7] = __uint_as_float(c1.w);
          sn[0] = __uint_as_float(s0.x), sn[1] = __uint_as_float(s0.y), sn[2] = __uint_as_float(s0.z);
          sn[3] = __uint_as_float(s0.w), sn[4] = __uint_as_float(s1.x), sn[5] = __uint_as_float(s1.y);
          sn[6] = __uint_as_float(s1.z), sn[7] = __uint_as_float(s1.w);
          // cos = freqs_cos[..., 0::2], sin = freqs_sin[..., 1::2]   (transformer_wan.py:83-87)
          float oq[8], ok[8];
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            oq[e] = __fsub_rn(__fmul_rn(q[e], cs[e]), __fmul_rn(q[e + 1], sn[e + 1]));
            oq[e + 1] = __fadd_rn(__fmul_rn(q[e], sn[e + 1]), __fmul_rn(q[e + 1], cs[e]));
            ok[e] = __fsub_rn(__fmul_rn(k[e], cs[e]), __fmul_rn(k[e + 1], sn[e + 1]));
            ok[e + 1] = __fadd_rn(__fmul_rn(k[e], sn[e + 1]), __fmul_rn(k[e + 1], cs[e]));
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) q[e] = oq[e], k[e] = ok[e];
        }
        if (SCATTER) {
          uint4* drow = reinterpret_cast<uint4*>(dst_base + (t0 + t) * p.dst_row_stride);
          drow[0] = pack8(q);
          drow[ichunks] = pack8(k);
          drow[2 * ichunks] = uv[t];
        } else {
          sts128(stage + t * row_bytes + c * 16, pack8(q));
          if (has_k) sts128(stage + k_off + t * row_bytes + c * 16, pack8(k));
        }
      }
    }
    if (SCATTER) {
      named_bar_sync(1, nthreads);  // every thread is done reading the stage
      if (threadIdx.x == 0) mbar_arrive(bars + 8u * (p.stages + s));
    } else {
      fence_proxy_async_smem();
      named_bar_sync(1, nthreads);
      if (threadIdx.x == 0) {
        for (int t = 0; t < nt; ++t) {
          bulk_store(p.q + (t0 + t) * p.q_stride, stage + t * row_bytes, row_bytes);
          if (has_k) bulk_store(p.k + (t0 + t) * p.k_stride, stage + k_off + t * row_bytes, row_bytes);
        }
        bulk_commit();
        if (prev_s >= 0) {
          bulk_wait_read<1>();
          mbar_arrive(bars + 8u * (p.stages + prev_s));
        }
      }
    }
    prev_s = s;
    if (++s == p.stages) {
      s = 0;
      phase ^= 1u;
    }
  }
  if (!SCATTER && threadIdx.x == 0) bulk_wait_read<0>();
}

template <bool SCATTER>
static int launch_qk_tma(QkTmaParams& p, cudaStream_t stream) {
  const int dim = p.heads * p.head_dim;
  const bool has_k = p.k != nullptr, rope = p.cos != nullptr;
  const int stage_bytes =
      ((QT * dim * 2 * (has_k ? 2 : 1) + (rope ? 2 * QT * p.head_dim * 4 : 0)) + 127) & ~127;
  int stages = kRingBudget / stage_bytes;
  if (stages > 4) stages = 4;
  FINO_CHECK_ARG(stages >= 2, "qk_rows_tma: rows too wide for a 2-stage ring");
  p.stages = stages;
  const int smem = stages * stage_bytes + 16 * stages + RB * MAX_WARPS * 4 + 256;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    FINO_CHECK_CUDA(cudaFuncSetAttribute(qk_rows_tma_kernel<SCATTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  const int threads = ((dim / 8 + 31) / 32) * 32 + 32;
  const int64_t nblocks = (p.rows + QT - 1) / QT;
  int grid = num_sms();
  if (nblocks < grid) grid = (int)nblocks;
  qk_rows_tma_kernel<SCATTER><<<grid, threads, smem, stream>>>(p);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// eligibility: RMS across heads, 1024 <= dim <= 4096, head_dim % 8 == 0 and a multiple of 4 floats, 16-byte aligned rows
bool qk_tma_eligible(int64_t rows, int heads, int head_dim) {
  const int dim = heads * head_dim;
  return rows >= 256 && dim >= 1024 && dim <= MAX_WARPS * 32 * 8 && head_dim % 8 == 0;
}

int qk_rms_rope_tma(void* q, int64_t q_stride, void* k, int64_t k_stride, int64_t rows, const void* wq, const void* wk,
                    int heads, int head_dim, float eps, const float* cos, const float* sin, int64_t seq_len,
                    cudaStream_t stream) {
  QkTmaParams p = {};
  p.q = reinterpret_cast<__nv_bfloat16*>(q);
  p.k = reinterpret_cast<__nv_bfloat16*>(k);
  p.rows = rows;
  p.q_stride = q_stride;
  p.k_stride = k_stride;
  p.wq = reinterpret_cast<const __nv_bfloat16*>(wq);
  p.wk = reinterpret_cast<const __nv_bfloat16*>(wk);
  p.heads = heads;
  p.head_dim = head_dim;
  p.eps = eps;
  p.cos = cos;
  p.sin = sin;
  p.seq_len = seq_len > 0 ? seq_len : (int64_t)1 << 62;
  return launch_qk_tma<false>(p, stream);
}

int qkv_rms_rope_scatter_tma(const void* qkv, int64_t rows, int64_t row_stride, const void* wq, const void* wk,
                             int heads, int head_dim, float eps, const float* cos, const float* sin,
                             void* const* dst_ptrs, int world, int rank, int64_t rows_per_rank,
                             int64_t dst_row_stride, cudaStream_t stream) {
  const int dim = heads * head_dim;
  QkTmaParams p = {};
  p.q = const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(qkv));
  p.k = p.q + dim;
  p.v = p.q + 2 * dim;
  p.rows = rows;
  p.q_stride = p.k_stride = p.v_stride = row_stride;
  p.wq = reinterpret_cast<const __nv_bfloat16*>(wq);
  p.wk = reinterpret_cast<const __nv_bfloat16*>(wk);
  p.heads = heads;
  p.head_dim = head_dim;
  p.eps = eps;
  p.cos = cos;
  p.sin = sin;
  p.seq_len = (int64_t)1 << 62;  // cos/sin are the rows of the local tokens
  for (int r = 0; r < 8; ++r) p.dst[r] = reinterpret_cast<__nv_bfloat16*>(dst_ptrs[r < world ? r : 0]);
  p.world = world;
  p.rank = rank;
  p.inner = dim / world;
  p.rows_per_rank = rows_per_rank;
  p.dst_row_stride = dst_row_stride;
  return launch_qk_tma<true>(p, stream);
}

}  // namespace fino
