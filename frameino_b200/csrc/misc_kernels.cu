// Small helpers around the transformer body: patchify / unpatchify (the Conv3d/Conv2d patch embedding becomes a GEMM
// on the patchified rows), the sinusoidal timestep embedding, a small-M linear for the de-duplicated time MLP and
// the per-layer AdaLN modulation table.
//
//   patchify          reference transformer_wan.py:486-487 (Conv3d k=s=(1,2,2)), embeddings.py:734-738 (Conv2d k=s=2)
//   unpatchify        reference transformer_wan.py:539-543, cogvideox_transformer_3d.py:549-550
//   timestep_embedding reference embeddings.py:27-78
//   linear_small_m    reference transformer_wan.py:182-183 (time_embedder / time_proj on unique timesteps)
//   build_mod_table   reference transformer_wan.py:317-331, 520-527 (scale_shift_table + temb)
#include "common.cuh"
#include "ptx.cuh"

namespace fino {

// ------------------------------------------------------------------------------------------------
// patchify: x[b,c,f,h,w] (arbitrary strides) -> rows[(b,f/pt,h/ph,w/pw), (c,pt,ph,pw)]
// ------------------------------------------------------------------------------------------------
struct PatchParams {
  int B, C, F, H, W;
  int pt, ph, pw;
  int64_t sb, sc, sf, sh, sw;  // element strides of the 5-D tensor
  int64_t ld;                  // row stride of the 2-D matrix
};

__global__ void __launch_bounds__(256)
patchify_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ rows, const PatchParams p,
                int64_t total) {
  const int pf = p.F / p.pt, phh = p.H / p.ph, pww = p.W / p.pw;
  const int kdim = p.C * p.pt * p.ph * p.pw;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % kdim);
    int64_t tok = idx / kdim;
    const int wq = (int)(tok % pww);
    tok /= pww;
    const int hq = (int)(tok % phh);
    tok /= phh;
    const int fq = (int)(tok % pf);
    const int b = (int)(tok / pf);
    int kk = k;
    const int iw = kk % p.pw;
    kk /= p.pw;
    const int ih = kk % p.ph;
    kk /= p.ph;
    const int it = kk % p.pt;
    const int c = kk / p.pt;
    const int64_t src = (int64_t)b * p.sb + (int64_t)c * p.sc + (int64_t)(fq * p.pt + it) * p.sf +
                        (int64_t)(hq * p.ph + ih) * p.sh + (int64_t)(wq * p.pw + iw) * p.sw;
    rows[(idx / kdim) * p.ld + k] = x[src];
  }
}

int patchify(const void* x, void* rows, int B, int C, int F, int H, int W, int pt, int ph, int pw, int64_t sb,
             int64_t sc, int64_t sf, int64_t sh, int64_t sw, int64_t ld, cudaStream_t stream) {
  FINO_CHECK_ARG(x && rows, "patchify: null pointer");
  FINO_CHECK_ARG(B > 0 && C > 0 && F > 0 && H > 0 && W > 0 && pt > 0 && ph > 0 && pw > 0, "patchify: bad shape");
  FINO_CHECK_ARG(F % pt == 0 && H % ph == 0 && W % pw == 0, "patchify: dims not divisible by the patch size");
  PatchParams p{B, C, F, H, W, pt, ph, pw, sb, sc, sf, sh, sw, ld};
  const int64_t total = (int64_t)B * C * F * H * W;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  patchify_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)rows, p, total);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// unpatchify: rows[(b,fq,hq,wq), last] -> out[b,c,f,h,w] (arbitrary output strides)
//   channel_last = 1: last dim ordered (pt,ph,pw,c)  (Wan)      channel_last = 0: (c,pt,ph,pw)  (CogVideoX)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
unpatchify_kernel(const __nv_bfloat16* __restrict__ rows, __nv_bfloat16* __restrict__ out, const PatchParams p,
                  int channel_last, int64_t total) {
  const int pf = p.F / p.pt, phh = p.H / p.ph, pww = p.W / p.pw;
  const int kdim = p.C * p.pt * p.ph * p.pw;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    // idx enumerates the OUTPUT in (b,c,f,h,w) order so that stores are coalesced
    int64_t r = idx;
    const int w = (int)(r % p.W);
    r /= p.W;
    const int h = (int)(r % p.H);
    r /= p.H;
    const int f = (int)(r % p.F);
    r /= p.F;
    const int c = (int)(r % p.C);
    const int b = (int)(r / p.C);
    const int fq = f / p.pt, it = f % p.pt, hq = h / p.ph, ih = h % p.ph, wq = w / p.pw, iw = w % p.pw;
    const int64_t tok = (((int64_t)b * pf + fq) * phh + hq) * pww + wq;
    int k;
    if (channel_last)
      k = ((it * p.ph + ih) * p.pw + iw) * p.C + c;
    else
      k = ((c * p.pt + it) * p.ph + ih) * p.pw + iw;
    out[(int64_t)b * p.sb + (int64_t)c * p.sc + (int64_t)f * p.sf + (int64_t)h * p.sh + (int64_t)w * p.sw] =
        rows[tok * p.ld + k];
    (void)kdim;
  }
}

int unpatchify(const void* rows, void* out, int B, int C, int F, int H, int W, int pt, int ph, int pw, int64_t sb,
               int64_t sc, int64_t sf, int64_t sh, int64_t sw, int64_t ld, int channel_last, cudaStream_t stream) {
  FINO_CHECK_ARG(rows && out, "unpatchify: null pointer");
  FINO_CHECK_ARG(B > 0 && C > 0 && F > 0 && H > 0 && W > 0 && pt > 0 && ph > 0 && pw > 0, "unpatchify: bad shape");
  FINO_CHECK_ARG(F % pt == 0 && H % ph == 0 && W % pw == 0, "unpatchify: dims not divisible by the patch size");
  PatchParams p{B, C, F, H, W, pt, ph, pw, sb, sc, sf, sh, sw, ld};
  const int64_t total = (int64_t)B * C * F * H * W;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  unpatchify_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)rows, (__nv_bfloat16*)out, p,
                                                          channel_last, total);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// sinusoidal timestep embedding (fp32)
// ------------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int n, int dim,
                                          int flip_sin_to_cos, float downscale_freq_shift, float scale,
                                          float max_period) {
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * half) return;
  const int r = idx / half, j = idx % half;
  const float exponent = -logf(max_period) * (float)j / ((float)half - downscale_freq_shift);
  const float arg = scale * (t[r] * expf(exponent));
  const float s = sinf(arg), c = cosf(arg);
  float* o = out + (int64_t)r * dim;
  if (flip_sin_to_cos) {
    o[j] = c;
    o[half + j] = s;
  } else {
    o[j] = s;
    o[half + j] = c;
  }
  if ((dim & 1) && j == 0) o[dim - 1] = 0.f;
}

int timestep_embedding(const float* t, float* out, int n, int dim, int flip_sin_to_cos, float downscale_freq_shift,
                       float scale, float max_period, cudaStream_t stream) {
  FINO_CHECK_ARG(t && out && n > 0 && dim >= 2, "timestep_embedding: bad arguments");
  const int total = n * (dim / 2);
  timestep_embedding_kernel<<<(total + 255) / 256, 256, 0, stream>>>(t, out, n, dim, flip_sin_to_cos,
                                                                     downscale_freq_shift, scale, max_period);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// y[m,n] = act_out( sum_k act_in(x[m,k]) * w[n,k] + b[n] ).  One warp per output column and 8-row chunk
//   (blockIdx.y): built for m <= 8 (the de-duplicated time MLP); larger m runs as ceil(m/8) independent chunks, each
//   row computed exactly as in the m <= 8 case (fp32 FMA chain over k in the same order).
//   x, y fp32; w/b fp32 or bf16.  act: 0 none, 1 SiLU.  round_in / round_out emulate a bf16 module.
// ------------------------------------------------------------------------------------------------
constexpr int SMALL_M_MAX = 8;

template <typename WT>
__device__ __forceinline__ float wload(const WT* p);
template <>
__device__ __forceinline__ float wload<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float wload<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }

// 4 consecutive weights of one output column as floats
__device__ __forceinline__ void wload4(const float* p, float (&w)[4]) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
}
__device__ __forceinline__ void wload4(const __nv_bfloat16* p, float (&w)[4]) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  w[0] = __uint_as_float(v.x << 16), w[1] = __uint_as_float(v.x & 0xffff0000u);
  w[2] = __uint_as_float(v.y << 16), w[3] = __uint_as_float(v.y & 0xffff0000u);
}

// VEC: k % 128 == 0 -> every lane walks the column in float4 steps (4 x fewer load instructions, 4 independent FMA
// chains per row); the accumulation order inside a row differs from the scalar path only in how the k index is dealt to
// the 32 lanes x 4 sub-chains, and is the same for every row and every launch (results do not depend on m).
template <typename WT, bool VEC>
__global__ void __launch_bounds__(256)
linear_small_m_kernel(const float* __restrict__ x, const WT* __restrict__ w, const WT* __restrict__ b,
                      float* __restrict__ y, int m_total, int n, int k, int act_in, int act_out, int round_in,
                      int round_out) {
  extern __shared__ float xs[];  // [m, k] after the input activation
  const int row0 = blockIdx.y * SMALL_M_MAX;
  const int m = min(SMALL_M_MAX, m_total - row0);
  x += (int64_t)row0 * k;
  y += (int64_t)row0 * n;
  for (int i = threadIdx.x; i < m * k; i += blockDim.x) {
    float v = x[i];
    if (act_in == 1) v = silu_exact(v);
    if (round_in) v = __bfloat162float(__float2bfloat16_rn(v));
    xs[i] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (col >= n) return;
  float acc[SMALL_M_MAX];
#pragma unroll
  for (int i = 0; i < SMALL_M_MAX; ++i) acc[i] = 0.f;
  const WT* wr = w + (int64_t)col * k;
  if (VEC) {
    float a4[SMALL_M_MAX][4];
#pragma unroll
    for (int i = 0; i < SMALL_M_MAX; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) a4[i][e] = 0.f;
#pragma unroll 2
    for (int kk = lane * 4; kk < k; kk += 128) {
      float wv[4];
      wload4(wr + kk, wv);
#pragma unroll
      for (int i = 0; i < SMALL_M_MAX; ++i)
        if (i < m) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + i * k + kk);
          a4[i][0] = fmaf(xv.x, wv[0], a4[i][0]);
          a4[i][1] = fmaf(xv.y, wv[1], a4[i][1]);
          a4[i][2] = fmaf(xv.z, wv[2], a4[i][2]);
          a4[i][3] = fmaf(xv.w, wv[3], a4[i][3]);
        }
    }
#pragma unroll
    for (int i = 0; i < SMALL_M_MAX; ++i) acc[i] = (a4[i][0] + a4[i][1]) + (a4[i][2] + a4[i][3]);
  } else {
    for (int kk = lane; kk < k; kk += 32) {
      const float wv = wload<WT>(wr + kk);
#pragma unroll
      for (int i = 0; i < SMALL_M_MAX; ++i)
        if (i < m) acc[i] = fmaf(xs[i * k + kk], wv, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < SMALL_M_MAX; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  if (lane == 0) {
    const float bv = b ? wload<WT>(b + col) : 0.f;
    for (int i = 0; i < m; ++i) {
      float v = acc[i] + bv;
      if (round_out) v = __bfloat162float(__float2bfloat16_rn(v));
      if (act_out == 1) {
        v = silu_exact(v);
        if (round_out) v = __bfloat162float(__float2bfloat16_rn(v));
      }
      y[(int64_t)i * n + col] = v;
    }
  }
}

int linear_small_m(const float* x, const void* w, const void* b, float* y, int m, int n, int k, int w_is_bf16,
                   int act_in, int act_out, int round_in, int round_out, cudaStream_t stream) {
  FINO_CHECK_ARG(x && w && y, "linear_small_m: null pointer");
  FINO_CHECK_ARG(m > 0 && m <= SMALL_M_MAX * 65535, "linear_small_m: m=%d out of range", m);
  FINO_CHECK_ARG(n > 0 && k > 0, "linear_small_m: bad shape");
  const size_t smem = (size_t)(m < SMALL_M_MAX ? m : SMALL_M_MAX) * k * sizeof(float);
  FINO_CHECK_ARG(smem <= 200 * 1024, "linear_small_m: m*k too large for shared memory");
  const int warps = 8;
  dim3 grid((n + warps - 1) / warps, (m + SMALL_M_MAX - 1) / SMALL_M_MAX);
  const bool vec = (k % 128 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
#define FINO_SMALL_M(WT_, VEC_)                                                                                       \
  do {                                                                                                                \
    FINO_CHECK_CUDA(cudaFuncSetAttribute(linear_small_m_kernel<WT_, VEC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         200 * 1024));                                                               \
    linear_small_m_kernel<WT_, VEC_><<<grid, warps * 32, smem, stream>>>(x, (const WT_*)w, (const WT_*)b, y, m, n, k,   \
                                                                        act_in, act_out, round_in, round_out);        \
  } while (0)
  if (w_is_bf16) {
    if (vec) FINO_SMALL_M(__nv_bfloat16, true);
    else FINO_SMALL_M(__nv_bfloat16, false);
  } else {
    if (vec) FINO_SMALL_M(float, true);
    else FINO_SMALL_M(float, false);
  }
#undef FINO_SMALL_M
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// De-duplication of the per-token timesteps on the device (replaces the torch.unique radix sort + host sync at the top
// of every forward): the FrameINO sampler passes two distinct values (0 on the clean first frame, t elsewhere;
// pipeline_wan_i2v_motion_FrameINO.py:832-843), the general case is capped at 8.
//   uniq[0..8)   the distinct values in ascending order (torch.unique order), unused slots repeat the largest
//   row_index[i] the position of t[i] in uniq
//   count[0]     the number of distinct values; 9 = more than 8 were seen (uniq/row_index are then unusable)
// One 1024-thread block: a block-wide membership test per 1024-element chunk; a chunk that holds an unseen value adds
// ONE value (that of the lowest thread that holds one) and re-tests, so the list grows in first-appearance order.
// ------------------------------------------------------------------------------------------------
constexpr int DEDUP_MAX = 8;

__global__ void __launch_bounds__(1024)
timestep_dedup_kernel(const float* __restrict__ t, int64_t n, float* __restrict__ uniq, int32_t* __restrict__ row_index,
                      int32_t* __restrict__ count) {
  __shared__ float s_list[DEDUP_MAX];
  __shared__ int s_n, s_pick, s_over;
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_n = 0;
    s_pick = 0x7fffffff;
    s_over = 0;
  }
  __syncthreads();
  for (int64_t base = 0; base < n && !s_over; base += blockDim.x) {
    const int64_t i = base + tid;
    const bool has = i < n;
    const float x = has ? t[i] : 0.f;
    while (true) {
      const int cnt = s_n;
      bool found = !has;
      for (int j = 0; j < cnt; ++j) found |= (s_list[j] == x);
      if (!__syncthreads_or(!found)) break;
      if (!found) atomicMin(&s_pick, tid);
      __syncthreads();
      if (tid == s_pick) {
        if (s_n < DEDUP_MAX) s_list[s_n++] = x;
        else s_over = 1;
        s_pick = 0x7fffffff;
      }
      __syncthreads();
      if (s_over) break;
    }
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) {
    const int cnt = s_n;
    for (int a = 1; a < cnt; ++a) {  // insertion sort, ascending
      const float v = s_list[a];
      int b = a - 1;
      while (b >= 0 && s_list[b] > v) {
        s_list[b + 1] = s_list[b];
        --b;
      }
      s_list[b + 1] = v;
    }
    for (int j = 0; j < DEDUP_MAX; ++j) uniq[j] = cnt > 0 ? s_list[j < cnt ? j : cnt - 1] : 0.f;
    count[0] = s_over ? DEDUP_MAX + 1 : cnt;
  }
  __syncthreads();
  const int cnt = s_n;
  for (int64_t i = tid; i < n; i += blockDim.x) {
    const float x = t[i];
    int idx = 0;
    for (int j = 0; j < cnt; ++j)
      if (s_list[j] == x) idx = j;
    row_index[i] = idx;
  }
}

int timestep_dedup(const float* t, int64_t n, float* uniq, int32_t* row_index, int32_t* count, cudaStream_t stream) {
  FINO_CHECK_ARG(t && uniq && row_index && count && n > 0, "timestep_dedup: bad arguments");
  timestep_dedup_kernel<<<1, 1024, 0, stream>>>(t, n, uniq, row_index, count);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// out[l, r, c] = table[l, c] + proj[r, c]      (fp32; l layers, r unique timesteps, c = chunks*dim columns)
// ------------------------------------------------------------------------------------------------
__global__ void build_mod_table_kernel(const float* __restrict__ table, const float* __restrict__ proj,
                                       float* __restrict__ out, int layers, int r, int cols, int64_t table_layer_stride) {
  const int64_t total = (int64_t)layers * r * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cols);
    const int64_t t = idx / cols;
    const int ri = (int)(t % r);
    const int l = (int)(t / r);
    out[idx] = table[(int64_t)l * table_layer_stride + c] + proj[(int64_t)ri * cols + c];
  }
}

int build_mod_table(const float* table, const float* proj, float* out, int layers, int r, int cols,
                    int64_t table_layer_stride, cudaStream_t stream) {
  FINO_CHECK_ARG(table && proj && out && layers > 0 && r > 0 && cols > 0, "build_mod_table: bad arguments");
  const int64_t total = (int64_t)layers * r * cols;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  build_mod_table_kernel<<<(unsigned)blocks, 256, 0, stream>>>(table, proj, out, layers, r, cols, table_layer_stride);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// out[b][a][:] = in[a][b][:]   (16-byte vectors; `inner` bf16 elements per row, multiple of 8)
// The (de)interleave step on either side of the Ulysses all-to-all.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
swap01_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t A, int64_t B, int vec_per_row) {
  const int64_t total = A * B * vec_per_row;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    // idx enumerates the OUTPUT [b][a][v]
    const int v = (int)(idx % vec_per_row);
    const int64_t t = idx / vec_per_row;
    const int64_t a = t % A;
    const int64_t b = t / A;
    out[idx] = in[(a * B + b) * vec_per_row + v];
  }
}

int swap01(const void* in, void* out, int64_t A, int64_t B, int64_t inner, cudaStream_t stream) {
  FINO_CHECK_ARG(in && out && A > 0 && B > 0 && inner > 0 && inner % 8 == 0, "swap01: bad arguments");
  const int64_t total = A * B * (inner / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  swap01_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const uint4*)in, (uint4*)out, A, B, (int)(inner / 8));
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

}  // namespace fino
