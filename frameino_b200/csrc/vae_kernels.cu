// HBM-bound helpers of the Wan VAE (reference architecture/autoencoder_kl_wan.py), all on channels-last bf16
// activations [T, H, W, C] (a pixel = one contiguous row of C channels):
//
//   rms_act_cl         WanRMS_norm (:179-202: F.normalize over channels * sqrt(C) * gamma [+ bias]) fused with the SiLU
//                      that follows it in every residual block / head (:347-348, :363-364, :603-604, :897-898)
//   upsample2x_cl      WanUpsample(scale 2, "nearest-exact") (:205-217, :245-252)
//   dupup_add_cl       y += DupUp3D(x_copy) (:90-131, :709-710): the parameter-free shortcut of WanResidualUpBlock
//   avgdown_add_cl     y += AvgDown3D(x_copy) (:37-87, :502): the shortcut of WanResidualDownBlock
//   softmax_rows       softmax(S * scale) rows of the single-head attention of WanAttentionBlock (:402-414), fp32 -> bf16
//   vae_to_cl          [C, T, H, W] (fp32 | bf16, any strides) -> channels-last bf16, optional patchify(2) (:912-932),
//                      channels zero-padded to the stride
//   vae_from_cl        channels-last bf16 -> [C, T, H, W] fp32 | bf16, optional unpatchify(2) (:935-952) + clamp (:1224)
#include "common.cuh"
#include "ptx.cuh"

namespace fino {

namespace {
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pk2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack8f(const uint4& u, float* f) {
  f[0] = bf_lo(u.x), f[1] = bf_hi(u.x), f[2] = bf_lo(u.y), f[3] = bf_hi(u.y);
  f[4] = bf_lo(u.z), f[5] = bf_hi(u.z), f[6] = bf_lo(u.w), f[7] = bf_hi(u.w);
}
__device__ __forceinline__ uint4 pack8f(const float* f) {
  return make_uint4(pk2(f[0], f[1]), pk2(f[2], f[3]), pk2(f[4], f[5]), pk2(f[6], f[7]));
}
// streaming 16-byte load (read once: do not allocate in L1)
__device__ __forceinline__ uint4 ld_stream16(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
}  // namespace

// These kernels are HBM-bound with one pass over the data, so what sets their speed is the number of bytes each SM keeps
// in flight: ~6.5 TB/s x ~1.5 us of loaded-DRAM latency / 148 SMs = ~66 KB per SM. One 16-byte load per thread at full
// occupancy is 32 KB (measured: 3.0-3.5 TB/s); every kernel below therefore issues 4-8 independent 16-byte loads per
// thread before it consumes the first one.

// ------------------------------------------------------------------------------------------------
// out[r, :] = act( x[r, :] / max(||x[r, :]||_2, 1e-12) * sqrt(C) * gamma + bias ),  act = SiLU or identity.
// Warp per pixel row, C <= 1024 (multiple of 8): lane l holds the 16-byte chunks l, l+32, l+64, l+96.
// ------------------------------------------------------------------------------------------------
template <int CPL, int RPW>  // 16-byte chunks per lane and row; rows per warp (narrow rows: more bytes in flight per warp)
__global__ void __launch_bounds__(256, 3) rms_act_cl_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                        int64_t rows, int C, int64_t xs, int64_t os,
                                                        const float* __restrict__ gamma, const float* __restrict__ bias,
                                                        float scale, int silu) {
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW;
  if (row0 >= rows) return;
  const int nch = C >> 3;
  uint4 raw[RPW][CPL];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (row0 + r) * xs);
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = i * 32 + lane;
      raw[r][i] = (c < nch && row0 + r < rows) ? ld_stream16(xr + c) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  float sq[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    sq[r] = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      float v[8];
      unpack8f(raw[r][i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) sq[r] = fmaf(v[e], v[e], sq[r]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < RPW; ++r) sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], o);
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = i * 32 + lane;
    if (c >= nch) continue;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c);
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c + 1);
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float bs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * c);
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * c + 1);
      bs[0] = b0.x, bs[1] = b0.y, bs[2] = b0.z, bs[3] = b0.w, bs[4] = b1.x, bs[5] = b1.y, bs[6] = b1.z, bs[7] = b1.w;
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      if (row0 + r >= rows) continue;
      const float inv = scale / fmaxf(sqrtf(sq[r]), 1e-12f);  // F.normalize: x / max(||x||, eps), times sqrt(C)
      float v[8], o[8];
      unpack8f(raw[r][i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float y = fmaf(v[e] * inv, gm[e], bs[e]);
        if (silu) {  // y * sigmoid(y) = y / (1 + 2^(-y log2 e)): two MUFU ops (ex2, rcp) and three FP32 ops per element
          float ex, rc;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-1.4426950408889634f * y));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(1.0f + ex));
          y *= rc;
        }
        o[e] = y;
      }
      reinterpret_cast<uint4*>(out + (row0 + r) * os)[c] = pack8f(o);
    }
  }
}

int rms_act_cl(const void* x, void* out, int64_t rows, int C, int64_t x_stride, int64_t out_stride, const float* gamma,
               const float* bias, int silu, cudaStream_t stream) {
  FINO_CHECK_ARG(x && out && gamma && rows > 0, "rms_act_cl: bad arguments");
  FINO_CHECK_ARG(C >= 8 && C <= 1024 && C % 8 == 0, "rms_act_cl: C=%d (multiple of 8, <= 1024)", C);
  FINO_CHECK_ARG(x_stride % 8 == 0 && out_stride % 8 == 0, "rms_act_cl: strides must be multiples of 8");
  const float scale = sqrtf((float)C);
  const __nv_bfloat16* xi = (const __nv_bfloat16*)x;
  __nv_bfloat16* oo = (__nv_bfloat16*)out;
  const int cpl = (C / 8 + 31) / 32;
#define FINO_RMS(CPL_, RPW_)                                                                                    \
  rms_act_cl_kernel<CPL_, RPW_><<<(unsigned)((rows + 8 * RPW_ - 1) / (8 * RPW_)), 256, 0, stream>>>(             \
      xi, oo, rows, C, x_stride, out_stride, gamma, bias, scale, silu)
  if (cpl <= 1) FINO_RMS(1, 4);
  else if (cpl <= 2) FINO_RMS(2, 2);
  else if (cpl <= 3) FINO_RMS(3, 1);
  else FINO_RMS(4, 1);
#undef FINO_RMS
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// out[t, ho, wo, :] = in[t, ho / 2, wo / 2, :]   (nearest-exact, scale 2: floor((dst + 0.5) / 2) = dst / 2)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample2x_cl_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int T,
                                                           int H, int W, int vec) {
  // one thread = four input vectors (4 independent loads), each stored to its 2 x 2 output pixels
  const int64_t total = (int64_t)T * H * W * vec;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t idx = base + u * stride;
      if (idx < total) v[u] = ld_stream16(in + idx);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t idx = base + u * stride;
      if (idx >= total) continue;
      const int c = (int)(idx % vec);
      int64_t r = idx / vec;
      const int w = (int)(r % W);
      r /= W;
      const int h = (int)(r % H);
      const int t = (int)(r / H);
      uint4* o = out + (((int64_t)t * 2 * H + 2 * h) * 2 * W + 2 * w) * vec + c;
      o[0] = v[u];
      o[vec] = v[u];
      o[(int64_t)2 * W * vec] = v[u];
      o[(int64_t)2 * W * vec + vec] = v[u];
    }
  }
}

int upsample2x_cl(const void* in, void* out, int T, int H, int W, int C, cudaStream_t stream) {
  FINO_CHECK_ARG(in && out && T > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "upsample2x_cl: bad arguments");
  const int64_t total = (int64_t)T * H * W * (C / 8);
  int64_t blocks = (total + 4 * 256 - 1) / (4 * 256);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  upsample2x_cl_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const uint4*)in, (uint4*)out, T, H, W, C / 8);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// y[to, ho, wo, c] += src[ti, hi, wi, ((c*ft + a)*fs + b)*fs + d) / repeats]      (DupUp3D, :109-131)
//   global frame = to + t_drop; ti = frame / ft, a = frame % ft; hi = ho / fs, b = ho % fs; wi = wo / fs, d = wo % fs
//   t_drop = ft - 1 for the first chunk (its first ft - 1 duplicated frames are dropped, :129-130), else 0.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dupup_add_cl_kernel(__nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ src, int To, int Ho, int Wo, int Co,
                    int Hi, int Wi, int Ci, int ft, int fs, int repeats, int rshift, int t_drop) {
  const int groups = Co >> 3;  // 8 output channels (one 16-byte vector of y) per item, 4 items per thread
  const int64_t total = (int64_t)To * Ho * Wo * groups;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += 4 * stride) {
    uint4 yv[4];
    uint4* yp[4];
    const __nv_bfloat16* sp[4];
    int kb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t idx = base + u * stride;
      yp[u] = nullptr;
      if (idx >= total) continue;
      const int c0 = (int)(idx % groups) * 8;
      int64_t r = idx / groups;
      const int wo = (int)(r % Wo);
      r /= Wo;
      const int ho = (int)(r % Ho);
      const int to = (int)(r / Ho);
      const int fr = to + t_drop;
      const int ti = fr / ft, a = fr % ft, hi = ho / fs, b = ho % fs, wi = wo / fs, d = wo % fs;
      sp[u] = src + (((int64_t)ti * Hi + hi) * Wi + wi) * Ci;
      yp[u] = reinterpret_cast<uint4*>(y + (((int64_t)to * Ho + ho) * Wo + wo) * Co + c0);
      kb[u] = ((c0 * ft + a) * fs + b) * fs + d;  // folded channel of output channel c0; + e * ft*fs*fs for c0 + e
      yv[u] = *yp[u];
    }
    const int kstep = ft * fs * fs;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (yp[u] == nullptr) continue;
      float v[8];
      unpack8f(yv[u], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = kb[u] + e * kstep;
        v[e] += __bfloat162float(sp[u][rshift >= 0 ? (k >> rshift) : (k / repeats)]);
      }
      *yp[u] = pack8f(v);
    }
  }
}

int dupup_add_cl(void* y, const void* src, int To, int Ho, int Wo, int Co, int Ti, int Hi, int Wi, int Ci, int ft, int fs,
                 int t_drop, cudaStream_t stream) {
  FINO_CHECK_ARG(y && src && To > 0 && Ho > 0 && Wo > 0 && Co > 0 && Co % 8 == 0 && Ci > 0, "dupup_add_cl: bad arguments");
  FINO_CHECK_ARG(ft >= 1 && fs >= 1 && (Co * ft * fs * fs) % Ci == 0, "dupup_add_cl: out_channels*factor %% in_channels");
  FINO_CHECK_ARG(Ho == Hi * fs && Wo == Wi * fs && To + t_drop <= Ti * ft, "dupup_add_cl: shape mismatch");
  const int repeats = Co * ft * fs * fs / Ci;
  int rshift = -1;
  for (int sft = 0; sft < 16; ++sft)
    if ((1 << sft) == repeats) rshift = sft;
  const int64_t total = (int64_t)To * Ho * Wo * (Co / 8);
  int64_t blocks = (total + 4 * 256 - 1) / (4 * 256);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  dupup_add_cl_kernel<<<(unsigned)blocks, 256, 0, stream>>>((__nv_bfloat16*)y, (const __nv_bfloat16*)src, To, Ho, Wo, Co,
                                                           Hi, Wi, Ci, ft, fs, repeats, rshift, t_drop);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// y[to, ho, wo, co] += mean_g src_folded[co*group + g]   (AvgDown3D, :55-87) where the folded channel
//   k = (c*ft + a)*fs*fs + b*fs + d  of output pixel (to, ho, wo) is src[c, to*ft + a - t_pad, ho*fs + b, wo*fs + d]
//   (frames before the start, t_pad = (ft - T % ft) % ft of them, read as zero: F.pad front, :56-58).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
avgdown_add_cl_kernel(__nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ src, int To, int Ho, int Wo, int Co,
                      int Ti, int Hi, int Wi, int Ci, int ft, int fs, int group, int t_pad) {
  // one thread = 8 consecutive output channels of one output pixel (one 16-byte vector of y): the pixel coordinates are
  // decoded once, the folded-channel index k = co*group + g only needs small divisions by ft / fs
  const int groups8 = Co >> 3;
  const int64_t total = (int64_t)To * Ho * Wo * groups8;
  const float inv = 1.0f / (float)group;
  const int fss = fs * fs;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(idx % groups8) * 8;
    int64_t r = idx / groups8;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int to = (int)(r / Ho);
    uint4* yp = reinterpret_cast<uint4*>(y + (((int64_t)to * Ho + ho) * Wo + wo) * Co + c0);
    float v[8];
    unpack8f(*yp, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float acc = 0.f;
      for (int gidx = 0; gidx < group; ++gidx) {
        const int k = (c0 + e) * group + gidx;
        const int sp = k % fss;  // position inside the fs x fs block
        const int ca = k / fss;  // c * ft + a
        const int a = ca % ft, c = ca / ft;
        const int ti = to * ft + a - t_pad;
        if (ti >= 0)
          acc += __bfloat162float(src[(((int64_t)ti * Hi + ho * fs + sp / fs) * Wi + wo * fs + sp % fs) * Ci + c]);
      }
      v[e] += acc * inv;
    }
    *yp = pack8f(v);
  }
}

int avgdown_add_cl(void* y, const void* src, int To, int Ho, int Wo, int Co, int Ti, int Hi, int Wi, int Ci, int ft, int fs,
                   cudaStream_t stream) {
  FINO_CHECK_ARG(y && src && To > 0 && Ho > 0 && Wo > 0 && Co > 0 && Co % 8 == 0 && Ci > 0, "avgdown_add_cl: bad arguments");
  FINO_CHECK_ARG(ft >= 1 && fs >= 1 && (Ci * ft * fs * fs) % Co == 0, "avgdown_add_cl: in_channels*factor %% out_channels");
  const int t_pad = (ft - Ti % ft) % ft;
  FINO_CHECK_ARG(Hi == Ho * fs && Wi == Wo * fs && (Ti + t_pad) == To * ft, "avgdown_add_cl: shape mismatch");
  const int group = Ci * ft * fs * fs / Co;
  const int64_t total = (int64_t)To * Ho * Wo * (Co / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  avgdown_add_cl_kernel<<<(unsigned)blocks, 256, 0, stream>>>((__nv_bfloat16*)y, (const __nv_bfloat16*)src, To, Ho, Wo, Co,
                                                             Ti, Hi, Wi, Ci, ft, fs, group, t_pad);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// p[r, c] = softmax_c(s[r, :] * scale)   fp32 in (row stride ls), bf16 out (row stride lp); columns [cols, lp) -> 0.
// One CTA of 256 threads per row (cols <= 16384: the row stays in shared memory).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p,
                                                          int cols, int64_t ls, int64_t lp, float scale_log2) {
  extern __shared__ float row_s[];
  __shared__ float red[8];
  const float* sr = s + (int64_t)blockIdx.x * ls;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) {
    const float v = sr[c];
    row_s[c] = v;
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) {
    const float e = exp2f((row_s[c] - mx) * scale_log2);
    row_s[c] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  __nv_bfloat16* pr = p + (int64_t)blockIdx.x * lp;
  for (int c = threadIdx.x; c < (int)lp; c += 256) pr[c] = __float2bfloat16_rn(c < cols ? row_s[c] * inv : 0.f);
}

int softmax_rows(const float* s, void* p, int64_t rows, int cols, int64_t ls, int64_t lp, float scale, cudaStream_t stream) {
  FINO_CHECK_ARG(s && p && rows > 0 && cols > 0 && cols <= 16384 && lp >= cols && ls >= cols, "softmax_rows: bad arguments");
  FINO_CHECK_ARG(rows < ((int64_t)1 << 31), "softmax_rows: too many rows");
  softmax_rows_kernel<<<(unsigned)rows, 256, (size_t)cols * sizeof(float), stream>>>(
      s, (__nv_bfloat16*)p, cols, ls, lp, scale * 1.4426950408889634f);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// [C, T, H, W] (element strides sc, st, sh, sw; fp32 or bf16) -> channels-last bf16 [T, H/ps, W/ps, cpad]:
//   channel c*ps*ps + (w % ps)*ps + (h % ps) of pixel (t, h / ps, w / ps) (patchify, :912-932); channels
//   [C*ps*ps, cpad) are zero.
// ------------------------------------------------------------------------------------------------
template <typename TI>
__global__ void __launch_bounds__(256) vae_to_cl_kernel(const TI* __restrict__ in, __nv_bfloat16* __restrict__ out, int C,
                                                       int T, int H, int W, int64_t sc, int64_t st, int64_t sh,
                                                       int64_t sw, int ps, int cpad) {
  const int Ho = H / ps, Wo = W / ps;
  const int64_t total = (int64_t)T * Ho * Wo * cpad;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % cpad);
    int64_t r = idx / cpad;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int t = (int)(r / Ho);
    float v = 0.f;
    if (k < C * ps * ps) {
      const int i = k % ps, j = (k / ps) % ps, c = k / (ps * ps);
      v = (float)in[(int64_t)c * sc + (int64_t)t * st + (int64_t)(ho * ps + i) * sh + (int64_t)(wo * ps + j) * sw];
    }
    out[idx] = __float2bfloat16_rn(v);
  }
}

int vae_to_cl(const void* in, int in_fp32, void* out, int C, int T, int H, int W, int64_t sc, int64_t st, int64_t sh,
              int64_t sw, int ps, int cpad, cudaStream_t stream) {
  FINO_CHECK_ARG(in && out && C > 0 && T > 0 && H > 0 && W > 0 && ps >= 1 && H % ps == 0 && W % ps == 0 &&
                     cpad >= C * ps * ps, "vae_to_cl: bad arguments");
  const int64_t total = (int64_t)T * (H / ps) * (W / ps) * cpad;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  if (in_fp32) vae_to_cl_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>((const float*)in, (__nv_bfloat16*)out, C, T, H, W, sc, st, sh, sw, ps, cpad);
  else vae_to_cl_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, C, T, H, W, sc, st, sh, sw, ps, cpad);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// channels-last bf16 [T, Hi, Wi, cstride] -> out[c, t, hi*ps + i, wi*ps + j] = in[t, hi, wi, c*ps*ps + j*ps + i]
// (unpatchify, :935-952), optionally clamped to [-1, 1] (:1224); out [C, T, Hi*ps, Wi*ps] with channel stride out_sc
// (frames, rows and pixels contiguous), fp32 or bf16.
// ------------------------------------------------------------------------------------------------
template <typename TO>
__global__ void __launch_bounds__(256) vae_from_cl_kernel(const __nv_bfloat16* __restrict__ in, TO* __restrict__ out, int C,
                                                         int T, int Hi, int Wi, int cstride, int ps, int clamp,
                                                         int64_t out_sc) {
  const int H = Hi * ps, W = Wi * ps;
  const int64_t total = (int64_t)C * T * H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(idx % W);
    int64_t r = idx / W;
    const int h = (int)(r % H);
    r /= H;
    const int t = (int)(r % T);
    const int c = (int)(r / T);
    const int k = c * ps * ps + (w % ps) * ps + (h % ps);
    float v = __bfloat162float(in[(((int64_t)t * Hi + h / ps) * Wi + w / ps) * cstride + k]);
    if (clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
    out[(int64_t)c * out_sc + ((int64_t)t * H + h) * W + w] = (TO)v;
  }
}

int vae_from_cl(const void* in, void* out, int out_fp32, int C, int T, int Hi, int Wi, int cstride, int ps, int clamp,
                int64_t out_sc, cudaStream_t stream) {
  FINO_CHECK_ARG(in && out && C > 0 && T > 0 && Hi > 0 && Wi > 0 && ps >= 1 && cstride >= C * ps * ps,
                 "vae_from_cl: bad arguments");
  const int64_t total = (int64_t)C * T * Hi * ps * Wi * ps;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  if (out_fp32) vae_from_cl_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)in, (float*)out, C, T, Hi, Wi, cstride, ps, clamp, out_sc);
  else vae_from_cl_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, C, T, Hi, Wi, cstride, ps, clamp, out_sc);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

}  // namespace fino
