// Causal 3-D / 2-D convolution over channels-last activations as an IMPLICIT GEMM on the tcgen05 tensor cores
// (Wan VAE: reference architecture/autoencoder_kl_wan.py:134-176 WanCausalConv3d, :245-262 the Conv2d of WanResample).
//
//   Y[t, h, w, :] = bias + sum_{kt,kh,kw} W[:, kt, kh, kw, :] . X[t*st + kt, h*s + kh - ph, w*s + kw - pw, :]
//
// X is [T_in, H_in, W_in, C_in] bf16 (channels last), W is [C_out, KT*KH*KW * C_inP] bf16 (tap-major, channels padded
// to a multiple of 64 per tap), the accumulator is fp32. The convolution is "valid" along time: the caller prepends
// the causal history (two cached frames, zeros for the first chunk), exactly the reference's feat_cache mechanism
// (:169-176). Spatial zero padding costs nothing: TMA zero-fills out-of-bounds box elements, negative coordinates
// included, so no im2col buffer and no padded copy of the activations ever exists.
//
// Mapping onto the GEMM of gemm_tcgen05.cu (same warp roles, same 128 x BN tile, same fused epilogue):
//   M tile = 128 output pixels of ONE frame = an 8 x 16 (h x w) patch;   N tile = BN output channels
//   K loop = taps x 64-channel chunks: per step the producer issues ONE 4-D TMA box load {64 ch, 16 w, 8 h, 1 t} at the
//            tap's offset (element stride 2 along h / w for the stride-2 downsampling convs) and one 2-D load of the
//            matching 64-column slice of W. The box lands in shared memory as 128 rows x 128 B, 128-byte swizzled —
//            byte for byte the K-major A tile tcgen05.mma expects.
//   epilogue: accumulator row r -> pixel (h0 + r/16, w0 + r%16); rows outside the frame are masked.
#include "gemm_common.cuh"

namespace fino {

constexpr int CONV_TH = 8;   // tile height (pixels)
constexpr int CONV_TW = 16;  // tile width

struct ConvGeom {
  int T, Ho, Wo;   // output frames / height / width
  int nhb, nwb;    // tiles along h and w
  int KT, KH, KW;
  int pad_h, pad_w;       // leading spatial zero padding
  int stride_hw, stride_t;
  int cchunks;            // 64-channel chunks per tap
  int64_t row_t, row_h, row0;  // "virtual row" of output pixel (t, h, w) = row0 + t*row_t + h*row_h + w: the epilogue
                               // addresses C / residual as row * ldc, ldc = pixel stride (frames may be interleaved)
};

template <int BN>
struct ConvCfg {
  static constexpr int kStages = (BN > 128) ? 4 : 6;
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;
  static constexpr int kBBytes = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv_cl_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmParams p, const ConvGeom g) {
  using Cfg = ConvCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t bars = smem_base + kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * kStages + 2 + s); };
  uint32_t tmem_ptr_smem = bars + 8u * (2 * kStages + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int taps = g.KT * g.KH * g.KW;
  const int num_kb = taps * g.cchunks;
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
        const int wb = mt % g.nwb;
        const int hb = (mt / g.nwb) % g.nhb;
        const int t = mt / (g.nwb * g.nhb);
        const int w_in0 = wb * CONV_TW * g.stride_hw - g.pad_w;
        const int h_in0 = hb * CONV_TH * g.stride_hw - g.pad_h;
        const int t_in0 = t * g.stride_t;
        int kb = 0;
        for (int kt = 0; kt < g.KT; ++kt)
          for (int kh = 0; kh < g.KH; ++kh)
            for (int kw = 0; kw < g.KW; ++kw)
              for (int cc = 0; cc < g.cchunks; ++cc, ++kb) {
                mbar_wait_relaxed(empty_bar(stage), phase ^ 1u, 100 + stage);
                const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
                const uint32_t b_dst = a_dst + Cfg::kABytes;
                mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
                tma_load_4d(a_dst, &tmap_a, full_bar(stage), cc * GEMM_BK, w_in0 + kw, h_in0 + kh, t_in0 + kt);
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
          const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t adesc = make_sdesc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = make_sdesc_sw128(b_addr + k * 32, 16, 1024);
            umma_ss_w(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit_w(empty_bar(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit_w(tfull_bar(acc));
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int mt, nt;
      tile_coords(tile, p.num_m_tiles, p.num_n_tiles, mt, nt);
      const int wb = mt % g.nwb;
      const int hb = (mt / g.nwb) % g.nhb;
      const int t = mt / (g.nwb * g.nhb);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tfull_bar(acc), acc_phase, 400 + acc);
      tc_fence_after();
      const int r = q * 32 + lane;
      const int h = hb * CONV_TH + (r >> 4);
      const int w = wb * CONV_TW + (r & 15);
      const bool row_ok = h < g.Ho && w < g.Wo;
      const int64_t row = g.row0 + (int64_t)t * g.row_t + (int64_t)h * g.row_h + w;
      const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < (BN + 31) / 32; ++c) {
        const int col0 = nt * BN + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t rr[32];
        ResidualChunk rc;
        tmem_ld_32x32b_x32(taddr_row + c * 32, rr);
        load_residual_chunk<EPI>(p, row, col0, row_ok, nullptr, rc);
        tmem_wait_ld();
        if (row_ok) epilogue_chunk<EPI>(p, row, col0, nullptr, rr, p.bias ? p.bias + col0 : nullptr, rc);
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

template <int BN, int EPI>
static int launch_conv(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, const ConvGeom& g,
                       cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  FINO_CHECK_CUDA(cudaFuncSetAttribute(conv_cl_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));  // per launch: cheap, and correct on every device
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv_cl_kernel<BN, EPI><<<grid, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(ta, tb, p, g);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

template <int EPI>
static int launch_conv_bn(int BN, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, const ConvGeom& g,
                          cudaStream_t stream) {
  if (BN == 256) return launch_conv<256, EPI>(ta, tb, p, g, stream);
  if (BN == 160) return launch_conv<160, EPI>(ta, tb, p, g, stream);
  if (BN == 128) return launch_conv<128, EPI>(ta, tb, p, g, stream);
  return launch_conv<32, EPI>(ta, tb, p, g, stream);
}

// epilogue: 0 = bias only, 3 = residual + bf16(conv + bias) (EPI_GATE_RESIDUAL without a gate)
int conv3d_cl(const void* x, int t_in, int h_in, int w_in, int c_in, int64_t in_st, int64_t in_sh, int64_t in_sw,
              const void* w, int64_t ldw, const void* bias, void* y, int t_out, int h_out, int w_out, int c_out,
              int64_t out_st, int64_t out_sh, int64_t out_sw, int kt, int kh, int kw, int pad_h, int pad_w,
              int stride_hw, int stride_t, const void* residual, int epilogue, cudaStream_t stream) {
  FINO_CHECK_ARG(x && w && y, "conv3d_cl: null pointer");
  FINO_CHECK_ARG(t_in > 0 && h_in > 0 && w_in > 0 && c_in > 0 && t_out > 0 && h_out > 0 && w_out > 0 && c_out > 0,
                 "conv3d_cl: non-positive shape");
  FINO_CHECK_ARG(kt >= 1 && kt <= 3 && kh >= 1 && kh <= 3 && kw >= 1 && kw <= 3, "conv3d_cl: kernel size 1..3");
  FINO_CHECK_ARG((stride_hw == 1 || stride_hw == 2) && (stride_t == 1 || stride_t == 2), "conv3d_cl: stride 1 or 2");
  FINO_CHECK_ARG(c_in % 8 == 0 && c_out % 8 == 0, "conv3d_cl: channel counts must be multiples of 8");
  FINO_CHECK_ARG(in_sw % 8 == 0 && in_sh % 8 == 0 && in_st % 8 == 0, "conv3d_cl: input strides must be multiples of 8");
  FINO_CHECK_ARG(out_sw % 8 == 0 && out_sh % out_sw == 0 && out_st % out_sw == 0,
                 "conv3d_cl: output row / frame strides must be multiples of the pixel stride (a multiple of 8)");
  FINO_CHECK_ARG((t_out - 1) * stride_t + kt <= t_in, "conv3d_cl: the convolution is valid along time: need "
                 "(t_out-1)*stride_t + kt <= t_in (prepend the causal history frames)");
  FINO_CHECK_ARG(epilogue == EPI_NONE || (epilogue == EPI_GATE_RESIDUAL && residual != nullptr),
                 "conv3d_cl: epilogue 0 (bias) or 3 (residual add, needs a residual pointer)");
  const int cchunks = (c_in + GEMM_BK - 1) / GEMM_BK;
  const int taps = kt * kh * kw;
  FINO_CHECK_ARG(ldw >= (int64_t)taps * cchunks * GEMM_BK && ldw % 8 == 0,
                 "conv3d_cl: weight rows hold taps * ceil(c_in/64)*64 = %d columns, ldw = %lld", taps * cchunks * GEMM_BK,
                 (long long)ldw);
  FINO_CHECK_ARG(((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual) |
                   reinterpret_cast<uintptr_t>(bias)) & 15) == 0, "conv3d_cl: y / residual / bias must be 16-byte aligned");
  // N tile: 160 for the encoder widths (160 / 320 / 640 = whole tiles; 256-wide tiles would idle 37 % / 17 % of the MMA)
  const int BN = (c_out % 160 == 0 && c_out % 256 != 0) ? 160 : c_out > 128 ? 256 : (c_out > 32 ? 128 : 32);
  ConvGeom g;
  g.T = t_out, g.Ho = h_out, g.Wo = w_out;
  g.nhb = (h_out + CONV_TH - 1) / CONV_TH;
  g.nwb = (w_out + CONV_TW - 1) / CONV_TW;
  g.KT = kt, g.KH = kh, g.KW = kw;
  g.pad_h = pad_h, g.pad_w = pad_w;
  g.stride_hw = stride_hw, g.stride_t = stride_t;
  g.cchunks = cchunks;
  g.row_t = out_st / out_sw;
  g.row_h = out_sh / out_sw;
  g.row0 = 0;
  const int64_t m_tiles = (int64_t)t_out * g.nhb * g.nwb;
  FINO_CHECK_ARG(m_tiles * ((c_out + BN - 1) / BN) < ((int64_t)1 << 30), "conv3d_cl: too many tiles");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = m_tiles * GEMM_BM;
  p.N = c_out;
  p.K = taps * cchunks * GEMM_BK;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.C = y;
  p.ldc = out_sw;
  p.out_fp32 = 0;
  p.epilogue = epilogue;
  p.flags = 0;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.ldr = out_sw;
  p.gate = nullptr;
  p.rows_per_group = (int64_t)1 << 62;
  p.num_m_tiles = (int)m_tiles;
  p.num_n_tiles = (c_out + BN - 1) / BN;
  p.num_full = p.num_m_tiles * p.num_n_tiles;
  p.splits = 1;
  CUtensorMap ta, tb;
  {
    uint64_t dims[4] = {(uint64_t)c_in, (uint64_t)w_in, (uint64_t)h_in, (uint64_t)t_in};
    uint64_t strides[3] = {(uint64_t)in_sw * 2, (uint64_t)in_sh * 2, (uint64_t)in_st * 2};
    uint32_t box[4] = {GEMM_BK, (uint32_t)(CONV_TW * stride_hw), (uint32_t)(CONV_TH * stride_hw), 1};
    uint32_t estr[4] = {1, (uint32_t)stride_hw, (uint32_t)stride_hw, 1};
    int r = encode_tmap_bf16(&ta, x, 4, dims, strides, box, estr);
    if (r) return r;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)c_out};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {GEMM_BK, (uint32_t)BN};
    int r = encode_tmap_bf16(&tb, w, 2, dims, strides, box);
    if (r) return r;
  }
  if (epilogue == EPI_GATE_RESIDUAL) return launch_conv_bn<EPI_GATE_RESIDUAL>(BN, ta, tb, p, g, stream);
  return launch_conv_bn<EPI_NONE>(BN, ta, tb, p, g, stream);
}

}  // namespace fino
