// Denoise-loop glue of the Wan2.2 FrameINO sampler, fused into one kernel before and one kernel after the two
// transformer forwards of a scheduler step. Both are HBM-bound gathers over ~11 MB tensors.
//
//   wan_pack_model_input   reference pipelines/pipeline_wan_i2v_motion_FrameINO.py:829-830 (first-frame mask blend),
//                          :854 (frame-wise ID concat), :858 (channel-wise trajectory concat + cast), followed by the
//                          patchify of transformer_wan.py:486-487 -> GEMM rows, without materialising the 5-D input
//   wan_cfg_euler_step     reference :882 (classifier-free guidance), :886 (drop the ID frames), :891 (scheduler step,
//                          flow-match Euler x += (sigma_next - sigma) * v), reading the two forwards' proj_out rows
//                          (transformer_wan.py:537) directly, i.e. fused with the un-patchify of :539-543
//
// fp32 arithmetic uses the explicit round-to-nearest intrinsics so that no FMA contraction happens: results are bit
// identical to the reference's separate fp32 tensor ops.
#include "common.cuh"

namespace fino {

struct LoopGeom {
  int B, C, F, NID, H, W;  // latent batch, channels, generated frames, ID frames, height, width
  int pt, ph, pw;
};

// rows[(b, fq, hq, wq), (c2, it, ih, iw)], c2 in [0, 2C): c2 < C -> blended latent / ID latent, else trajectory latent
__global__ void __launch_bounds__(256)
wan_pack_model_input_kernel(const float* __restrict__ latents, const float* __restrict__ condition,
                            const float* __restrict__ mask, const float* __restrict__ id_latents,
                            const float* __restrict__ traj, __nv_bfloat16* __restrict__ rows, const LoopGeom g,
                            int64_t ld, int64_t total) {
  const int FT = g.F + g.NID;
  const int pf = FT / g.pt, phh = g.H / g.ph, pww = g.W / g.pw;
  const int kdim = 2 * g.C * g.pt * g.ph * g.pw;
  const int64_t hw = (int64_t)g.H * g.W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % kdim);
    const int64_t row = idx / kdim;
    int64_t tok = row;
    const int wq = (int)(tok % pww);
    tok /= pww;
    const int hq = (int)(tok % phh);
    tok /= phh;
    const int fq = (int)(tok % pf);
    const int b = (int)(tok / pf);
    int kk = k;
    const int iw = kk % g.pw;
    kk /= g.pw;
    const int ih = kk % g.ph;
    kk /= g.ph;
    const int it = kk % g.pt;
    const int c2 = kk / g.pt;
    const int f = fq * g.pt + it, h = hq * g.ph + ih, w = wq * g.pw + iw;
    const int64_t sp = (int64_t)h * g.W + w;
    float v;
    if (c2 >= g.C) {
      v = traj[(((int64_t)b * g.C + (c2 - g.C)) * FT + f) * hw + sp];
    } else if (f >= g.F) {
      v = id_latents[(((int64_t)b * g.C + c2) * g.NID + (f - g.F)) * hw + sp];
    } else {
      const int64_t o = (((int64_t)b * g.C + c2) * g.F + f) * hw + sp;
      const float m = mask[(int64_t)f * hw + sp];
      v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, m), condition[o]), __fmul_rn(m, latents[o]));
    }
    rows[row * ld + k] = __float2bfloat16_rn(v);
  }
}

int wan_pack_model_input(const float* latents, const float* condition, const float* mask, const float* id_latents,
                         const float* traj, void* rows, int B, int C, int F, int NID, int H, int W, int pt, int ph,
                         int pw, int64_t ld, cudaStream_t stream) {
  FINO_CHECK_ARG(latents && condition && mask && traj && rows, "wan_pack_model_input: null pointer");
  FINO_CHECK_ARG(NID == 0 || id_latents, "wan_pack_model_input: id_latents is null but n_id > 0");
  FINO_CHECK_ARG(B > 0 && C > 0 && F > 0 && NID >= 0 && H > 0 && W > 0 && pt > 0 && ph > 0 && pw > 0,
                 "wan_pack_model_input: bad shape");
  FINO_CHECK_ARG((F + NID) % pt == 0 && H % ph == 0 && W % pw == 0,
                 "wan_pack_model_input: dims not divisible by the patch size");
  const int64_t kdim = (int64_t)2 * C * pt * ph * pw;
  FINO_CHECK_ARG(ld >= kdim, "wan_pack_model_input: row stride %lld < %lld columns", (long long)ld, (long long)kdim);
  LoopGeom g{B, C, F, NID, H, W, pt, ph, pw};
  const int64_t total = (int64_t)B * ((F + NID) / pt) * (H / ph) * (W / pw) * kdim;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  wan_pack_model_input_kernel<<<(unsigned)blocks, 256, 0, stream>>>(latents, condition, mask, id_latents, traj,
                                                                    (__nv_bfloat16*)rows, g, ld, total);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// latents[b,c,f,h,w] += dsigma * cfg, cfg = u + guidance * (v - u) evaluated IN BF16 like the reference (the pipeline
// combines the two bf16 forward outputs with tensor ops, pipeline_wan_i2v_motion_FrameINO.py:882: three roundings —
// the difference, its product with the scalar, the sum); v/u = proj_out rows of the cond / uncond forwards with the
// (pt, ph, pw, c) column order of transformer_wan.py:539-543; rows of the ID frames (f >= F) are never read.
__global__ void __launch_bounds__(256)
wan_cfg_euler_step_kernel(const __nv_bfloat16* __restrict__ y_cond, const __nv_bfloat16* __restrict__ y_uncond,
                          float* __restrict__ latents, const LoopGeom g, int64_t ld, float guidance, float dsigma,
                          int64_t total) {
  const int FT = g.F + g.NID;
  const int pf = FT / g.pt, phh = g.H / g.ph, pww = g.W / g.pw;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx;
    const int w = (int)(r % g.W);
    r /= g.W;
    const int h = (int)(r % g.H);
    r /= g.H;
    const int f = (int)(r % g.F);
    r /= g.F;
    const int c = (int)(r % g.C);
    const int b = (int)(r / g.C);
    const int fq = f / g.pt, it = f % g.pt, hq = h / g.ph, ih = h % g.ph, wq = w / g.pw, iw = w % g.pw;
    const int64_t tok = (((int64_t)b * pf + fq) * phh + hq) * pww + wq;
    const int k = ((it * g.ph + ih) * g.pw + iw) * g.C + c;
    float v = __bfloat162float(y_cond[tok * ld + k]);
    if (y_uncond != nullptr) {
      const float u = __bfloat162float(y_uncond[tok * ld + k]);
      const float d = __bfloat162float(__float2bfloat16_rn(__fsub_rn(v, u)));
      const float m = __bfloat162float(__float2bfloat16_rn(__fmul_rn(guidance, d)));
      v = __bfloat162float(__float2bfloat16_rn(__fadd_rn(u, m)));
    }
    latents[idx] = __fadd_rn(latents[idx], __fmul_rn(dsigma, v));
  }
}

int wan_cfg_euler_step(const void* y_cond, const void* y_uncond, int64_t ld, float* latents, int B, int C, int F,
                       int NID, int H, int W, int pt, int ph, int pw, float guidance, float dsigma,
                       cudaStream_t stream) {
  FINO_CHECK_ARG(y_cond && latents, "wan_cfg_euler_step: null pointer");
  FINO_CHECK_ARG(B > 0 && C > 0 && F > 0 && NID >= 0 && H > 0 && W > 0 && pt > 0 && ph > 0 && pw > 0,
                 "wan_cfg_euler_step: bad shape");
  FINO_CHECK_ARG((F + NID) % pt == 0 && F % pt == 0 && H % ph == 0 && W % pw == 0,
                 "wan_cfg_euler_step: dims not divisible by the patch size");
  FINO_CHECK_ARG(ld >= (int64_t)C * pt * ph * pw, "wan_cfg_euler_step: row stride too small");
  LoopGeom g{B, C, F, NID, H, W, pt, ph, pw};
  const int64_t total = (int64_t)B * C * F * H * W;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  wan_cfg_euler_step_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
      (const __nv_bfloat16*)y_cond, (const __nv_bfloat16*)y_uncond, latents, g, ld, guidance, dsigma, total);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

}  // namespace fino
