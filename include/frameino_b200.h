/* frameino_b200 — C ABI of the B200-native FrameINO denoise-step kernels (sm_100a).
 *
 * The reference (UVA-Computer-Vision-Lab/FrameINO) has no FFI of its own: its hot path is PyTorch library calls
 * issued from architecture/transformer_wan.py, architecture/cogvideox_transformer_3d.py and
 * architecture/attention_processor.py. Each entry point below replaces one of those call sites (cited per
 * function) and is what a reference-side binding (ctypes) would bind.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; tensors are bf16 unless the type says float/int32_t;
 *   - strides are in ELEMENTS; `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 = ok, 1 = invalid argument, 2 = CUDA error, 3 = unsupported; the message is available from
 *     fino_last_error() (thread local). There is no CPU fallback: without a GPU every compute entry point fails.
 *   - "modulation rows": AdaLN tables are fp32 [R, row_stride]; the row used for token `row` is
 *     row_index[row] when row_index != NULL, else row / rows_per_group.
 */
#ifndef FRAMEINO_B200_H_
#define FRAMEINO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FINO_ABI_VERSION 1

/* status / bookkeeping ------------------------------------------------------------------------------------ */
int fino_abi_version(void);
const char* fino_last_error(void);
/* Binds the library's CUDA runtime to `device` (one process per GPU: call once with LOCAL_RANK). */
int fino_set_device(int device);
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
int64_t fino_launch_count(void);

/* epilogues of fino_gemm_bf16 */
#define FINO_EPI_NONE 0          /* C = A W^T + bias                                                         */
#define FINO_EPI_GELU_TANH 1     /* C = gelu_tanh(bf16(A W^T + bias))      FeedForward, transformer_wan.py:347 */
#define FINO_EPI_SILU 2          /* C = silu(bf16(A W^T + bias))                                             */
#define FINO_EPI_GATE_RESIDUAL 3 /* C = residual + bf16(A W^T + bias) * gate   transformer_wan.py:336,341,348 */
#define FINO_GEMM_FLAG_ROUND_PRODUCT 1 /* round gate*y to bf16 first (CogVideoX, cogvideox_transformer_3d.py:146) */

/* C[m,n] = epilogue(A[m,k] * W[n,k]^T + bias[n]); tcgen05/TMEM/TMA persistent GEMM.
 * Replaces nn.Linear on the hot path: transformer_wan.py:60-62,117,347,486,537; attention_processor.py:2837-2839,2870.
 * out_fp32: 0 -> C is bf16, 1 -> C is float. k, n, lda, ldw, ldc (and ldr) must be multiples of 8. */
int fino_gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* c, int64_t ldc,
                   int64_t m, int n, int k, int epilogue, int out_fp32, int flags, const void* residual, int64_t ldr,
                   const float* gate, int64_t gate_row_stride, const int32_t* row_index, int64_t rows_per_group,
                   void* stream);

/* Tuning / test hook: 0 = choose per shape (default), 1 = single-CTA kernel, 2 = CTA pair (cta_group::2) with
 * 256x256 cluster tiles, 3 = CTA pair with 512x256 cluster tiles. */
int fino_gemm_set_mode(int mode);

/* Work decomposition of the CTA-pair GEMM kernel (256 x 256 tiles on sms/2 CTA pairs, persistent). When the tile count
 * leaves a partly filled last round (8-way Ulysses: M = 3520, N = 3072 -> 168 tiles on 74 pairs = 2.27 rounds), the
 * tiles of that round are split along K into `splits` slices each (raw fp32 partials in a library-owned per-device
 * workspace; a small second kernel adds the slices and applies the fused epilogue). mode: -1 = automatic (default),
 * 0 = never, 2..16 = split the whole last round that many ways (test hook). The workspace is per (device, stream):
 * calls on different streams do not share it. */
int fino_gemm_set_split(int mode);
/* The decomposition fino_gemm_bf16 would use for an m x n x k problem on a device with `sms` SMs (host arithmetic). */
int fino_gemm_plan(int64_t m, int n, int k, int sms, int mode, int* num_full, int* splits);

/* O = softmax(Q K^T * scale) V, non-causal, no mask; tcgen05 flash attention, head_dim 64 or 128.
 * Replaces F.scaled_dot_product_attention: transformer_wan.py:108-110, attention_processor.py:2863.
 * Q/K/V/O are [batch, n, heads*head_dim] views (heads contiguous inside a row). */
int fino_attention_fwd(const void* q, const void* k, const void* v, void* o, int batch, int heads, int64_t nq,
                       int64_t nk, int head_dim, int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride,
                       int64_t o_row_stride, int64_t q_batch_stride, int64_t k_batch_stride, int64_t v_batch_stride,
                       int64_t o_batch_stride, float scale, void* stream);

/* Tuning / test hook: scheduling variant of the attention kernel (0 = default; 1..15 see attention_tcgen05.cu; 6..15
 * apply to head_dim 64 only). */
int fino_attention_set_variant(int variant);

/* Work decomposition of fino_attention_fwd. The kernel runs one CTA per 256-query-row tile of one (batch, head); when
 * the tile count leaves a partly filled last wave on the device's SMs (8-way Ulysses: 3 heads x 110 tiles = 330 CTAs on
 * 148 SMs), the tiles of that wave are split along the key axis into `splits` partial CTAs each (fp32 partials in a
 * library-owned per-device workspace, merged by a small second kernel). mode: -1 = automatic (default), 0 = never,
 * 2..64 = split EVERY tile that many ways (test hook). The workspace is per (device, stream). */
int fino_attention_set_split(int mode);
/* The decomposition fino_attention_fwd would use on a device with `sms` SMs (pure host arithmetic; no GPU needed):
 * CTAs [0, n_full) run whole tiles, the remaining tiles run as `splits` partial CTAs each. */
int fino_attention_plan(int64_t nq, int64_t nk, int heads, int batch, int sms, int mode, int* n_full, int* splits);
/* Same for a given head_dim: 128 -> 256-query-row tiles against 128-key tiles (what fino_attention_plan describes);
 * 64 -> the four-tile kernel's 512-query-row tiles against 64-key steps. *tile_rows receives the rows per CTA tile. */
int fino_attention_plan_hd(int64_t nq, int64_t nk, int heads, int batch, int head_dim, int sms, int mode, int* n_full,
                           int* splits, int* tile_rows);

/* Tuning / test hook: LayerNorm kernel 0 = warp per row, 1 = block per row, 2 = batched block per row (default for
 * 1024 <= dim <= 3072), 3 = packed warp per row (default); q/k-norm kernel 0 = generic warp per row, 1 = block per
 * token, 2 = packed warp per row (default; per-head LayerNorm(64): the 6-groups-in-flight kernel), 3 / 4 = per-head
 * LayerNorm(64) kernel with 12 / 4 groups in flight. */
int fino_rows_set_variant(int ln_block, int qk_block);
/* Tuning / test hook: 1 = wide rows (1024 <= dim <= 4096) go through the experimental TMA-staged persistent row
 * kernels (bulk async copies through a shared-memory ring); 0 (default) = the register-resident row kernels. */
int fino_rows_set_tma(int on);

#define FINO_LN_FLAG_BF16_STEPS 1 /* emulate the bf16 module flow of CogVideoXLayerNormZero / AdaLayerNorm */

/* out = LayerNorm(x) [*gamma + beta] [*(1+scale) + shift], fp32 math, one rounding.
 * Replaces FP32LayerNorm + modulate: transformer_wan.py:334,339,344-346,536; cogvideox_transformer_3d.py:134,150,536-541.
 * gamma/beta/shift/scale are float and optional (NULL). */
int fino_ln_modulate(const void* x, void* out, int64_t rows, int dim, int64_t x_stride, int64_t out_stride, float eps,
                     const float* gamma, const float* beta, const float* shift, const float* scale,
                     int64_t mod_row_stride, const int32_t* row_index, int64_t rows_per_group, int flags, void* stream);

/* out = x + y * gate (gate optional). Replaces transformer_wan.py:336,341,348 when not fused into a GEMM epilogue. */
int fino_gate_residual(const void* x, const void* y, void* out, int64_t rows, int dim, int64_t x_stride,
                       int64_t y_stride, int64_t out_stride, const float* gate, int64_t mod_row_stride,
                       const int32_t* row_index, int64_t rows_per_group, int round_product, void* stream);

#define FINO_QK_RMS_ACROSS_HEADS 0   /* RMSNorm over heads*head_dim  (Wan, attention_processor.py:208-211)   */
#define FINO_QK_LAYERNORM_PER_HEAD 1 /* LayerNorm(head_dim) per head (CogVideoX, attention_processor.py:195-197) */
#define FINO_ROPE_NONE 0
#define FINO_ROPE_WAN 1       /* transformer_wan.py:75-90       */
#define FINO_ROPE_COGVIDEOX 2 /* embeddings.py:1219-1258        */

/* In-place q/k normalisation + rotary embedding for up to two tensors (x1 may be NULL).
 * Replaces norm_q/norm_k + apply_rotary_emb: transformer_wan.py:64-90; attention_processor.py:2848-2860.
 * weights/biases are bf16 ([heads*head_dim] for RMS, [head_dim] for per-head LayerNorm); cos/sin are float
 * [*, head_dim] tables exactly as the reference rope modules produce them; token s of every `seq_len`-long sequence
 * uses table row s - rope_skip and tokens with s < rope_skip are not rotated. */
int fino_qk_norm_rope(void* x0, int64_t rows0, int64_t stride0, const void* w0, const void* b0, int rope0, void* x1,
                      int64_t rows1, int64_t stride1, const void* w1, const void* b1, int rope1, int heads,
                      int head_dim, int norm_mode, float eps, int rope_mode, const float* cos, const float* sin,
                      int64_t seq_len, int64_t rope_skip, void* stream);

/* x[b,c,f,h,w] (element strides sb..sw) -> rows[(b,f/pt,h/ph,w/pw), (c,pt,ph,pw)], row stride ld.
 * Turns the patch-embedding convolutions into a GEMM: transformer_wan.py:486-487; embeddings.py:734-738. */
int fino_patchify(const void* x, void* rows, int b, int c, int f, int h, int w, int pt, int ph, int pw, int64_t sb,
                  int64_t sc, int64_t sf, int64_t sh, int64_t sw, int64_t ld, void* stream);

/* rows -> out[b,c,f,h,w] (element strides sb..sw). channel_last = 1: row layout (pt,ph,pw,c)
 * (transformer_wan.py:539-543); 0: (c,pt,ph,pw) (cogvideox_transformer_3d.py:549-550). */
int fino_unpatchify(const void* rows, void* out, int b, int c, int f, int h, int w, int pt, int ph, int pw, int64_t sb,
                    int64_t sc, int64_t sf, int64_t sh, int64_t sw, int64_t ld, int channel_last, void* stream);

/* Sinusoidal timestep embedding, float in/out [n] -> [n, dim]. Replaces embeddings.py:27-78. */
int fino_timestep_embedding(const float* t, float* out, int n, int dim, int flip_sin_to_cos,
                            float downscale_freq_shift, float scale, float max_period, void* stream);

/* y[m,n] = act_out(act_in(x)[m,k] * w[n,k]^T + b), float activations and accumulation, w/b float or bf16; built for
 * m <= 8 (larger m runs as independent 8-row chunks, every row computed identically).
 * The de-duplicated time MLP: transformer_wan.py:182-183; cogvideox_transformer_3d.py:485. act: 0 none, 1 SiLU. */
int fino_linear_small_m(const float* x, const void* w, const void* b, float* y, int m, int n, int k, int w_is_bf16,
                        int act_in, int act_out, int round_in, int round_out, void* stream);

/* De-duplicates the per-token timesteps t[n] (float; transformer_wan.py:490-494 flattens them and runs the time MLP per
 * token) on the device, without a host round trip: uniq8[8] receives the distinct values in ascending order (unused
 * slots repeat the largest), row_index[i] the slot of t[i], count[0] the number of distinct values or 9 when more than 8
 * were seen (the outputs are then unusable and the caller must de-duplicate another way). The FrameINO sampler passes two
 * values: 0 on the clean first frame, t elsewhere (pipeline_wan_i2v_motion_FrameINO.py:832-843). */
int fino_timestep_dedup(const float* t, int64_t n, float* uniq8, int32_t* row_index, int32_t* count, void* stream);

/* out[l,r,c] = table[l*table_layer_stride + c] + proj[r,c] (float): the per-layer AdaLN rows
 * scale_shift_table + temb of transformer_wan.py:317-331, 520-527. */
int fino_build_mod_table(const float* table, const float* proj, float* out, int layers, int r, int cols,
                         int64_t table_layer_stride, void* stream);

/* ---- denoise-loop glue of the Wan2.2 FrameINO sampler, one kernel before and one after the
 * transformer forwards of a scheduler step. All 5-D tensors are contiguous float [b, c, frames, h, w]. ----
 *
 * rows[(b, f/pt, h/ph, w/pw), (c2, pt, ph, pw)] (bf16, row stride ld, c2 in [0, 2c)) = the patchified transformer input
 *   c2 <  c, frame <  f : (1 - mask) * condition + mask * latents     pipeline_wan_i2v_motion_FrameINO.py:829-830
 *   c2 <  c, frame >= f : id_latents[b, c2, frame - f]                 :854 (frame-wise ID concat)
 *   c2 >= c             : traj_latents[b, c2 - c, frame]               :858 (channel-wise trajectory concat)
 * mask is float [f, h, w] (the reference's [1,1,f,h,w] first_frame_mask); traj_latents has f + n_id frames;
 * id_latents may be NULL when n_id == 0. Fuses the concat/cast chain with the patchify of transformer_wan.py:486. */
int fino_wan_pack_model_input(const float* latents, const float* condition, const float* mask,
                              const float* id_latents, const float* traj_latents, void* rows, int b, int c, int f,
                              int n_id, int h, int w, int pt, int ph, int pw, int64_t ld, void* stream);

/* latents += dsigma * (y_uncond + guidance * (y_cond - y_uncond)) — the guidance combine evaluated in bf16 with the
 * reference's three roundings (it combines the bf16 forward outputs with tensor ops), the step in fp32 — over the f generated frames, reading the bf16
 * proj_out rows [(b, (f+n_id)/pt, h/ph, w/pw), (pt, ph, pw, c)] of the two forwards (row stride ld): classifier-free
 * guidance :882, ID-frame drop :886, flow-match Euler scheduler step :891, fused with the un-patchify of
 * transformer_wan.py:539-543. y_uncond may be NULL (no guidance). fp32, no FMA contraction. */
int fino_wan_cfg_euler_step(const void* y_cond, const void* y_uncond, int64_t ld, float* latents, int b, int c, int f,
                            int n_id, int h, int w, int pt, int ph, int pw, float guidance, float dsigma,
                            void* stream);

/* out[b][a][:] = in[a][b][:] with `inner` contiguous bf16 elements (multiple of 8) per (a,b): the pack / unpack step
 * around the Ulysses head<->sequence all-to-all (new capability, no reference counterpart: SURVEY.md 8e). */
int fino_swap01(const void* in, void* out, int64_t a, int64_t b, int64_t inner, void* stream);

/* ---- Ulysses exchange over NVLink/NVSwitch peer memory (new capability, no reference counterpart: SURVEY.md 8e) ----
 * One process per GPU. Each rank allocates its exchange buffers with fino_peer_alloc, exports a 64-byte CUDA IPC
 * handle per buffer, swaps the handles with its peers on the host (any transport) and maps theirs with
 * fino_peer_import. `*_ptrs` arguments below are HOST arrays of `world` device pointers, indexed by rank (the entry
 * for the calling rank is its own local pointer). world <= 8. */
int fino_peer_alloc(int64_t bytes, void** ptr);         /* cudaMalloc + zero fill                                  */
int fino_peer_free(void* ptr);
int fino_peer_export(const void* ptr, void* handle64);  /* handle64: 64-byte host buffer                           */
int fino_peer_import(const void* handle64, void** ptr); /* maps a peer's buffer (enables peer access lazily)       */
int fino_peer_release(void* ptr);                       /* unmaps an imported buffer                               */

/* Barrier between the ranks' streams through peer memory: flag_ptrs[r] = rank r's flag array (>= 16 uint32, zeroed).
 * `epoch` must increase by one per call (same sequence on every rank). Stream-ordered; no host synchronisation. A wait
 * that exceeds FINO_PEER_TIMEOUT_S seconds (environment, default 600, 0 = forever) is given up WITHOUT trapping: the
 * kernel records the epoch in word 8 + t of the rank's own flag array and returns; fino_peer_status reads those words. */
int fino_peer_barrier(void* const* flag_ptrs, int rank, int world, uint32_t epoch, void* stream);
/* status8[t] (HOST array of 8) = 0, or the epoch at which a barrier on this rank gave up waiting for rank t.
 * own_flags = this rank's flag array. Synchronises `stream`. */
int fino_peer_status(const void* own_flags, uint32_t* status8, void* stream);

/* Halo rows of a row-parallel 3x3 convolution (Wan VAE split by frame rows across GPUs; no reference counterpart — the
 * reference VAE, architecture/autoencoder_kl_wan.py, is single-GPU). frames: bf16 [t, hl + 2, W, C] (frame stride
 * frame_stride_bytes, rows of row_bytes = W*C*2), rows 1..hl written. One launch: rows 1 / hl of every frame are pushed
 * into the mailbox of the rank above / below (`up` / `down`: peer-mapped, NULL at the image border), flags raised to
 * `seq`, then rows 0 / hl + 1 are filled from this rank's mailbox `own` once the neighbours' pushes `seq` arrived.
 * Mailbox = fino_peer_alloc(256 + 4 * slot_bytes), zeroed; `seq` starts at 1 and increases by one per call, the same
 * sequence on every rank. Waits give up after FINO_PEER_TIMEOUT_S like fino_peer_barrier (words 4 / 5 of `own`). */
int fino_halo_exchange(void* frames, int t, int hl, int64_t row_bytes, int64_t frame_stride_bytes, void* up, void* down,
                       void* own, uint32_t seq, int64_t slot_bytes, int rank, void* stream);

/* First all-to-all fused into the q/k prologue: RMSNorm across heads (+ Wan RoPE when cos/sin != NULL; transformer_wan.py
 * :64-90) of the local fused projection rows qkv[rows, row_stride] (q | k | v, heads*head_dim columns each), stored
 * straight into the owning ranks' buffers: head group g = columns [g*inner, (g+1)*inner), inner = heads*head_dim/world,
 * of local token t lands in dst_ptrs[g] row rank*rows_per_rank + t as (q | k | v), inner columns each. cos/sin are the
 * float [rows, head_dim] table rows of the LOCAL tokens. */
int fino_qkv_norm_rope_scatter(const void* qkv, int64_t rows, int64_t row_stride, const void* wq, const void* wk,
                               int heads, int head_dim, float eps, const float* cos, const float* sin,
                               void* const* dst_ptrs, int world, int rank, int64_t rows_per_rank,
                               int64_t dst_row_stride, void* stream);

/* The CogVideoX form of the fused first all-to-all: per-head LayerNorm(64) with affine (attention_processor.py:2848-2851)
 * and the CogVideoX RoPE on local rows >= rope_skip (the text rows are not rotated, :2858-2860; cos/sin are the float
 * [rows - rope_skip, 64] table rows of the LOCAL video tokens) of qkv[rows, row_stride] (q | k | v), stored into the
 * owning ranks' buffers exactly as fino_qkv_norm_rope_scatter lays them out. head_dim must be 64. */
int fino_qkv_ln_rope_scatter(const void* qkv, int64_t rows, int64_t row_stride, const void* wq, const void* bq,
                             const void* wk, const void* bk, int heads, int head_dim, float eps, const float* cos,
                             const float* sin, int64_t rope_skip, void* const* dst_ptrs, int world, int rank,
                             int64_t rows_per_rank, int64_t dst_row_stride, void* stream);

/* fino_attention_fwd whose epilogue is the second all-to-all: query row g is stored into
 * o_owners[g / rows_per_owner] at local row g % rows_per_owner (row stride o_row_stride; pass the column offset of
 * this rank's heads in each pointer). o_owners: HOST array of num_owners (<= 8) device pointers. */
int fino_attention_fwd_scatter(const void* q, const void* k, const void* v, void* const* o_owners, int num_owners,
                               int64_t rows_per_owner, int batch, int heads, int64_t nq, int64_t nk, int head_dim,
                               int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride, int64_t o_row_stride,
                               int64_t q_batch_stride, int64_t k_batch_stride, int64_t v_batch_stride,
                               int64_t o_batch_stride, float scale, void* stream);

/* ---- Wan VAE encode / decode (reference architecture/autoencoder_kl_wan.py). Activations are
 * channels-last bf16 [T, H, W, C]: a pixel is a contiguous row of C channels. ----
 *
 * Causal 3-D / 2-D convolution as an implicit GEMM on the tcgen05 tensor cores (WanCausalConv3d :134-176, the Conv2d
 * of WanResample :245-262):  y[t,h,w,:] = bias + sum_{kt,kh,kw} W[:,kt,kh,kw,:] . x[t*stride_t + kt, h*s + kh - pad_h,
 * w*s + kw - pad_w, :].  "Valid" along time — the caller prepends the causal history frames (the reference's
 * feat_cache, :169-176); spatial out-of-range reads are zero. x: [t_in, h_in, w_in, c_in], element strides in_st /
 * in_sh / in_sw (frame / row / pixel). w: [c_out, kt*kh*kw * ceil(c_in/64)*64] (tap-major, each tap's channels
 * zero-padded to a multiple of 64), row stride ldw. y (and residual, same geometry): element strides out_st / out_sh /
 * out_sw with out_st, out_sh multiples of out_sw. c_in, c_out multiples of 8. epilogue: FINO_EPI_NONE, or
 * FINO_EPI_GATE_RESIDUAL = residual + bf16(conv + bias) (the block's skip connection, :382). */
int fino_conv3d_cl_bf16(const void* x, int t_in, int h_in, int w_in, int c_in, int64_t in_st, int64_t in_sh,
                        int64_t in_sw, const void* w, int64_t ldw, const void* bias, void* y, int t_out, int h_out,
                        int w_out, int c_out, int64_t out_st, int64_t out_sh, int64_t out_sw, int kt, int kh, int kw,
                        int pad_h, int pad_w, int stride_hw, int stride_t, const void* residual, int epilogue,
                        void* stream);
/* out[r,:] = act(x[r,:] / max(||x[r,:]||, 1e-12) * sqrt(c) * gamma + bias): WanRMS_norm (:179-202) fused with the SiLU
 * that follows it (silu = 1) or alone (silu = 0, the attention block's norm :406). gamma / bias float [c], bias optional. */
int fino_rms_act_cl(const void* x, void* out, int64_t rows, int c, int64_t x_stride, int64_t out_stride,
                    const float* gamma, const float* bias, int silu, void* stream);
/* out[t, 2h+i, 2w+j, :] = in[t, h, w, :]: WanUpsample(scale_factor 2, "nearest-exact") (:205-217). */
int fino_upsample2x_cl(const void* in, void* out, int t, int h, int w, int c, void* stream);
/* y += DupUp3D(src) (:90-131, :709-710), y [to,ho,wo,co], src [ti,hi,wi,ci], ho = hi*fs, wo = wi*fs; t_drop = ft - 1 for
 * the first chunk (first_chunk=True drops the first ft - 1 duplicated frames), else 0. */
int fino_dupup_add_cl(void* y, const void* src, int to, int ho, int wo, int co, int ti, int hi, int wi, int ci, int ft,
                      int fs, int t_drop, void* stream);
/* y += AvgDown3D(src) (:37-87, :502), y [to,ho,wo,co], src [ti,hi,wi,ci], hi = ho*fs, wi = wo*fs, frames zero-padded at
 * the front to a multiple of ft. */
int fino_avgdown_add_cl(void* y, const void* src, int to, int ho, int wo, int co, int ti, int hi, int wi, int ci, int ft,
                        int fs, void* stream);
/* p[r,:] = softmax(s[r,:cols] * scale) (float in, row stride ls; bf16 out, row stride lp, columns [cols, lp) zeroed):
 * the single-head attention of WanAttentionBlock (:413). cols <= 16384. */
int fino_softmax_rows(const float* s, void* p, int64_t rows, int cols, int64_t ls, int64_t lp, float scale, void* stream);
/* [c,t,h,w] (float when in_fp32 else bf16; element strides sc/st/sh/sw) -> channels-last bf16 [t, h/ps, w/ps, cpad] with
 * the patchify of :912-932 (channel c*ps*ps + (w%ps)*ps + h%ps), channels [c*ps*ps, cpad) zero. */
int fino_vae_to_cl(const void* in, int in_fp32, void* out, int c, int t, int h, int w, int64_t sc, int64_t st, int64_t sh,
                   int64_t sw, int ps, int cpad, void* stream);
/* channels-last bf16 [t, hi, wi, cstride] -> [c, t, hi*ps, wi*ps] (float when out_fp32 else bf16; channel stride out_sc
 * elements, the rest contiguous — a chunk of frames written into the full video): the unpatchify of :935-952, clamped
 * to [-1, 1] when clamp != 0 (:1224). */
int fino_vae_from_cl(const void* in, void* out, int out_fp32, int c, int t, int hi, int wi, int cstride, int ps, int clamp,
                     int64_t out_sc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FRAMEINO_B200_H_ */
