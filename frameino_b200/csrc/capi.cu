// extern "C" surface declared in include/frameino_b200.h. Plain pointers and sizes only.
#include "../../include/frameino_b200.h"
#include "common.cuh"
#include <atomic>

namespace fino {
int gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* c, int64_t ldc,
              int64_t m, int n, int k, int epilogue, int out_fp32, int flags, const void* residual, int64_t ldr,
              const float* gate, int64_t gate_row_stride, const int32_t* row_index, int64_t rows_per_group,
              cudaStream_t stream);
int attention_fwd(const void* q, const void* k, const void* v, void* o, int batch, int heads, int64_t nq, int64_t nk,
                  int head_dim, int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride, int64_t o_row_stride,
                  int64_t q_batch_stride, int64_t k_batch_stride, int64_t v_batch_stride, int64_t o_batch_stride,
                  float scale, cudaStream_t stream);
int ln_modulate(const void* x, void* out, int64_t rows, int dim, int64_t x_stride, int64_t out_stride, float eps,
                const float* gamma, const float* beta, const float* shift, const float* scale, int64_t mod_row_stride,
                const int32_t* row_index, int64_t rows_per_group, int flags, cudaStream_t stream);
int gate_residual(const void* x, const void* y, void* out, int64_t rows, int dim, int64_t x_stride, int64_t y_stride,
                  int64_t out_stride, const float* gate, int64_t mod_row_stride, const int32_t* row_index,
                  int64_t rows_per_group, int round_product, cudaStream_t stream);
int qk_norm_rope(void* x0, int64_t rows0, int64_t stride0, const void* w0, const void* b0, int rope0, void* x1,
                 int64_t rows1, int64_t stride1, const void* w1, const void* b1, int rope1, int heads, int head_dim,
                 int norm_mode, float eps, int rope_mode, const float* cos, const float* sin, int64_t seq_len,
                 int64_t rope_skip, cudaStream_t stream);
int patchify(const void* x, void* rows, int B, int C, int F, int H, int W, int pt, int ph, int pw, int64_t sb,
             int64_t sc, int64_t sf, int64_t sh, int64_t sw, int64_t ld, cudaStream_t stream);
int unpatchify(const void* rows, void* out, int B, int C, int F, int H, int W, int pt, int ph, int pw, int64_t sb,
               int64_t sc, int64_t sf, int64_t sh, int64_t sw, int64_t ld, int channel_last, cudaStream_t stream);
int timestep_embedding(const float* t, float* out, int n, int dim, int flip_sin_to_cos, float downscale_freq_shift,
                       float scale, float max_period, cudaStream_t stream);
int linear_small_m(const float* x, const void* w, const void* b, float* y, int m, int n, int k, int w_is_bf16,
                   int act_in, int act_out, int round_in, int round_out, cudaStream_t stream);
int timestep_dedup(const float* t, int64_t n, float* uniq, int32_t* row_index, int32_t* count, cudaStream_t stream);
int build_mod_table(const float* table, const float* proj, float* out, int layers, int r, int cols,
                    int64_t table_layer_stride, cudaStream_t stream);
int swap01(const void* in, void* out, int64_t A, int64_t B, int64_t inner, cudaStream_t stream);
int attention_fwd_owners(const void* q, const void* k, const void* v, void* const* o_owners, int num_owners,
                         int64_t rows_per_owner, int batch, int heads, int64_t nq, int64_t nk, int head_dim,
                         int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride, int64_t o_row_stride,
                         int64_t q_batch_stride, int64_t k_batch_stride, int64_t v_batch_stride,
                         int64_t o_batch_stride, float scale, cudaStream_t stream);
int peer_alloc(int64_t bytes, void** ptr);
int peer_free(void* ptr);
int peer_export(const void* ptr, void* handle64);
int peer_import(const void* handle64, void** ptr);
int peer_release(void* ptr);
int peer_barrier(void* const* flag_ptrs, int rank, int world, uint32_t epoch, cudaStream_t stream);
int peer_status(const void* own_flags, uint32_t* status8, cudaStream_t stream);
int halo_exchange(void* frames, int t, int hl, int64_t row_bytes, int64_t frame_stride_bytes, void* up, void* down,
                  void* own, uint32_t seq, int64_t slot_bytes, int rank, cudaStream_t stream);
int qkv_norm_rope_scatter(const void* qkv, int64_t rows, int64_t row_stride, const void* wq, const void* wk,
                          int heads, int head_dim, float eps, const float* cos, const float* sin,
                          void* const* dst_ptrs, int world, int rank, int64_t rows_per_rank, int64_t dst_row_stride,
                          cudaStream_t stream);
int qkv_ln_rope_scatter(const void* qkv, int64_t rows, int64_t row_stride, const void* wq, const void* bq, const void* wk,
                        const void* bk, int heads, int head_dim, float eps, const float* cos, const float* sin,
                        int64_t rope_skip, void* const* dst_ptrs, int world, int rank, int64_t rows_per_rank,
                        int64_t dst_row_stride, cudaStream_t stream);
int wan_pack_model_input(const float* latents, const float* condition, const float* mask, const float* id_latents,
                         const float* traj, void* rows, int B, int C, int F, int NID, int H, int W, int pt, int ph,
                         int pw, int64_t ld, cudaStream_t stream);
int wan_cfg_euler_step(const void* y_cond, const void* y_uncond, int64_t ld, float* latents, int B, int C, int F,
                       int NID, int H, int W, int pt, int ph, int pw, float guidance, float dsigma,
                       cudaStream_t stream);
int conv3d_cl(const void* x, int t_in, int h_in, int w_in, int c_in, int64_t in_st, int64_t in_sh, int64_t in_sw,
              const void* w, int64_t ldw, const void* bias, void* y, int t_out, int h_out, int w_out, int c_out,
              int64_t out_st, int64_t out_sh, int64_t out_sw, int kt, int kh, int kw, int pad_h, int pad_w,
              int stride_hw, int stride_t, const void* residual, int epilogue, cudaStream_t stream);
int rms_act_cl(const void* x, void* out, int64_t rows, int C, int64_t x_stride, int64_t out_stride, const float* gamma,
               const float* bias, int silu, cudaStream_t stream);
int upsample2x_cl(const void* in, void* out, int T, int H, int W, int C, cudaStream_t stream);
int dupup_add_cl(void* y, const void* src, int To, int Ho, int Wo, int Co, int Ti, int Hi, int Wi, int Ci, int ft, int fs,
                 int t_drop, cudaStream_t stream);
int avgdown_add_cl(void* y, const void* src, int To, int Ho, int Wo, int Co, int Ti, int Hi, int Wi, int Ci, int ft, int fs,
                   cudaStream_t stream);
int softmax_rows(const float* s, void* p, int64_t rows, int cols, int64_t ls, int64_t lp, float scale, cudaStream_t stream);
int vae_to_cl(const void* in, int in_fp32, void* out, int C, int T, int H, int W, int64_t sc, int64_t st, int64_t sh,
              int64_t sw, int ps, int cpad, cudaStream_t stream);
int vae_from_cl(const void* in, void* out, int out_fp32, int C, int T, int Hi, int Wi, int cstride, int ps, int clamp,
                int64_t out_sc, cudaStream_t stream);
void gemm_set_mode(int mode);
void gemm_set_split(int mode);
void gemm_plan(int tiles, int num_kb, int clusters, int mode, int* num_full, int* splits);
void attention_set_variant(int v);
void attention_set_split(int s);
void attention_plan(int n_tiles, int T, int sms, int mode, int* n_full, int* splits);
void rows_set_variant(int ln_block, int qk_block);
void rows_set_tma(int on);
void scatter_set_tma(int on);
void scatter_set_packed(int on);
}  // namespace fino

static std::atomic<int64_t> g_launches{0};
static int g_device = -1;

// Every compute entry point: make the bound device current for the duration of the call and put the caller's current
// device back afterwards (the library must not move torch's current device behind its back; per-device state such as
// the dynamic-shared-memory opt-in and the split workspaces is keyed by the device that is current inside the call).
struct DeviceGuard {
  int prev = -1;
  int status = 0;
  DeviceGuard() {
    if (g_device >= 0) {
      if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
      if (prev != g_device) {
        cudaError_t e = cudaSetDevice(g_device);
        if (e != cudaSuccess) {
          fino::set_last_error("cudaSetDevice(%d) failed: %s", g_device, cudaGetErrorString(e));
          status = fino::FINO_ERR_CUDA;
          prev = -1;
        }
      } else {
        prev = -1;  // nothing to restore
      }
      return;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
      fino::set_last_error("no CUDA device available (%s); frameino_b200 has no CPU fallback",
                           e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
      status = fino::FINO_ERR_CUDA;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static int ensure_device() {  // allocation / IPC entry points: bind without restoring (they are device management)
  if (g_device >= 0) {
    cudaError_t e = cudaSetDevice(g_device);
    if (e != cudaSuccess) {
      fino::set_last_error("cudaSetDevice(%d) failed: %s", g_device, cudaGetErrorString(e));
      return fino::FINO_ERR_CUDA;
    }
    return 0;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    fino::set_last_error("no CUDA device available (%s); frameino_b200 has no CPU fallback",
                         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return fino::FINO_ERR_CUDA;
  }
  return 0;
}

#define FINO_ENTRY(call)                  \
  do {                                    \
    DeviceGuard _g;                       \
    if (_g.status) return _g.status;      \
    int _r = (call);                      \
    if (_r == 0) g_launches.fetch_add(1, std::memory_order_relaxed); \
    return _r;                            \
  } while (0)

extern "C" {

int fino_abi_version(void) { return FINO_ABI_VERSION; }
const char* fino_last_error(void) { return fino::get_last_error(); }
int64_t fino_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int fino_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    fino::set_last_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
    return fino::FINO_ERR_CUDA;
  }
  g_device = device;
  return 0;
}

int fino_gemm_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* c, int64_t ldc,
                   int64_t m, int n, int k, int epilogue, int out_fp32, int flags, const void* residual, int64_t ldr,
                   const float* gate, int64_t gate_row_stride, const int32_t* row_index, int64_t rows_per_group,
                   void* stream) {
  FINO_ENTRY(fino::gemm_bf16(a, lda, w, ldw, bias, c, ldc, m, n, k, epilogue, out_fp32, flags, residual, ldr, gate,
                             gate_row_stride, row_index, rows_per_group, (cudaStream_t)stream));
}

int fino_attention_fwd(const void* q, const void* k, const void* v, void* o, int batch, int heads, int64_t nq,
                       int64_t nk, int head_dim, int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride,
                       int64_t o_row_stride, int64_t q_batch_stride, int64_t k_batch_stride, int64_t v_batch_stride,
                       int64_t o_batch_stride, float scale, void* stream) {
  FINO_ENTRY(fino::attention_fwd(q, k, v, o, batch, heads, nq, nk, head_dim, q_row_stride, k_row_stride, v_row_stride,
                                 o_row_stride, q_batch_stride, k_batch_stride, v_batch_stride, o_batch_stride, scale,
                                 (cudaStream_t)stream));
}

int fino_ln_modulate(const void* x, void* out, int64_t rows, int dim, int64_t x_stride, int64_t out_stride, float eps,
                     const float* gamma, const float* beta, const float* shift, const float* scale,
                     int64_t mod_row_stride, const int32_t* row_index, int64_t rows_per_group, int flags,
                     void* stream) {
  FINO_ENTRY(fino::ln_modulate(x, out, rows, dim, x_stride, out_stride, eps, gamma, beta, shift, scale, mod_row_stride,
                               row_index, rows_per_group, flags, (cudaStream_t)stream));
}

int fino_gate_residual(const void* x, const void* y, void* out, int64_t rows, int dim, int64_t x_stride,
                       int64_t y_stride, int64_t out_stride, const float* gate, int64_t mod_row_stride,
                       const int32_t* row_index, int64_t rows_per_group, int round_product, void* stream) {
  FINO_ENTRY(fino::gate_residual(x, y, out, rows, dim, x_stride, y_stride, out_stride, gate, mod_row_stride, row_index,
                                 rows_per_group, round_product, (cudaStream_t)stream));
}

int fino_qk_norm_rope(void* x0, int64_t rows0, int64_t stride0, const void* w0, const void* b0, int rope0, void* x1,
                      int64_t rows1, int64_t stride1, const void* w1, const void* b1, int rope1, int heads,
                      int head_dim, int norm_mode, float eps, int rope_mode, const float* cos, const float* sin,
                      int64_t seq_len, int64_t rope_skip, void* stream) {
  FINO_ENTRY(fino::qk_norm_rope(x0, rows0, stride0, w0, b0, rope0, x1, rows1, stride1, w1, b1, rope1, heads, head_dim,
                                norm_mode, eps, rope_mode, cos, sin, seq_len, rope_skip, (cudaStream_t)stream));
}

int fino_patchify(const void* x, void* rows, int b, int c, int f, int h, int w, int pt, int ph, int pw, int64_t sb,
                  int64_t sc, int64_t sf, int64_t sh, int64_t sw, int64_t ld, void* stream) {
  FINO_ENTRY(fino::patchify(x, rows, b, c, f, h, w, pt, ph, pw, sb, sc, sf, sh, sw, ld, (cudaStream_t)stream));
}

int fino_unpatchify(const void* rows, void* out, int b, int c, int f, int h, int w, int pt, int ph, int pw, int64_t sb,
                    int64_t sc, int64_t sf, int64_t sh, int64_t sw, int64_t ld, int channel_last, void* stream) {
  FINO_ENTRY(fino::unpatchify(rows, out, b, c, f, h, w, pt, ph, pw, sb, sc, sf, sh, sw, ld, channel_last,
                              (cudaStream_t)stream));
}

int fino_timestep_embedding(const float* t, float* out, int n, int dim, int flip_sin_to_cos,
                            float downscale_freq_shift, float scale, float max_period, void* stream) {
  FINO_ENTRY(fino::timestep_embedding(t, out, n, dim, flip_sin_to_cos, downscale_freq_shift, scale, max_period,
                                      (cudaStream_t)stream));
}

int fino_linear_small_m(const float* x, const void* w, const void* b, float* y, int m, int n, int k, int w_is_bf16,
                        int act_in, int act_out, int round_in, int round_out, void* stream) {
  FINO_ENTRY(fino::linear_small_m(x, w, b, y, m, n, k, w_is_bf16, act_in, act_out, round_in, round_out,
                                  (cudaStream_t)stream));
}

int fino_timestep_dedup(const float* t, int64_t n, float* uniq8, int32_t* row_index, int32_t* count, void* stream) {
  FINO_ENTRY(fino::timestep_dedup(t, n, uniq8, row_index, count, (cudaStream_t)stream));
}

int fino_build_mod_table(const float* table, const float* proj, float* out, int layers, int r, int cols,
                         int64_t table_layer_stride, void* stream) {
  FINO_ENTRY(fino::build_mod_table(table, proj, out, layers, r, cols, table_layer_stride, (cudaStream_t)stream));
}

int fino_gemm_set_mode(int mode) {
  if (mode < 0 || mode > 3) {
    fino::set_last_error("fino_gemm_set_mode: mode %d out of range (0 auto, 1 single-CTA, 2 CTA-pair 256x256, 3 CTA-pair 512x256)", mode);
    return fino::FINO_ERR_INVALID;
  }
  fino::gemm_set_mode(mode);
  return 0;
}

int fino_gemm_set_split(int mode) {
  if (mode < -1 || mode == 1 || mode > 16) {
    fino::set_last_error("fino_gemm_set_split: mode %d (want -1 auto, 0 off, or 2..16 forced K slices)", mode);
    return fino::FINO_ERR_INVALID;
  }
  fino::gemm_set_split(mode);
  return 0;
}

int fino_gemm_plan(int64_t m, int n, int k, int sms, int mode, int* num_full, int* splits) {
  if (m <= 0 || n <= 0 || k <= 0 || sms < 2 || !num_full || !splits) {
    fino::set_last_error("fino_gemm_plan: bad arguments");
    return fino::FINO_ERR_INVALID;
  }
  const int64_t tiles = ((m + 255) / 256) * ((n + 255) / 256);
  if (tiles >= ((int64_t)1 << 30)) {
    fino::set_last_error("fino_gemm_plan: too many tiles");
    return fino::FINO_ERR_INVALID;
  }
  fino::gemm_plan((int)tiles, (k + 63) / 64, sms / 2, mode, num_full, splits);
  return 0;
}

int fino_attention_set_variant(int variant) {
  if (variant < 0 || variant > 15) {
    fino::set_last_error("fino_attention_set_variant: variant %d out of range (0..15)", variant);
    return fino::FINO_ERR_INVALID;
  }
  fino::attention_set_variant(variant);
  return 0;
}

int fino_attention_set_split(int mode) {
  if (mode < -1 || mode == 1 || mode > 64) {
    fino::set_last_error("fino_attention_set_split: mode %d (want -1 auto, 0 off, or 2..64 forced splits)", mode);
    return fino::FINO_ERR_INVALID;
  }
  fino::attention_set_split(mode);
  return 0;
}

int fino_attention_plan(int64_t nq, int64_t nk, int heads, int batch, int sms, int mode, int* n_full, int* splits) {
  if (nq <= 0 || nk <= 0 || heads <= 0 || batch <= 0 || sms <= 0 || !n_full || !splits) {
    fino::set_last_error("fino_attention_plan: bad arguments");
    return fino::FINO_ERR_INVALID;
  }
  const int64_t tiles = (nq + 255) / 256 * heads * batch;
  if (tiles >= ((int64_t)1 << 30)) {
    fino::set_last_error("fino_attention_plan: too many query tiles");
    return fino::FINO_ERR_INVALID;
  }
  fino::attention_plan((int)tiles, (int)((nk + 127) / 128), sms, mode, n_full, splits);
  return 0;
}

int fino_attention_plan_hd(int64_t nq, int64_t nk, int heads, int batch, int head_dim, int sms, int mode, int* n_full,
                           int* splits, int* tile_rows) {
  if (nq <= 0 || nk <= 0 || heads <= 0 || batch <= 0 || sms <= 0 || !n_full || !splits || !tile_rows ||
      (head_dim != 64 && head_dim != 128)) {
    fino::set_last_error("fino_attention_plan_hd: bad arguments (head_dim 64 or 128)");
    return fino::FINO_ERR_INVALID;
  }
  const int rows = head_dim == 64 ? 512 : 256, keys = head_dim == 64 ? 64 : 128;
  const int64_t tiles = (nq + rows - 1) / rows * heads * batch;
  if (tiles >= ((int64_t)1 << 30)) {
    fino::set_last_error("fino_attention_plan_hd: too many query tiles");
    return fino::FINO_ERR_INVALID;
  }
  *tile_rows = rows;
  fino::attention_plan((int)tiles, (int)((nk + keys - 1) / keys), sms, mode, n_full, splits);
  return 0;
}

int fino_rows_set_variant(int ln_block, int qk_block) {
  fino::rows_set_variant(ln_block, qk_block);
  fino::scatter_set_packed(qk_block == 2);
  return 0;
}

int fino_rows_set_tma(int on) {
  fino::rows_set_tma(on);
  fino::scatter_set_tma(on);
  return 0;
}

int fino_swap01(const void* in, void* out, int64_t a, int64_t b, int64_t inner, void* stream) {
  FINO_ENTRY(fino::swap01(in, out, a, b, inner, (cudaStream_t)stream));
}

int fino_attention_fwd_scatter(const void* q, const void* k, const void* v, void* const* o_owners, int num_owners,
                               int64_t rows_per_owner, int batch, int heads, int64_t nq, int64_t nk, int head_dim,
                               int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride, int64_t o_row_stride,
                               int64_t q_batch_stride, int64_t k_batch_stride, int64_t v_batch_stride,
                               int64_t o_batch_stride, float scale, void* stream) {
  FINO_ENTRY(fino::attention_fwd_owners(q, k, v, o_owners, num_owners, rows_per_owner, batch, heads, nq, nk, head_dim,
                                        q_row_stride, k_row_stride, v_row_stride, o_row_stride, q_batch_stride,
                                        k_batch_stride, v_batch_stride, o_batch_stride, scale, (cudaStream_t)stream));
}

int fino_qkv_norm_rope_scatter(const void* qkv, int64_t rows, int64_t row_stride, const void* wq, const void* wk,
                               int heads, int head_dim, float eps, const float* cos, const float* sin,
                               void* const* dst_ptrs, int world, int rank, int64_t rows_per_rank,
                               int64_t dst_row_stride, void* stream) {
  FINO_ENTRY(fino::qkv_norm_rope_scatter(qkv, rows, row_stride, wq, wk, heads, head_dim, eps, cos, sin, dst_ptrs, world,
                                         rank, rows_per_rank, dst_row_stride, (cudaStream_t)stream));
}

int fino_qkv_ln_rope_scatter(const void* qkv, int64_t rows, int64_t row_stride, const void* wq, const void* bq,
                             const void* wk, const void* bk, int heads, int head_dim, float eps, const float* cos,
                             const float* sin, int64_t rope_skip, void* const* dst_ptrs, int world, int rank,
                             int64_t rows_per_rank, int64_t dst_row_stride, void* stream) {
  FINO_ENTRY(fino::qkv_ln_rope_scatter(qkv, rows, row_stride, wq, bq, wk, bk, heads, head_dim, eps, cos, sin, rope_skip,
                                       dst_ptrs, world, rank, rows_per_rank, dst_row_stride, (cudaStream_t)stream));
}

int fino_peer_barrier(void* const* flag_ptrs, int rank, int world, uint32_t epoch, void* stream) {
  FINO_ENTRY(fino::peer_barrier(flag_ptrs, rank, world, epoch, (cudaStream_t)stream));
}

int fino_peer_status(const void* own_flags, uint32_t* status8, void* stream) {
  int d = ensure_device();
  return d ? d : fino::peer_status(own_flags, status8, (cudaStream_t)stream);
}

int fino_halo_exchange(void* frames, int t, int hl, int64_t row_bytes, int64_t frame_stride_bytes, void* up, void* down,
                       void* own, uint32_t seq, int64_t slot_bytes, int rank, void* stream) {
  FINO_ENTRY(fino::halo_exchange(frames, t, hl, row_bytes, frame_stride_bytes, up, down, own, seq, slot_bytes, rank,
                                 (cudaStream_t)stream));
}

int fino_wan_pack_model_input(const float* latents, const float* condition, const float* mask,
                              const float* id_latents, const float* traj_latents, void* rows, int b, int c, int f,
                              int n_id, int h, int w, int pt, int ph, int pw, int64_t ld, void* stream) {
  FINO_ENTRY(fino::wan_pack_model_input(latents, condition, mask, id_latents, traj_latents, rows, b, c, f, n_id, h, w,
                                        pt, ph, pw, ld, (cudaStream_t)stream));
}

int fino_wan_cfg_euler_step(const void* y_cond, const void* y_uncond, int64_t ld, float* latents, int b, int c, int f,
                            int n_id, int h, int w, int pt, int ph, int pw, float guidance, float dsigma,
                            void* stream) {
  FINO_ENTRY(fino::wan_cfg_euler_step(y_cond, y_uncond, ld, latents, b, c, f, n_id, h, w, pt, ph, pw, guidance, dsigma,
                                      (cudaStream_t)stream));
}

// ---- Wan VAE ----
int fino_conv3d_cl_bf16(const void* x, int t_in, int h_in, int w_in, int c_in, int64_t in_st, int64_t in_sh,
                        int64_t in_sw, const void* w, int64_t ldw, const void* bias, void* y, int t_out, int h_out,
                        int w_out, int c_out, int64_t out_st, int64_t out_sh, int64_t out_sw, int kt, int kh, int kw,
                        int pad_h, int pad_w, int stride_hw, int stride_t, const void* residual, int epilogue,
                        void* stream) {
  FINO_ENTRY(fino::conv3d_cl(x, t_in, h_in, w_in, c_in, in_st, in_sh, in_sw, w, ldw, bias, y, t_out, h_out, w_out, c_out,
                             out_st, out_sh, out_sw, kt, kh, kw, pad_h, pad_w, stride_hw, stride_t, residual, epilogue,
                             (cudaStream_t)stream));
}
int fino_rms_act_cl(const void* x, void* out, int64_t rows, int c, int64_t x_stride, int64_t out_stride,
                    const float* gamma, const float* bias, int silu, void* stream) {
  FINO_ENTRY(fino::rms_act_cl(x, out, rows, c, x_stride, out_stride, gamma, bias, silu, (cudaStream_t)stream));
}
int fino_upsample2x_cl(const void* in, void* out, int t, int h, int w, int c, void* stream) {
  FINO_ENTRY(fino::upsample2x_cl(in, out, t, h, w, c, (cudaStream_t)stream));
}
int fino_dupup_add_cl(void* y, const void* src, int to, int ho, int wo, int co, int ti, int hi, int wi, int ci, int ft,
                      int fs, int t_drop, void* stream) {
  FINO_ENTRY(fino::dupup_add_cl(y, src, to, ho, wo, co, ti, hi, wi, ci, ft, fs, t_drop, (cudaStream_t)stream));
}
int fino_avgdown_add_cl(void* y, const void* src, int to, int ho, int wo, int co, int ti, int hi, int wi, int ci, int ft,
                        int fs, void* stream) {
  FINO_ENTRY(fino::avgdown_add_cl(y, src, to, ho, wo, co, ti, hi, wi, ci, ft, fs, (cudaStream_t)stream));
}
int fino_softmax_rows(const float* s, void* p, int64_t rows, int cols, int64_t ls, int64_t lp, float scale, void* stream) {
  FINO_ENTRY(fino::softmax_rows(s, p, rows, cols, ls, lp, scale, (cudaStream_t)stream));
}
int fino_vae_to_cl(const void* in, int in_fp32, void* out, int c, int t, int h, int w, int64_t sc, int64_t st, int64_t sh,
                   int64_t sw, int ps, int cpad, void* stream) {
  FINO_ENTRY(fino::vae_to_cl(in, in_fp32, out, c, t, h, w, sc, st, sh, sw, ps, cpad, (cudaStream_t)stream));
}
int fino_vae_from_cl(const void* in, void* out, int out_fp32, int c, int t, int hi, int wi, int cstride, int ps, int clamp,
                     int64_t out_sc, void* stream) {
  FINO_ENTRY(fino::vae_from_cl(in, out, out_fp32, c, t, hi, wi, cstride, ps, clamp, out_sc, (cudaStream_t)stream));
}

// allocation / IPC: no kernel launch, not counted
int fino_peer_alloc(int64_t bytes, void** ptr) {
  int d = ensure_device();
  return d ? d : fino::peer_alloc(bytes, ptr);
}
int fino_peer_free(void* ptr) {
  int d = ensure_device();
  return d ? d : fino::peer_free(ptr);
}
int fino_peer_export(const void* ptr, void* handle64) {
  int d = ensure_device();
  return d ? d : fino::peer_export(ptr, handle64);
}
int fino_peer_import(const void* handle64, void** ptr) {
  int d = ensure_device();
  return d ? d : fino::peer_import(handle64, ptr);
}
int fino_peer_release(void* ptr) {
  int d = ensure_device();
  return d ? d : fino::peer_release(ptr);
}

}  // extern "C"
