// Ulysses head<->sequence exchange over NVLink / NVSwitch PEER MEMORY, fused into the kernels on either side of it
// (new capability — the reference is single-GPU). One process per GPU; every rank cudaMalloc's its
// exchange buffers, exports them with CUDA IPC and maps the buffers of all peers, so a kernel can store straight into
// another GPU's HBM:
//
//   qkv_norm_rope_scatter   RMSNorm-across-heads + 3-D RoPE of the local tokens' q/k (reference transformer_wan.py:64-90)
//                           whose stores ARE the first all-to-all: head group g of every local token row is written
//                           into rank g's [N, 3*inner] buffer at the token's global row (v rides along un-normalised).
//   attention (scatter O)   attention_tcgen05.cu with one output "owner" per rank: the epilogue stores each query
//                           row's heads into the owning rank's [n_loc, D] buffer — the second all-to-all.
//   peer_barrier            flag exchange through peer memory between the two (all ranks' stores have landed).
//
// Nothing here calls NCCL; torch.distributed is only used once, on the host, to swap the 64-byte IPC handles.
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"

namespace fino {

constexpr int PEER_MAX_RANKS = 8;

// ------------------------------------------------------------------------------------------------
// allocation + CUDA IPC
// ------------------------------------------------------------------------------------------------
int peer_alloc(int64_t bytes, void** ptr) {
  FINO_CHECK_ARG(bytes > 0 && ptr != nullptr, "peer_alloc: bad arguments");
  void* p = nullptr;
  FINO_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    set_last_error("peer_alloc: cudaMemset failed: %s", cudaGetErrorString(e));
    return FINO_ERR_CUDA;
  }
  *ptr = p;
  return FINO_OK;
}

int peer_free(void* ptr) {
  if (ptr != nullptr) FINO_CHECK_CUDA(cudaFree(ptr));
  return FINO_OK;
}

int peer_export(const void* ptr, void* handle64) {
  FINO_CHECK_ARG(ptr != nullptr && handle64 != nullptr, "peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  FINO_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64, &h, 64);
  return FINO_OK;
}

int peer_import(const void* handle64, void** ptr) {
  FINO_CHECK_ARG(handle64 != nullptr && ptr != nullptr, "peer_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  FINO_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return FINO_OK;
}

int peer_release(void* ptr) {
  if (ptr != nullptr) FINO_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// barrier through peer memory
// ------------------------------------------------------------------------------------------------
struct BarrierParams {
  uint32_t* flags[PEER_MAX_RANKS];  // flags[r] = rank r's flag array (>= 2 * PEER_MAX_RANKS words), peer-mapped
  int rank, world;
  uint32_t epoch;
  unsigned long long timeout_ns;  // 0 = wait forever
};

// Word PEER_MAX_RANKS + t of a rank's own flag array: non-zero once a wait on rank t timed out (epoch it was waiting
// for). The host reads it (fino_peer_status); the kernel does NOT trap — a trap is a sticky context error on this rank
// and a hang on the others — it gives up the wait, so the step that follows computes on stale data and the host check
// turns that into an exception.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Thread t < world: tell rank t that `rank` reached `epoch`, then wait until rank t told us the same. The kernel is
// stream-ordered after the producer kernel, whose (peer) stores are complete when it retires; the release/acquire
// pair at system scope orders them against the consumer kernel that follows the barrier on every rank.
__global__ void peer_barrier_kernel(const __grid_constant__ BarrierParams p) {
  const int t = threadIdx.x;
  if (t < p.world) {
    __threadfence_system();
    uint32_t* remote = p.flags[t] + p.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(p.epoch) : "memory");
    const uint32_t* mine = p.flags[p.rank] + t;
    uint32_t v;
    unsigned long long t0 = 0;
    unsigned spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - p.epoch) >= 0) break;
      __nanosleep(64);
      if ((++spins & 0xfffu) == 0 && p.timeout_ns != 0) {  // look at the clock every 4096 polls
        const unsigned long long now = global_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > p.timeout_ns) {
          printf("fino peer_barrier timeout: rank %d gave up waiting for rank %d at epoch %u (saw %u)\n", p.rank, t,
                 p.epoch, v);
          p.flags[p.rank][PEER_MAX_RANKS + t] = p.epoch ? p.epoch : 1u;
          break;
        }
      }
    } while (true);
    __threadfence_system();
  }
}

// Wait limit of the barrier: FINO_PEER_TIMEOUT_S seconds (default 600; 0 = wait forever). Rank skew far beyond a kernel's
// duration is legitimate (lazy module loads, host-side preprocessing on one rank, a debugger), so the default is generous.
static unsigned long long peer_timeout_ns() {
  static long long cached = -1;
  if (cached < 0) {
    const char* e = getenv("FINO_PEER_TIMEOUT_S");
    double s = e ? atof(e) : 600.0;
    if (s < 0) s = 0;
    cached = (long long)(s * 1e9);
  }
  return (unsigned long long)cached;
}

int peer_barrier(void* const* flag_ptrs, int rank, int world, uint32_t epoch, cudaStream_t stream) {
  FINO_CHECK_ARG(flag_ptrs != nullptr && world >= 1 && world <= PEER_MAX_RANKS && rank >= 0 && rank < world,
                 "peer_barrier: bad rank/world (%d/%d)", rank, world);
  BarrierParams p;
  for (int r = 0; r < PEER_MAX_RANKS; ++r) p.flags[r] = reinterpret_cast<uint32_t*>(flag_ptrs[r < world ? r : 0]);
  for (int r = 0; r < world; ++r) FINO_CHECK_ARG(flag_ptrs[r] != nullptr, "peer_barrier: null flag pointer %d", r);
  p.rank = rank;
  p.world = world;
  p.epoch = epoch;
  p.timeout_ns = peer_timeout_ns();
  peer_barrier_kernel<<<1, 32, 0, stream>>>(p);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// Copies the calling rank's timeout words (see above) to the host: status[t] != 0 means a barrier gave up waiting for
// rank t. Synchronises the stream.
int peer_status(const void* own_flags, uint32_t* status8, cudaStream_t stream) {
  FINO_CHECK_ARG(own_flags && status8, "peer_status: null pointer");
  FINO_CHECK_CUDA(cudaMemcpyAsync(status8, reinterpret_cast<const uint32_t*>(own_flags) + PEER_MAX_RANKS,
                                  PEER_MAX_RANKS * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  FINO_CHECK_CUDA(cudaStreamSynchronize(stream));
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// halo rows of a row-parallel convolution (VAE, frameino_b200/vae.py RowParallel), pushed through peer memory
// ------------------------------------------------------------------------------------------------
// Every rank owns a band of the frame rows; a 3x3 convolution's input frames [t, hl + 2, W, C] carry one halo row above
// and below the band. One launch per convolution and rank: (1) PUSH the band's first row of every frame into the rank
// above's mailbox ("from below" slot) and its last row into the rank below's ("from above" slot), (2) the last CTA to
// finish raises the neighbours' flags to `seq`, (3) wait until both own flags reached `seq`, (4) PULL the two mailbox
// slots into the halo rows. Mailbox (peer-mapped, zeroed): words 0 / 1 = flag from above / from below, word 2 = CTA
// arrival counter, words 4 / 5 = time-out records; data at byte 256: slot (seq & 1, direction) of `slot_bytes` each.
// Two slot parities suffice: a neighbour can run at most one exchange ahead (its exchange seq + 1 needs this rank's
// push seq + 1, which is stream-ordered behind this rank's pull seq). All CTAs must be co-resident (they spin in (3)
// while the last one signals in (2)): the grid is at most HALO_MAX_CTAS.
constexpr int HALO_DATA_OFF = 256;
constexpr int HALO_MAX_CTAS = 32;
struct HaloParams {
  char* frames;          // first frame of the chunk, halo row 0
  int t, hl;             // frames, band rows
  long long row_bytes, frame_stride;
  char *up, *down, *own; // mailboxes of the rank above / below (null at the image border) and this rank's
  uint32_t seq;
  long long slot_bytes;
  unsigned long long timeout_ns;
  int rank;
};

__device__ __forceinline__ void halo_wait(const uint32_t* flag, uint32_t seq, const HaloParams& p, int which) {
  uint32_t v;
  unsigned long long t0 = 0;
  unsigned spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if ((int32_t)(v - seq) >= 0) break;
    __nanosleep(32);
    if ((++spins & 0xfffu) == 0 && p.timeout_ns != 0) {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > p.timeout_ns) {
        if (blockIdx.x == 0) {
          printf("fino halo_exchange timeout: rank %d gave up waiting for the rank %s at exchange %u (saw %u)\n", p.rank,
                 which ? "below" : "above", seq, v);
          reinterpret_cast<uint32_t*>(p.own)[4 + which] = seq ? seq : 1u;
        }
        break;
      }
    }
  } while (true);
}

__global__ void __launch_bounds__(256) halo_exchange_kernel(const __grid_constant__ HaloParams p) {
  const long long cpr = p.row_bytes / 16, total = (long long)p.t * cpr;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const int par = (int)(p.seq & 1u);
  uint32_t* own_words = reinterpret_cast<uint32_t*>(p.own);
  // (1) push
  if (p.up != nullptr) {
    int4* dst = reinterpret_cast<int4*>(p.up + HALO_DATA_OFF + (par * 2 + 1) * p.slot_bytes);
    for (long long i = tid; i < total; i += nth) {
      const long long f = i / cpr, c = i - f * cpr;
      dst[i] = *reinterpret_cast<const int4*>(p.frames + f * p.frame_stride + p.row_bytes + c * 16);
    }
  }
  if (p.down != nullptr) {
    int4* dst = reinterpret_cast<int4*>(p.down + HALO_DATA_OFF + (par * 2 + 0) * p.slot_bytes);
    for (long long i = tid; i < total; i += nth) {
      const long long f = i / cpr, c = i - f * cpr;
      dst[i] = *reinterpret_cast<const int4*>(p.frames + f * p.frame_stride + (long long)p.hl * p.row_bytes + c * 16);
    }
  }
  // (2) last CTA signals
  __threadfence_system();
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(own_words + 2, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    own_words[2] = 0u;
    __threadfence_system();
    if (p.up != nullptr)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(reinterpret_cast<uint32_t*>(p.up) + 1), "r"(p.seq) : "memory");
    if (p.down != nullptr)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(reinterpret_cast<uint32_t*>(p.down) + 0), "r"(p.seq) : "memory");
  }
  // (3) wait for the neighbours' rows
  if (threadIdx.x == 0) {
    if (p.up != nullptr) halo_wait(own_words + 0, p.seq, p, 0);
    if (p.down != nullptr) halo_wait(own_words + 1, p.seq, p, 1);
    __threadfence_system();
  }
  __syncthreads();
  // (4) pull (L2 is the coherence point of the peer's stores: bypass L1)
  if (p.up != nullptr) {
    const int4* src = reinterpret_cast<const int4*>(p.own + HALO_DATA_OFF + (par * 2 + 0) * p.slot_bytes);
    for (long long i = tid; i < total; i += nth) {
      const long long f = i / cpr, c = i - f * cpr;
      *reinterpret_cast<int4*>(p.frames + f * p.frame_stride + c * 16) = __ldcg(src + i);
    }
  }
  if (p.down != nullptr) {
    const int4* src = reinterpret_cast<const int4*>(p.own + HALO_DATA_OFF + (par * 2 + 1) * p.slot_bytes);
    for (long long i = tid; i < total; i += nth) {
      const long long f = i / cpr, c = i - f * cpr;
      *reinterpret_cast<int4*>(p.frames + f * p.frame_stride + (long long)(p.hl + 1) * p.row_bytes + c * 16) = __ldcg(src + i);
    }
  }
}

int halo_exchange(void* frames, int t, int hl, int64_t row_bytes, int64_t frame_stride_bytes, void* up, void* down,
                  void* own, uint32_t seq, int64_t slot_bytes, int rank, cudaStream_t stream) {
  FINO_CHECK_ARG(frames != nullptr && own != nullptr && t > 0 && hl > 0, "halo_exchange: bad arguments");
  FINO_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0 && frame_stride_bytes % 16 == 0 &&
                 frame_stride_bytes >= (int64_t)(hl + 2) * row_bytes, "halo_exchange: rows must be multiples of 16 bytes "
                 "and frames hold hl + 2 rows");
  FINO_CHECK_ARG(((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(own) | reinterpret_cast<uintptr_t>(up) |
                   reinterpret_cast<uintptr_t>(down)) & 15) == 0, "halo_exchange: pointers must be 16-byte aligned");
  FINO_CHECK_ARG(slot_bytes % 16 == 0 && (int64_t)t * row_bytes <= slot_bytes,
                 "halo_exchange: %lld bytes of rows do not fit a %lld-byte mailbox slot", (long long)t * row_bytes,
                 (long long)slot_bytes);
  if (up == nullptr && down == nullptr) return FINO_OK;
  HaloParams p;
  p.frames = reinterpret_cast<char*>(frames);
  p.t = t, p.hl = hl;
  p.row_bytes = row_bytes, p.frame_stride = frame_stride_bytes;
  p.up = reinterpret_cast<char*>(up), p.down = reinterpret_cast<char*>(down), p.own = reinterpret_cast<char*>(own);
  p.seq = seq;
  p.slot_bytes = slot_bytes;
  p.timeout_ns = peer_timeout_ns();
  p.rank = rank;
  const long long chunks = (long long)t * (row_bytes / 16);
  int grid = (int)((chunks + 256 * 4 - 1) / (256 * 4));
  grid = grid < 1 ? 1 : (grid > HALO_MAX_CTAS ? HALO_MAX_CTAS : grid);
  halo_exchange_kernel<<<grid, 256, 0, stream>>>(p);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

// ------------------------------------------------------------------------------------------------
// q/k RMSNorm-across-heads + RoPE, with the stores scattered to the ranks that own each head group
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float rbf_f(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ void unpack8_f(const uint4& u, float* f) {
  f[0] = bf16_lo_to_f32(u.x);
  f[1] = bf16_hi_to_f32(u.x);
  f[2] = bf16_lo_to_f32(u.y);
  f[3] = bf16_hi_to_f32(u.y);
  f[4] = bf16_lo_to_f32(u.z);
  f[5] = bf16_hi_to_f32(u.z);
  f[6] = bf16_lo_to_f32(u.w);
  f[7] = bf16_hi_to_f32(u.w);
}
__device__ __forceinline__ uint4 pack8_f(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

struct ScatterParams {
  const __nv_bfloat16* qkv;  // local [rows, row_stride]: q at column 0, k at dim, v at 2*dim
  int64_t rows, row_stride;
  const __nv_bfloat16* wq;
  const __nv_bfloat16* wk;
  int heads, head_dim;
  float eps;
  const float* cos;  // [rows, head_dim] fp32 rows of the LOCAL tokens (null: no RoPE)
  const float* sin;
  __nv_bfloat16* dst[PEER_MAX_RANKS];  // rank g's [world*rows_per_rank, dst_row_stride] buffer (q | k | v, inner each)
  int world, rank;
  int64_t rows_per_rank;   // row of local token t in every destination: rank*rows_per_rank + t
  int64_t dst_row_stride;  // elements (>= 3*inner)
  int inner;               // dim / world
};

constexpr int SCATTER_TOKENS_PER_BLOCK = 8;

// Thread c owns 16-byte chunk c (8 columns) of the q, k and v rows of each token the CTA walks over: the RMSNorm
// weights of its columns stay in registers, cos/sin are fetched once per token and used for q and k, and the two
// sum-of-squares reductions share one __syncthreads. All 8 columns of a chunk belong to one head, hence one rank.
__global__ void __launch_bounds__(512) qkv_norm_rope_scatter_kernel(const __grid_constant__ ScatterParams p) {
  __shared__ float red[2][2][16];
  const int c = threadIdx.x;
  const int lane = c & 31, warp = c >> 5;
  const int nwarps = blockDim.x >> 5;
  const int dim = p.heads * p.head_dim;
  const bool active = c < (dim >> 3);
  const float inv_dim = 1.0f / (float)dim;
  float wq[8], wk[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) wq[e] = wk[e] = 1.f;
  if (active && p.wq != nullptr) unpack8_f(__ldg(reinterpret_cast<const uint4*>(p.wq) + c), wq);
  if (active && p.wk != nullptr) unpack8_f(__ldg(reinterpret_cast<const uint4*>(p.wk) + c), wk);
  const int col = c * 8;
  const int off = col % p.head_dim;
  const int g = active ? col / p.inner : 0;  // destination rank of this chunk
  __nv_bfloat16* dst_base = p.dst[g] + (int64_t)p.rank * p.rows_per_rank * p.dst_row_stride + (col - g * p.inner);
  const int64_t t0 = (int64_t)blockIdx.x * SCATTER_TOKENS_PER_BLOCK;
  const int64_t t1 = min(p.rows, t0 + (int64_t)SCATTER_TOKENS_PER_BLOCK);
  uint4 uq = make_uint4(0, 0, 0, 0), uk = uq, uv = uq;
  if (active && t0 < t1) {
    const uint4* src = reinterpret_cast<const uint4*>(p.qkv + t0 * p.row_stride);
    uq = src[c];
    uk = src[c + (dim >> 3)];
    uv = src[c + 2 * (dim >> 3)];
  }
  int it = 0;
  for (int64_t tok = t0; tok < t1; ++tok, ++it) {
    float q[8], k[8];
    unpack8_f(uq, q);
    unpack8_f(uk, k);
    const uint4 v_now = uv;
    if (active && tok + 1 < t1) {  // prefetch the next token's rows
      const uint4* src = reinterpret_cast<const uint4*>(p.qkv + (tok + 1) * p.row_stride);
      uq = src[c];
      uk = src[c + (dim >> 3)];
      uv = src[c + 2 * (dim >> 3)];
    }
    float cs[8], sn[8];
    if (p.cos != nullptr && active) {
      const float4* cp = reinterpret_cast<const float4*>(p.cos + tok * p.head_dim + off);
      const float4* sp = reinterpret_cast<const float4*>(p.sin + tok * p.head_dim + off);
      const float4 a = __ldg(cp), b = __ldg(cp + 1), s0 = __ldg(sp), s1 = __ldg(sp + 1);
      cs[0] = a.x, cs[1] = a.y, cs[2] = a.z, cs[3] = a.w, cs[4] = b.x, cs[5] = b.y, cs[6] = b.z, cs[7] = b.w;
      sn[0] = s0.x, sn[1] = s0.y, sn[2] = s0.z, sn[3] = s0.w, sn[4] = s1.x, sn[5] = s1.y, sn[6] = s1.z, sn[7] = s1.w;
    }
    float sq = 0.f, sk = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sq += q[e] * q[e];
      sk += k[e] * k[e];
    }
    sq = warp_sum_f(sq);
    sk = warp_sum_f(sk);
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
      // RMSNorm (upstream diffusers): fp32 x*rsqrt -> bf16 -> * weight in bf16
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        q[e] = p.wq ? rbf_f(rbf_f(q[e] * rq) * wq[e]) : rbf_f(q[e] * rq);
        k[e] = p.wk ? rbf_f(rbf_f(k[e] * rk) * wk[e]) : rbf_f(k[e] * rk);
      }
      uint4* drow = reinterpret_cast<uint4*>(dst_base + tok * p.dst_row_stride);
      const int ichunks = p.inner >> 3;
      if (p.cos != nullptr) {
        // cos = freqs_cos[..., 0::2], sin = freqs_sin[..., 1::2]   (transformer_wan.py:83-87)
        float oq[8], ok[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          oq[e] = __fsub_rn(__fmul_rn(q[e], cs[e]), __fmul_rn(q[e + 1], sn[e + 1]));
          oq[e + 1] = __fadd_rn(__fmul_rn(q[e], sn[e + 1]), __fmul_rn(q[e + 1], cs[e]));
          ok[e] = __fsub_rn(__fmul_rn(k[e], cs[e]), __fmul_rn(k[e + 1], sn[e + 1]));
          ok[e + 1] = __fadd_rn(__fmul_rn(k[e], sn[e + 1]), __fmul_rn(k[e + 1], cs[e]));
        }
        drow[0] = pack8_f(oq);
        drow[ichunks] = pack8_f(ok);
      } else {
        drow[0] = pack8_f(q);
        drow[ichunks] = pack8_f(k);
      }
      drow[2 * ichunks] = v_now;
    }
  }
}

// Packed warp-per-row form of the kernel above (see ln_modulate_kernel2 / qk_rms_rope_kernel2 in norm_kernels.cu for
// the layout): head_dim == 128, GPL == heads, HPR == heads per rank. A lane owns 4 columns of every head, so group i
// of a row IS head i: its destination rank (i / HPR) and column inside that rank's buffer are compile-time constants
// after unrolling, one float4 of cos / sin per token serves all heads, and each warp-wide store writes one head's 256
// contiguous bytes into the owner's buffer. CTA x handles q (x % 3 == 0), k (1) or v (2, a plain copy) rows.
template <int GPL, int HPR>
__global__ void __launch_bounds__(256, (GPL <= 24 ? 2 : 1))
qkv_norm_rope_scatter_kernel2(const __grid_constant__ ScatterParams p) {
  constexpr int dim = GPL * 128;
  constexpr int inner = HPR * 128;
  const int lane = threadIdx.x & 31;
  const int which = blockIdx.x % 3;
  const int64_t row = (int64_t)(blockIdx.x / 3) * 8 + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const uint2* src = reinterpret_cast<const uint2*>(p.qkv + row * p.row_stride + which * dim) + lane;
  uint2 v[GPL];
#pragma unroll
  for (int i = 0; i < GPL; ++i)
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v[i].x), "=r"(v[i].y) : "l"(src + i * 32));
  const int64_t drow_off = ((int64_t)p.rank * p.rows_per_rank + row) * p.dst_row_stride + which * inner + lane * 4;
  if (which == 2) {
#pragma unroll
    for (int i = 0; i < GPL; ++i)
      *reinterpret_cast<uint2*>(p.dst[i / HPR] + drow_off + (i % HPR) * 128) = v[i];
    return;
  }
  const bool rope = p.cos != nullptr;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), sn = cs;
  if (rope) {
    cs = __ldg(reinterpret_cast<const float4*>(p.cos + row * 128) + lane);
    sn = __ldg(reinterpret_cast<const float4*>(p.sin + row * 128) + lane);
  }
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
  for (int i = 0; i < GPL; ++i) {
    const float a0 = bf16_lo_to_f32(v[i].x), a1 = bf16_hi_to_f32(v[i].x);
    const float a2 = bf16_lo_to_f32(v[i].y), a3 = bf16_hi_to_f32(v[i].y);
    ffma2(q0, q1, a0, a1, a0, a1, q0, q1);
    ffma2(q2, q3, a2, a3, a2, a3, q2, q3);
  }
  const float rstd = rsqrtf(warp_sum_f((q0 + q1) + (q2 + q3)) * (1.0f / (float)dim) + p.eps);
  const __nv_bfloat16* wt = which ? p.wk : p.wq;
  const uint2* wr = reinterpret_cast<const uint2*>(wt) + lane;
  const float c01 = cs.x, s01 = sn.y, c23 = cs.z, s23 = sn.w;
#pragma unroll
  for (int i = 0; i < GPL; ++i) {
    float a0, a1, a2, a3;
    fmul2(a0, a1, bf16_lo_to_f32(v[i].x), bf16_hi_to_f32(v[i].x), rstd, rstd);
    fmul2(a2, a3, bf16_lo_to_f32(v[i].y), bf16_hi_to_f32(v[i].y), rstd, rstd);
    uint32_t n01 = pack_bf16x2(a0, a1), n23 = pack_bf16x2(a2, a3);
    if (wt != nullptr) {
      const uint2 w = __ldg(wr + i * 32);
      fmul2(a0, a1, bf16_lo_to_f32(n01), bf16_hi_to_f32(n01), bf16_lo_to_f32(w.x), bf16_hi_to_f32(w.x));
      fmul2(a2, a3, bf16_lo_to_f32(n23), bf16_hi_to_f32(n23), bf16_lo_to_f32(w.y), bf16_hi_to_f32(w.y));
      n01 = pack_bf16x2(a0, a1);
      n23 = pack_bf16x2(a2, a3);
    }
    if (rope) {
      const float x0 = bf16_lo_to_f32(n01), x1 = bf16_hi_to_f32(n01);
      const float x2 = bf16_lo_to_f32(n23), x3 = bf16_hi_to_f32(n23);
      float pa, pb, ra, rb, o0, o1, o2, o3;
      fmul2(pa, pb, x0, x1, c01, c01);
      fmul2(ra, rb, x1, x0, -s01, s01);
      fadd2(o0, o1, pa, pb, ra, rb);
      fmul2(pa, pb, x2, x3, c23, c23);
      fmul2(ra, rb, x3, x2, -s23, s23);
      fadd2(o2, o3, pa, pb, ra, rb);
      n01 = pack_bf16x2(o0, o1);
      n23 = pack_bf16x2(o2, o3);
    }
    *reinterpret_cast<uint2*>(p.dst[i / HPR] + drow_off + (i % HPR) * 128) = make_uint2(n01, n23);
  }
}

// ------------------------------------------------------------------------------------------------
// CogVideoX prologue of the fused exchange: per-head LayerNorm(64) (+ CogVideoX RoPE on the video rows) of the local
// fused projection rows, stored straight into the owning ranks' buffers (attention_processor.py:2848-2860 followed by
// the first Ulysses all-to-all). Same lane layout as qk_ln64_rope_kernel (norm_kernels.cu): lane l holds the 16-byte
// chunk i*32 + l of every 256-column group i, a 64-wide head is 8 consecutive lanes of one group, so the chunk's
// destination rank (head / heads_per_rank) and its column inside that rank's (q | k | v) row are cheap integer math
// and every 8-lane group writes one head's 128 contiguous bytes. CTA x handles q (x % 3 == 0), k (1) or v (2: copy).
// ------------------------------------------------------------------------------------------------
struct LnScatterParams {
  const __nv_bfloat16* qkv;
  int64_t rows, row_stride;
  const __nv_bfloat16 *wq, *bq, *wk, *bk;  // [64] each (null: no affine)
  int heads;
  float eps;
  const float* cos;  // [rows - rope_skip, 64] fp32 rows of the LOCAL video tokens (null: no RoPE)
  const float* sin;
  int64_t rope_skip;  // leading local rows that are text (not rotated)
  __nv_bfloat16* dst[PEER_MAX_RANKS];
  int world, rank;
  int64_t rows_per_rank, dst_row_stride;
  int inner;  // heads * 64 / world
};

template <int CPL, int G>
__global__ void __launch_bounds__(256, 3) qkv_ln64_rope_scatter_kernel(const __grid_constant__ LnScatterParams p) {
  constexpr int dim = CPL * 256;
  const int lane = threadIdx.x & 31;
  const int which = blockIdx.x % 3;
  const int64_t row = (int64_t)(blockIdx.x / 3) * 8 + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const uint4* src = reinterpret_cast<const uint4*>(p.qkv + row * p.row_stride + which * dim) + lane;
  const int64_t drow = ((int64_t)p.rank * p.rows_per_rank + row) * p.dst_row_stride + which * p.inner;
  const int hc = lane & 7;
  const bool do_rope = which < 2 && p.cos != nullptr && row >= p.rope_skip;
  uint4 wq = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u), bq = make_uint4(0u, 0u, 0u, 0u);
  float cs[8], sn[8];
  if (which < 2) {
    const __nv_bfloat16* wt = which ? p.wk : p.wq;
    const __nv_bfloat16* bt = which ? p.bk : p.bq;
    if (wt != nullptr) wq = __ldg(reinterpret_cast<const uint4*>(wt) + hc);
    if (bt != nullptr) bq = __ldg(reinterpret_cast<const uint4*>(bt) + hc);
    if (do_rope) {
      const float4* cp = reinterpret_cast<const float4*>(p.cos + (row - p.rope_skip) * 64 + hc * 8);
      const float4* sp = reinterpret_cast<const float4*>(p.sin + (row - p.rope_skip) * 64 + hc * 8);
      const float4 a = __ldg(cp), b = __ldg(cp + 1), s0 = __ldg(sp), s1 = __ldg(sp + 1);
      cs[0] = a.x, cs[1] = a.y, cs[2] = a.z, cs[3] = a.w, cs[4] = b.x, cs[5] = b.y, cs[6] = b.z, cs[7] = b.w;
      sn[0] = s0.x, sn[1] = s0.y, sn[2] = s0.z, sn[3] = s0.w, sn[4] = s1.x, sn[5] = s1.y, sn[6] = s1.z, sn[7] = s1.w;
    }
  }
#pragma unroll 1
  for (int g0 = 0; g0 < CPL; g0 += G) {
    uint4 raw[G];
#pragma unroll
    for (int i = 0; i < G; ++i) raw[i] = src[(g0 + i) * 32];
    float mean[G], rstd[G];
    if (which < 2) {
#pragma unroll
      for (int i = 0; i < G; ++i) {
        float v[8];
        unpack8_f(raw[i], v);
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
        unpack8_f(raw[i], v);
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
    }
    float w[8], b[8];
    unpack8_f(wq, w);
    unpack8_f(bq, b);
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int col = (g0 + i) * 256 + lane * 8;
      const int g = col / p.inner;  // destination rank (a chunk never straddles a head, hence never a rank)
      uint4* d16 = reinterpret_cast<uint4*>(p.dst[g] + drow + (col - g * p.inner));
      if (which == 2) {
        *d16 = raw[i];
        continue;
      }
      const float r = rsqrtf(rstd[i] * (1.0f / 64.0f) + p.eps);
      float v[8];
      unpack8_f(raw[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = rbf_f((v[e] - mean[i]) * r * w[e] + b[e]);
      if (do_rope) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float xe = v[e], xo = v[e + 1];
          o[e] = __fadd_rn(__fmul_rn(xe, cs[e]), __fmul_rn(-xo, sn[e]));
          o[e + 1] = __fadd_rn(__fmul_rn(xo, cs[e + 1]), __fmul_rn(xe, sn[e + 1]));
        }
        *d16 = pack8_f(o);
      } else {
        *d16 = pack8_f(v);
      }
    }
  }
}

int qkv_ln_rope_scatter(const void* qkv, int64_t rows, int64_t row_stride, const void* wq, const void* bq, const void* wk,
                        const void* bk, int heads, int head_dim, float eps, const float* cos, const float* sin,
                        int64_t rope_skip, void* const* dst_ptrs, int world, int rank, int64_t rows_per_rank,
                        int64_t dst_row_stride, cudaStream_t stream) {
  FINO_CHECK_ARG(qkv != nullptr && rows > 0 && dst_ptrs != nullptr, "qkv_ln_rope_scatter: null / empty input");
  FINO_CHECK_ARG(world >= 1 && world <= PEER_MAX_RANKS && rank >= 0 && rank < world,
                 "qkv_ln_rope_scatter: bad rank/world (%d/%d)", rank, world);
  FINO_CHECK_ARG(head_dim == 64 && heads > 0 && heads % world == 0,
                 "qkv_ln_rope_scatter: head_dim 64 and heads divisible by the world size (%d heads x %d over %d)", heads,
                 head_dim, world);
  const int dim = heads * 64;
  FINO_CHECK_ARG(dim == 3072 || dim == 256 || dim == 512, "qkv_ln_rope_scatter: heads*64 = %d (built for 256 / 512 / 3072)", dim);
  FINO_CHECK_ARG(row_stride % 8 == 0 && row_stride >= 3 * dim, "qkv_ln_rope_scatter: row stride");
  const int inner = dim / world;
  FINO_CHECK_ARG(dst_row_stride % 8 == 0 && dst_row_stride >= 3 * inner, "qkv_ln_rope_scatter: dst row stride");
  FINO_CHECK_ARG(rows <= rows_per_rank, "qkv_ln_rope_scatter: more local rows than rows_per_rank");
  FINO_CHECK_ARG((cos == nullptr) == (sin == nullptr) && rope_skip >= 0, "qkv_ln_rope_scatter: cos / sin / rope_skip");
  LnScatterParams p;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  p.rows = rows;
  p.row_stride = row_stride;
  p.wq = reinterpret_cast<const __nv_bfloat16*>(wq);
  p.bq = reinterpret_cast<const __nv_bfloat16*>(bq);
  p.wk = reinterpret_cast<const __nv_bfloat16*>(wk);
  p.bk = reinterpret_cast<const __nv_bfloat16*>(bk);
  p.heads = heads;
  p.eps = eps;
  p.cos = cos;
  p.sin = sin;
  p.rope_skip = rope_skip;
  for (int r = 0; r < PEER_MAX_RANKS; ++r) {
    FINO_CHECK_ARG(r >= world || (dst_ptrs[r] != nullptr && (reinterpret_cast<uintptr_t>(dst_ptrs[r]) & 15) == 0),
                   "qkv_ln_rope_scatter: dst %d null or unaligned", r);
    p.dst[r] = reinterpret_cast<__nv_bfloat16*>(dst_ptrs[r < world ? r : 0]);
  }
  p.world = world;
  p.rank = rank;
  p.rows_per_rank = rows_per_rank;
  p.dst_row_stride = dst_row_stride;
  p.inner = inner;
  dim3 grid((unsigned)(3 * ((rows + 7) / 8)));
  if (dim == 3072) qkv_ln64_rope_scatter_kernel<12, 4><<<grid, 256, 0, stream>>>(p);
  else if (dim == 512) qkv_ln64_rope_scatter_kernel<2, 2><<<grid, 256, 0, stream>>>(p);
  else qkv_ln64_rope_scatter_kernel<1, 1><<<grid, 256, 0, stream>>>(p);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

template <int GPL, int HPR>
static void launch_scatter2(const ScatterParams& p, cudaStream_t stream) {
  dim3 grid((unsigned)(3 * ((p.rows + 7) / 8)));
  qkv_norm_rope_scatter_kernel2<GPL, HPR><<<grid, 256, 0, stream>>>(p);
}

bool qk_tma_eligible(int64_t rows, int heads, int head_dim);
int qkv_rms_rope_scatter_tma(const void* qkv, int64_t rows, int64_t row_stride, const void* wq, const void* wk,
                             int heads, int head_dim, float eps, const float* cos, const float* sin,
                             void* const* dst_ptrs, int world, int rank, int64_t rows_per_rank,
                             int64_t dst_row_stride, cudaStream_t stream);
static bool g_scatter_tma = false;  // see g_rows_tma in norm_kernels.cu
static bool g_scatter_packed = true;  // packed warp-per-row kernel for the Wan shapes (24 / 40 heads x 128)
void scatter_set_tma(int on) { g_scatter_tma = on != 0; }
void scatter_set_packed(int on) { g_scatter_packed = on != 0; }  // follows the q/k variant of fino_rows_set_variant

int qkv_norm_rope_scatter(const void* qkv, int64_t rows, int64_t row_stride, const void* wq, const void* wk,
                          int heads, int head_dim, float eps, const float* cos, const float* sin,
                          void* const* dst_ptrs, int world, int rank, int64_t rows_per_rank, int64_t dst_row_stride,
                          cudaStream_t stream) {
  FINO_CHECK_ARG(qkv != nullptr && rows > 0 && dst_ptrs != nullptr, "qkv_norm_rope_scatter: null / empty input");
  FINO_CHECK_ARG(world >= 1 && world <= PEER_MAX_RANKS && rank >= 0 && rank < world,
                 "qkv_norm_rope_scatter: bad rank/world (%d/%d)", rank, world);
  FINO_CHECK_ARG(heads > 0 && head_dim >= 8 && head_dim % 8 == 0 && heads % world == 0,
                 "qkv_norm_rope_scatter: heads %d x head_dim %d not divisible over %d ranks", heads, head_dim, world);
  const int dim = heads * head_dim;
  FINO_CHECK_ARG(row_stride % 8 == 0 && row_stride >= 3 * dim, "qkv_norm_rope_scatter: row stride");
  const int inner = dim / world;
  FINO_CHECK_ARG(dst_row_stride % 8 == 0 && dst_row_stride >= 3 * inner, "qkv_norm_rope_scatter: dst row stride");
  FINO_CHECK_ARG(rows <= rows_per_rank, "qkv_norm_rope_scatter: more local rows than rows_per_rank");
  FINO_CHECK_ARG((cos == nullptr) == (sin == nullptr), "qkv_norm_rope_scatter: cos and sin go together");
  for (int r = 0; r < world; ++r)
    FINO_CHECK_ARG(dst_ptrs[r] != nullptr && (reinterpret_cast<uintptr_t>(dst_ptrs[r]) & 15) == 0,
                   "qkv_norm_rope_scatter: dst %d null or unaligned", r);
  if (g_scatter_tma && qk_tma_eligible(rows, heads, head_dim) && head_dim % 4 == 0 &&
      ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(cos) | reinterpret_cast<uintptr_t>(sin)) & 15) == 0)
    return qkv_rms_rope_scatter_tma(qkv, rows, row_stride, wq, wk, heads, head_dim, eps, cos, sin, dst_ptrs, world, rank,
                                    rows_per_rank, dst_row_stride, stream);
  ScatterParams p;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  p.rows = rows;
  p.row_stride = row_stride;
  p.wq = reinterpret_cast<const __nv_bfloat16*>(wq);
  p.wk = reinterpret_cast<const __nv_bfloat16*>(wk);
  p.heads = heads;
  p.head_dim = head_dim;
  p.eps = eps;
  p.cos = cos;
  p.sin = sin;
  for (int r = 0; r < PEER_MAX_RANKS; ++r) {
    void* d = dst_ptrs[r < world ? r : 0];
    FINO_CHECK_ARG(d != nullptr && (reinterpret_cast<uintptr_t>(d) & 15) == 0, "qkv_norm_rope_scatter: dst %d", r);
    p.dst[r] = reinterpret_cast<__nv_bfloat16*>(d);
  }
  p.world = world;
  p.rank = rank;
  p.rows_per_rank = rows_per_rank;
  p.dst_row_stride = dst_row_stride;
  p.inner = inner;
  if (g_scatter_packed && head_dim == 128 && (heads == 24 || heads == 40) && (world == 2 || world == 4 || world == 8)) {
    const int hpr = heads / world;
    if (heads == 24) {
      if (hpr == 12) launch_scatter2<24, 12>(p, stream);
      else if (hpr == 6) launch_scatter2<24, 6>(p, stream);
      else launch_scatter2<24, 3>(p, stream);
    } else {
      if (hpr == 20) launch_scatter2<40, 20>(p, stream);
      else if (hpr == 10) launch_scatter2<40, 10>(p, stream);
      else launch_scatter2<40, 5>(p, stream);
    }
    FINO_CHECK_CUDA(cudaGetLastError());
    return FINO_OK;
  }
  FINO_CHECK_ARG(dim / 8 <= 512, "qkv_norm_rope_scatter: heads*head_dim <= 4096 (or 24 / 40 heads x 128)");
  const int threads = ((dim / 8 + 31) / 32) * 32;
  dim3 grid((unsigned)((rows + SCATTER_TOKENS_PER_BLOCK - 1) / SCATTER_TOKENS_PER_BLOCK));
  qkv_norm_rope_scatter_kernel<<<grid, threads, 0, stream>>>(p);
  FINO_CHECK_CUDA(cudaGetLastError());
  return FINO_OK;
}

}  // namespace fino
