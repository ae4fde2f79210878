#include "common.cuh"
#include <stdarg.h>
#include <mutex>

namespace fino {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_last_error; }

typedef CUresult (*PFN_cuTensorMapEncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                                 const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_cuTensorMapEncodeTiled_t get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_t fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      set_last_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_t>(p);
  }
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides) {
  PFN_cuTensorMapEncodeTiled_t fn = get_encode_fn();
  if (!fn) return FINO_ERR_CUDA;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_last_error("TMA base pointer %p is not 16-byte aligned", base);
    return FINO_ERR_INVALID;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    if (gstr[i] % 16 != 0) {
      set_last_error("TMA global stride %llu (dim %d) is not a multiple of 16 bytes", (unsigned long long)gstr[i],
                     i + 1);
      return FINO_ERR_INVALID;
    }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
                   rank > 1 ? box[1] : 0);
    return FINO_ERR_CUDA;
  }
  return FINO_OK;
}

struct WsEntry {
  int dev, tag;
  cudaStream_t stream;
  void* ptr;
  size_t bytes;
};
static WsEntry g_ws[64];
static int g_ws_n = 0;
static std::mutex g_ws_mu;

int stream_workspace(int tag, cudaStream_t stream, size_t bytes, void** out) {
  std::lock_guard<std::mutex> lock(g_ws_mu);
  int dev = 0;
  FINO_CHECK_CUDA(cudaGetDevice(&dev));
  WsEntry* e = nullptr;
  for (int i = 0; i < g_ws_n; ++i)
    if (g_ws[i].dev == dev && g_ws[i].tag == tag && g_ws[i].stream == stream) e = &g_ws[i];
  if (e == nullptr) {
    FINO_CHECK_ARG(g_ws_n < 64, "stream_workspace: more than 64 (device, stream, kind) combinations in use");
    e = &g_ws[g_ws_n++];
    *e = WsEntry{dev, tag, stream, nullptr, 0};
  }
  if (e->bytes < bytes) {
    if (e->ptr) {
      FINO_CHECK_CUDA(cudaStreamSynchronize(stream));  // earlier launches on this stream may still read the old block
      FINO_CHECK_CUDA(cudaFree(e->ptr));
      e->ptr = nullptr;
      e->bytes = 0;
    }
    const size_t want = bytes < ((size_t)20 << 20) ? ((size_t)20 << 20) : bytes;
    FINO_CHECK_CUDA(cudaMalloc(&e->ptr, want));
    e->bytes = want;
  }
  *out = e->ptr;
  return FINO_OK;
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

int num_sms() {
  static int n[64] = {};
  const int dev = current_device() & 63;
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

}  // namespace fino
