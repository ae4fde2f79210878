// Host-side helpers shared by the kernels' launchers: status codes, last-error text, TMA tensor-map encode.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace fino {

enum Status : int {
  FINO_OK = 0,
  FINO_ERR_INVALID = 1,   // bad argument (shape / alignment / null pointer)
  FINO_ERR_CUDA = 2,      // a CUDA runtime / driver call failed
  FINO_ERR_UNSUPPORTED = 3,
};

void set_last_error(const char* fmt, ...);
const char* get_last_error();

#define FINO_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      ::fino::set_last_error(__VA_ARGS__);   \
      return ::fino::FINO_ERR_INVALID;       \
    }                                        \
  } while (0)

#define FINO_CHECK_CUDA(expr)                                                                       \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ::fino::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ::fino::FINO_ERR_CUDA;                                                                 \
    }                                                                                               \
  } while (0)

// Encodes a bf16 tiled tensor map with 128-byte swizzle. dims/strides innermost first; strides in bytes for
// dims 1..rank-1. Returns 0 on success.
// elem_strides (optional): traversal stride per dimension (1 = every element); a box of `box[i]` elements then yields
// ceil(box[i] / elem_strides[i]) elements in shared memory. Out-of-bounds elements (also negative coordinates) read as 0.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides = nullptr);

// Library-owned scratch memory keyed by (current device, stream, tag): grown on demand, never shared between streams, so
// launches that use it on different streams (e.g. the side stream of prepare_text) cannot race on it. Launches on ONE
// stream reuse the same block (stream order protects it).
int stream_workspace(int tag, cudaStream_t stream, size_t bytes, void** out);

int num_sms();         // SM count of the CURRENT device (cached per device)
int current_device();  // cudaGetDevice

}  // namespace fino
