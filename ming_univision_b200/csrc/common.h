// Host-side helpers shared by the C-ABI translation units (error reporting, TMA descriptor encoding).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mingb200.h"

namespace mb {

// Thread-local last-error string returned by mb_last_error().
void set_error(const char* fmt, ...);

#define MB_CHECK_ARG(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      ::mb::set_error(__VA_ARGS__);        \
      return (code);                       \
    }                                      \
  } while (0)

#define MB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::mb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MB_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

// Encodes a 2-D bf16 row-major tensor map: dims {inner, outer}, row stride in elements, box {box_inner, box_outer},
// 128-byte swizzle, zero fill for out-of-bounds elements. Returns false (and sets the error) on failure.
bool make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                       uint32_t box_inner, uint32_t box_outer);

// 4-D bf16 view {d0 (contiguous), d1, d2, d3} with element strides {1, s1, s2, s3} and box {b0, b1, b2, b3}; 128-byte
// swizzle (b0 * 2 bytes <= 128), zero fill out of bounds.  Used for [batch][token][head][head_dim] attention operands.
bool make_tmap_4d_bf16(CUtensorMap* map, const void* base, const uint64_t dims[4], const uint64_t strides_elems[3],
                       const uint32_t box[4]);

int num_sms();

// Launch with the programmatic-stream-serialization attribute (PDL).  Only for kernels that call pdl_wait() before
// their first dependent memory access.  MB_NO_PDL=1 in the environment falls back to a plain launch.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace mb
