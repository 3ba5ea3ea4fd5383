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

int num_sms();

}  // namespace mb
