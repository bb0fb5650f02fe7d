// Host-side plumbing shared by every translation unit of libprotoclip_b200: error codes that cross the
// C ABI (include/protoclip_b200.h), the thread-local error string, CUDA error capture, and TMA tensor-map
// construction through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/protoclip_b200.h"

namespace pc {

// error codes: the PC_* enum of the public header (global namespace)

void set_error(const char* fmt, ...);
const char* get_error();

#define PC_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      pc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PC_ERR_CUDA;                                                                 \
    }                                                                                         \
  } while (0)

#define PC_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      pc::set_error(__VA_ARGS__);    \
      return (code);                 \
    }                                \
  } while (0)

#define PC_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != PC_OK) return _rc; \
  } while (0)

// 2-D fp16 tensor map: `inner` contiguous elements per row, `rows` rows, `row_stride_bytes` between rows.
// Box = box_inner x box_rows elements, 128B swizzle (box_inner must be 64 fp16 = 128 B).
int make_tmap_f16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows,
                     uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows);

// Generic form (elem_bytes 2 = fp16, 4 = fp32; box_inner * elem_bytes must be 128). Results are cached per thread.
int make_tmap_2d(CUtensorMap* out, const void* base, uint32_t elem_bytes, uint64_t inner, uint64_t rows,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows);

int device_sm_count();

}  // namespace pc
