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

// 3-D fp16 map [dim2][dim1][inner], box {box_inner, box_rows, 1}; box_inner = 64 -> 128B swizzle, 32 -> 64B swizzle
// (per-sequence tiles: OOB rows of a sequence are clipped on store / zero-filled on load).
int make_tmap_f16_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t dim1, uint64_t dim2,
                     uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box_inner, uint32_t box_rows);

// 4-D fp16 map over an NHWC activation [n][h][w][channels] (pixel stride in bytes >= 2 * channels), box
// {64 channels, box_w, box_h, box_n}, 128B swizzle: a rectangular pixel patch lands in shared memory as box_w*box_h*box_n
// consecutive 128-byte rows (x fastest); coordinates outside the image are zero-filled on load and clipped on store.
int make_tmap_f16_nhwc(CUtensorMap* out, const void* base, uint64_t channels, uint64_t w, uint64_t h, uint64_t n,
                       uint64_t pixel_stride_bytes, uint32_t box_w, uint32_t box_h, uint32_t box_n);

int device_sm_count();  // SM count of the CURRENT device (cached per device)

// The opt-in dynamic shared-memory limit (cudaFuncSetAttribute) is per device and a process may hold one pc_ctx per
// GPU: remember the largest size configured for (kernel, device) so that the attribute call stays off the launch path.
// `slots` is a zero-initialised static array of kMaxDevices atomics owned by the call site.
constexpr int kMaxDevices = 64;
int current_device();
template <typename Kern>
cudaError_t ensure_dynamic_smem(Kern kern, int bytes, int* slots) {
  const int dev = current_device();
  int* slot = &slots[dev < 0 || dev >= kMaxDevices ? 0 : dev];
  if (bytes <= __atomic_load_n(slot, __ATOMIC_ACQUIRE) && dev >= 0 && dev < kMaxDevices) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) __atomic_store_n(slot, bytes, __ATOMIC_RELEASE);
  return e;
}

// Programmatic dependent launch (PDL): kernels launched through launch_pdl() may be scheduled while the previous
// kernel of the stream is still draining; they run their prologue (barrier init, TMEM allocation, descriptor
// prefetch) and then block in griddep_wait() until the predecessor has completed and flushed its writes. Every
// kernel launched this way MUST call griddep_wait() before its first global-memory access, and calls
// griddep_launch_dependents() at its start so that its own successor can be scheduled early.
// PC_NO_PDL=1 in the environment turns the launch attribute off (plain stream order) for A/B timing.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                       Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace pc
