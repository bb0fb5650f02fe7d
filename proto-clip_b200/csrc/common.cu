#include "common.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <unordered_map>
#include <utility>
#include <vector>

namespace pc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

namespace {

struct TmapKey {
  const void* base;
  uint64_t inner, rows, stride;
  uint32_t box_inner, box_rows, elem_bytes;
  bool operator==(const TmapKey& o) const {
    return base == o.base && inner == o.inner && rows == o.rows && stride == o.stride && box_inner == o.box_inner &&
           box_rows == o.box_rows && elem_bytes == o.elem_bytes;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (k.inner + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.stride + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= ((static_cast<uint64_t>(k.box_inner) << 40 | static_cast<uint64_t>(k.box_rows) << 8 | k.elem_bytes) +
          0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    return static_cast<size_t>(h);
  }
};

}  // namespace

// Descriptors are pure functions of (address, geometry): the encoder re-uses the same activation buffers and
// weights on every call, so a small per-thread cache removes cuTensorMapEncodeTiled from the launch path.
int make_tmap_2d(CUtensorMap* out, const void* base, uint32_t elem_bytes, uint64_t inner, uint64_t rows,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  const TmapKey key{base, inner, rows, row_stride_bytes, box_inner, box_rows, elem_bytes};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return PC_OK;
  }
  EncodeTiledFn fn = get_encode_fn();
  PC_REQUIRE(fn != nullptr, PC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  PC_REQUIRE(elem_bytes == 2 || elem_bytes == 4, PC_ERR_ARG, "TMA element size %u unsupported", elem_bytes);
  PC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, PC_ERR_ALIGN,
             "TMA base %p is not 16-byte aligned", base);
  PC_REQUIRE((row_stride_bytes & 15) == 0, PC_ERR_ALIGN, "TMA row stride %llu is not a multiple of 16",
             (unsigned long long)row_stride_bytes);
  PC_REQUIRE(box_inner * elem_bytes == 128 && box_rows >= 1 && box_rows <= 256, PC_ERR_ARG,
             "TMA box %ux%u unsupported", box_inner, box_rows);
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PC_REQUIRE(r == CUDA_SUCCESS, PC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return PC_OK;
}

// 3-D fp16 map [dim2][dim1][inner] with a {box_inner, box_rows, 1} box (64 columns: 128B swizzle, 32: 64B swizzle): rows are clipped / zero-filled
// per dim-2 slice (used for per-sequence tiles: dim1 = tokens of one sequence, dim2 = sequences).
int make_tmap_f16_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t dim1, uint64_t dim2,
                     uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box_inner, uint32_t box_rows) {
  struct Key3 {
    const void* base;
    uint64_t inner, dim1, dim2, s1, s2;
    uint32_t box_inner, box_rows;
  };
  static thread_local std::vector<std::pair<Key3, CUtensorMap>> cache;
  for (const auto& e : cache) {
    const Key3& k = e.first;
    if (k.base == base && k.inner == inner && k.dim1 == dim1 && k.dim2 == dim2 && k.s1 == stride1_bytes &&
        k.s2 == stride2_bytes && k.box_inner == box_inner && k.box_rows == box_rows) {
      *out = e.second;
      return PC_OK;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  PC_REQUIRE(fn != nullptr, PC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  PC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (stride1_bytes & 15) == 0 && (stride2_bytes & 15) == 0,
             PC_ERR_ALIGN, "TMA 3-D map: base / strides must be multiples of 16 bytes");
  PC_REQUIRE(box_rows >= 1 && box_rows <= 256 && (box_inner == 64 || box_inner == 32), PC_ERR_ARG,
             "TMA box %ux%u unsupported", box_inner, box_rows);
  cuuint64_t dims[3] = {inner, dim1, dim2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PC_REQUIRE(r == CUDA_SUCCESS, PC_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
  if (cache.size() > 64) cache.clear();
  cache.emplace_back(Key3{base, inner, dim1, dim2, stride1_bytes, stride2_bytes, box_inner, box_rows}, *out);
  return PC_OK;
}

int make_tmap_f16_nhwc(CUtensorMap* out, const void* base, uint64_t channels, uint64_t w, uint64_t h, uint64_t n,
                       uint64_t pixel_stride_bytes, uint32_t box_w, uint32_t box_h, uint32_t box_n) {
  struct Key4 {
    const void* base;
    uint64_t c, w, h, n, ps;
    uint32_t bw, bh, bn;
  };
  static thread_local std::vector<std::pair<Key4, CUtensorMap>> cache;
  for (const auto& e : cache) {
    const Key4& k = e.first;
    if (k.base == base && k.c == channels && k.w == w && k.h == h && k.n == n && k.ps == pixel_stride_bytes &&
        k.bw == box_w && k.bh == box_h && k.bn == box_n) {
      *out = e.second;
      return PC_OK;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  PC_REQUIRE(fn != nullptr, PC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  PC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (pixel_stride_bytes & 15) == 0, PC_ERR_ALIGN,
             "TMA NHWC map: base / pixel stride must be multiples of 16 bytes");
  PC_REQUIRE(box_w >= 1 && box_h >= 1 && box_n >= 1 && box_w * box_h * box_n <= 256, PC_ERR_ARG,
             "TMA NHWC box %ux%ux%u unsupported", box_w, box_h, box_n);
  cuuint64_t dims[4] = {channels, w, h, n};
  cuuint64_t strides[3] = {pixel_stride_bytes, pixel_stride_bytes * w, pixel_stride_bytes * w * h};
  cuuint32_t box[4] = {64, box_w, box_h, box_n};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PC_REQUIRE(r == CUDA_SUCCESS, PC_ERR_CUDA, "cuTensorMapEncodeTiled (NHWC) failed with CUresult %d", (int)r);
  if (cache.size() > 256) cache.clear();
  cache.emplace_back(Key4{base, channels, w, h, n, pixel_stride_bytes, box_w, box_h, box_n}, *out);
  return PC_OK;
}

int make_tmap_f16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows,
                     uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows) {
  return make_tmap_2d(out, base, 2, inner, rows, row_stride_bytes, box_inner, box_rows);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PC_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  return dev;
}

int device_sm_count() {
  static int cache[kMaxDevices];  // 0 = unknown; racing writers store the same value
  const int dev = current_device();
  if (dev < 0) return 148;
  if (dev < kMaxDevices && cache[dev]) return cache[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev < kMaxDevices) cache[dev] = n;
  return n;
}

}  // namespace pc
