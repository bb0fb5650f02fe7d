// Softmax arithmetic shared by the two attention kernels (attention.cu, attention5.cu): exp2 on the MUFU pipe, packed
// fp32 pair math (one issue slot for two elements), and the per-16-column chunk helpers of the online softmax.
#pragma once
#include "ptx.cuh"

namespace pc {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  return static_cast<uint64_t>(__float_as_uint(lo)) | (static_cast<uint64_t>(__float_as_uint(hi)) << 32);
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 16 S columns of this thread's row. kmax = last valid key of THIS row relative to the chunk's first key.
template <bool FULL>
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[16], int kmax, float mx) {
  float m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    float a = __uint_as_float(v[j]), b = __uint_as_float(v[j + 1]);
    float c = __uint_as_float(v[j + 2]), d = __uint_as_float(v[j + 3]);
    if (!FULL) {
      a = (j <= kmax) ? a : -INFINITY;
      b = (j + 1 <= kmax) ? b : -INFINITY;
      c = (j + 2 <= kmax) ? c : -INFINITY;
      d = (j + 3 <= kmax) ? d : -INFINITY;
    }
    mx = fmaxf(mx, fmaxf(a, b));
    m1 = fmaxf(m1, fmaxf(c, d));
  }
  return fmaxf(mx, m1);
}
template <bool FULL, bool NOEXP = false>
__device__ __forceinline__ uint64_t chunk_exp(const uint32_t (&v)[16], uint32_t (&pk)[8], int kmax, uint64_t sc2,
                                              uint64_t nref2, uint64_t acc) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const uint64_t t = fma_f32x2(static_cast<uint64_t>(v[j]) | (static_cast<uint64_t>(v[j + 1]) << 32), sc2, nref2);
    float e0 = NOEXP ? __uint_as_float(static_cast<uint32_t>(t)) : ex2_approx(__uint_as_float(static_cast<uint32_t>(t)));
    float e1 = NOEXP ? __uint_as_float(static_cast<uint32_t>(t >> 32)) : ex2_approx(__uint_as_float(static_cast<uint32_t>(t >> 32)));
    if (!FULL) {
      e0 = (j <= kmax) ? e0 : 0.0f;
      e1 = (j + 1 <= kmax) ? e1 : 0.0f;
    }
    acc = add_f32x2(acc, pack_f32x2(e0, e1));
    pk[j >> 1] = pack_half2(e0, e1);
  }
  return acc;
}

}  // namespace pc
