// Micro-benchmarks that size the attention kernel's softmax stage on B200 (sm_100a): TMEM read / write
// throughput per SM (tcgen05.ld / tcgen05.st, 32x32b shapes), MUFU ex2 throughput, and the FMA-pipe cost of a
// polynomial exp2. One CTA per SM, W warps; cycles are clock64() deltas of warp 0 around a barrier pair.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e = (x);                                                            \
    if (e != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// mode 0: tcgen05.ld x32 stream; 1: tcgen05.st x16 stream; 2: ex2.approx stream; 3: ld + fmax (pass-1 shape);
// 4: ld + ffma + ex2 + add + cvt + st (pass-2 shape); 5: polynomial exp2 on the FMA pipe
__global__ void __launch_bounds__(512) ubench(int mode, int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_s)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 256;
  float acc = 0.0f;
  uint32_t v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = threadIdx.x * 33 + j;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        ld32(tb + c * 32, v);
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(v[i & 31]);
    }
  } else if (mode == 1) {
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = v[j];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 6; ++c) st16(tb + c * 16, w);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  } else if (mode == 2) {
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = -1e-3f * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int r = 0; r < 12; ++r) {
#pragma unroll
        for (int j = 0; j < 16; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += x[j];
  } else if (mode == 3) {
    float mx = -1e30f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        ld32(tb + c * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
      }
    }
    acc = mx;
  } else if (mode == 4) {
    float sum = 0.0f;
    const float sc = 0.18f, mxs = 3.0f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        ld32(tb + c * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float e0, e1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(v[j]), sc, -mxs)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(v[j + 1]), sc, -mxs)));
          sum += e0 + e1;
          asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk[j >> 1]) : "f"(e1), "f"(e0));
        }
        st16(tb + 192 + (c & 3) * 16, pk);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    acc = sum;
  } else if (mode == 5) {
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = -1e-3f * (threadIdx.x + j) - 0.3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int r = 0; r < 12; ++r) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          // 2^x, x <= 0: split x = n + f with the round-to-nearest magic constant, cubic on f, exponent add
          float xx = fmaxf(x[j], -126.0f);
          float t = xx + 12582912.0f;             // 1.5 * 2^23: integer part lands in the low mantissa bits
          float n = t - 12582912.0f;
          float f = xx - n;                       // [-0.5, 0.5]
          float p = fmaf(f, 0.0555041086f, 0.2402265069f);
          p = fmaf(p, f, 0.6931471806f);
          p = fmaf(p, f, 1.0f);
          x[j] = __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23)) - 1.0f;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += x[j];
  }
  else if (mode == 6) {  // cvt.rn.f16x2.f32 alone: which pipe / rate?
    float x[16];
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = 1e-3f * (threadIdx.x + j);
    uint32_t accu = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int r = 0; r < 12; ++r) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(o[j >> 1]) : "f"(x[j + 1]), "f"(x[j]));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) accu ^= o[j];
      }
    }
    acc = __uint_as_float(accu);
  } else if (mode == 7 || mode == 8) {
    // pass 2 with the next chunk's tcgen05.ld in flight; 7: cvt.rn.f16x2 pack, 8: exponent-rebias bit trick
    float sum = 0.0f;
    const float sc = 0.18f, mxs = 3.0f + (mode == 8 ? 108.0f : 0.0f);
    uint32_t vb[32];
    for (int i = 0; i < iters; ++i) {
      ld32(tb, v);
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t(&cur)[32] = (c & 1) ? vb : v;
        uint32_t(&nxt)[32] = (c & 1) ? v : vb;
        if (c + 1 < 6) ld32(tb + (c + 1) * 32, nxt);
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float e0, e1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(cur[j]), sc, -mxs)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(cur[j + 1]), sc, -mxs)));
          sum += e0 + e1;
          if (mode == 7) {
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk[j >> 1]) : "f"(e1), "f"(e0));
          } else {
            const uint32_t a = __float_as_uint(e0) >> 13, b = __float_as_uint(e1) << 3;
            pk[j >> 1] = (a & 0xFFFFu) | (b & 0xFFFF0000u);
          }
        }
        st16(tb + 192 + (c & 3) * 16, pk);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    acc = sum;
  } else if (mode == 9) {
    // pass 1 with prefetch and 3-input max
    float mx = -1e30f;
    uint32_t vb[32];
    for (int i = 0; i < iters; ++i) {
      ld32(tb, v);
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t(&cur)[32] = (c & 1) ? vb : v;
        uint32_t(&nxt)[32] = (c & 1) ? v : vb;
        if (c + 1 < 6) ld32(tb + (c + 1) * 32, nxt);
        float m0 = mx, m1 = -1e30f;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(cur[j + 2]), __uint_as_float(cur[j + 3])));
        }
        mx = fmaxf(m0, m1);
      }
    }
    acc = mx;
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512) : "memory");
  }
}

int main() {
  long long* cyc;
  float* sink;
  CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
  CK(cudaMalloc(&sink, 148 * 512 * sizeof(float)));
  const char* names[] = {"tcgen05.ld x32 (6 per wait)", "tcgen05.st x16 (6 per wait)", "ex2.approx",
                         "ld x32 + wait + fmax (pass 1)", "ld + ffma + ex2 + add + cvt + st (pass 2)", "poly exp2 (FMA pipe)",
                         "cvt.rn.f16x2.f32", "pass 2 prefetched, cvt pack", "pass 2 prefetched, bit-trick pack",
                         "pass 1 prefetched, 2 chains"};
  const int iters = 200;
  for (int mode = 0; mode < 10; ++mode) {
    for (int warps : {4, 8, 16}) {
      ubench<<<148, warps * 32, 0>>>(mode, iters, cyc, sink);
      CK(cudaDeviceSynchronize());
      ubench<<<148, warps * 32, 0>>>(mode, iters, cyc, sink);
      CK(cudaDeviceSynchronize());
      long long h[148];
      CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      double per_sm_elems = 0, bytes = 0;
      if (mode == 0 || mode == 3 || mode == 4 || mode >= 7) bytes = double(iters) * 6 * 32 * 32 * 4 * warps, per_sm_elems = bytes / 4;
      if (mode == 1) bytes = double(iters) * 6 * 16 * 32 * 4 * warps, per_sm_elems = bytes / 4;
      if (mode == 2 || mode == 5 || mode == 6) per_sm_elems = double(iters) * 12 * 16 * 32 * warps;
      printf("mode %d %-44s warps %2d: %9lld cycles  %.2f elem/clk/SM  %.1f B/clk/SM\n", mode, names[mode], warps, mx,
             per_sm_elems / mx, bytes / mx);
    }
  }
  return 0;
}
