// TMEM access latency probe (B200): what does one tcgen05.ld / tcgen05.st round trip cost a warp, and which
// waits expose which latencies? One CTA per SM, W warps, each warp on its own lanes / columns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I proto-clip_b200/csrc -o tools/ubench_tmem tools/ubench_tmem.cu
#include <stdio.h>

#include "ptx.cuh"

using namespace pc;

__device__ __forceinline__ float red16(const uint32_t (&v)[16], float m) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) m = fmaxf(m, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
  return m;
}
__device__ __forceinline__ float red32(const uint32_t (&v)[32], float m) {
  float m1 = m;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    m = fmaxf(m, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
    m1 = fmaxf(m1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
  }
  return fmaxf(m, m1);
}
__device__ __forceinline__ float exp16(const uint32_t (&v)[16], uint32_t (&pk)[8], float s) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(v[j]), 0.18f, -3.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(v[j + 1]), 0.18f, -3.0f)));
    s += e0 + e1;
    pk[j >> 1] = pack_half2(e0, e1);
  }
  return s;
}

// mode 0: ld x16, wait                                  (pure load round trip)
// mode 1: ld x32, wait
// mode 2: ld x16, wait, st x8, wait::st                  (load + store round trips, serial)
// mode 3: st x8 then ld x16 (other columns), wait::ld    (does wait::ld also cover the store?)
// mode 4: st x8, wait::st                                (pure store round trip)
// mode 5: pass-1 shape: ld x16 (next) in flight during a 16-column max
// mode 6: pass-2 shape as in the kernel: wait, ld next, exp, st
// mode 7: pass-2 with the store delayed by one chunk (issued after the next wait)
// mode 8: pass-1 shape with x32 chunks
// mode 9: 4 x ld x16 then one wait                        (batched loads)
__global__ void __launch_bounds__(640) probe(int mode, int iters, long long* cycles, float* sink, int spinners, int sleep_ns, int mma_mode) {
  extern __shared__ uint8_t dsm_raw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t never, done;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if ((threadIdx.x & 31) == 0) {
      mbar_init(&never, 1);
      mbar_init(&done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // warp w: lanes 32 * (w % 4), columns 128 * ((w / 4) % 4)
  const uint32_t tb = tmem_base_s + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  float acc = -1e30f;
  uint32_t a[16], b[16], pk[8], pk2[8], w32[32];
#pragma unroll
  for (int j = 0; j < 8; ++j) pk[j] = pk2[j] = threadIdx.x + j;
  // give the columns defined contents
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = __float_as_uint(0.01f * (threadIdx.x + j));
  for (int c = 0; c < 128; c += 16) tmem_st_32x16(tb + c, a);
  tmem_wait_st();
  __syncthreads();
  const int workers = (blockDim.x >> 5) - spinners;
  if (warp >= workers) {
    if (mma_mode && warp == workers) {
      // one warp keeps the tensor core busy like the attention kernel's issuers: S-shaped (SS, N = 208) and
      // PV-shaped (TS, N = 64) batches into columns [256, 512), until the workers are done
      uint8_t* dsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsm_raw) + 1023) & ~uintptr_t(1023));
      const uint32_t a_addr = smem_u32(dsm), b_addr = smem_u32(dsm + 16384);
      const uint64_t da = umma_desc_kmajor_sw128(a_addr), db = umma_desc_kmajor_sw128(b_addr);
      const uint64_t dv = umma_desc_mnmajor_sw128(b_addr, 1024);
      uint32_t ph = 0;
      while (!mbar_try_wait(&done, 0)) {
        if (elect_one()) {
          if (mma_mode & 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base_s + 256, da + 2 * k, db + 2 * k, umma_idesc_f16(128, 208, 0, 0), k ? 1u : 0u);
          }
          if (mma_mode & 2) {
#pragma unroll
            for (int k = 0; k < 13; ++k)
              umma_f16_ts(tmem_base_s + 448, tmem_base_s + 256 + 8 * k, dv + 128 * k, umma_idesc_f16(128, 64, 0, 1), k ? 1u : 0u);
          }
          umma_commit(&never);
        }
        __syncwarp();
        mbar_wait(&never, ph);
        ph ^= 1;
      }
      tc_fence_before();
      __syncthreads();
      if (warp == 0) tmem_dealloc(tmem_base_s, 512);
      return;
    }
    // spinner warps: poll a barrier that completes only when the workers are done (like the kernel's idle roles)
    uint32_t spins = 0;
    while (!mbar_try_wait(&done, 0)) {
      if (sleep_ns) __nanosleep(sleep_ns);
      if (++spins > (1u << 26)) break;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 512);
    return;
  }
  const long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) {
      tmem_ld_32x16(tb + (i & 7) * 16, a);
      tmem_wait_ld();
      acc = fmaxf(acc, __uint_as_float(a[i & 15]));
    }
  } else if (mode == 1) {
    for (int i = 0; i < iters; ++i) {
      tmem_ld_32x32(tb + (i & 3) * 32, w32);
      tmem_wait_ld();
      acc = fmaxf(acc, __uint_as_float(w32[i & 31]));
    }
  } else if (mode == 2) {
    for (int i = 0; i < iters; ++i) {
      tmem_ld_32x16(tb + (i & 3) * 16, a);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = a[j] ^ a[j + 8];
      tmem_st_32x8(tb + 64 + (i & 3) * 8, pk);
      tmem_wait_st();
    }
    acc = __uint_as_float(pk[0]);
  } else if (mode == 3) {
    for (int i = 0; i < iters; ++i) {
      tmem_st_32x8(tb + 64 + (i & 3) * 8, pk);
      tmem_ld_32x16(tb + (i & 3) * 16, a);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = a[j] ^ a[j + 8];
    }
    tmem_wait_st();
    acc = __uint_as_float(pk[0]);
  } else if (mode == 4) {
    for (int i = 0; i < iters; ++i) {
      tmem_st_32x8(tb + 64 + (i & 3) * 8, pk);
      tmem_wait_st();
    }
  } else if (mode == 5) {
    tmem_ld_32x16(tb, a);
    for (int i = 0; i < iters; i += 2) {
      tmem_wait_ld();
      tmem_ld_32x16(tb + ((i + 1) & 7) * 16, b);
      acc = red16(a, acc);
      tmem_wait_ld();
      tmem_ld_32x16(tb + ((i + 2) & 7) * 16, a);
      acc = red16(b, acc);
    }
    tmem_wait_ld();
  } else if (mode == 6) {
    float s = 0.0f;
    tmem_ld_32x16(tb, a);
    for (int i = 0; i < iters; i += 2) {
      tmem_wait_ld();
      tmem_ld_32x16(tb + 32 + ((i + 1) & 3) * 16, b);
      s = exp16(a, pk, s);
      tmem_st_32x8(tb + (i & 3) * 8, pk);
      tmem_wait_ld();
      tmem_ld_32x16(tb + 32 + ((i + 2) & 3) * 16, a);
      s = exp16(b, pk, s);
      tmem_st_32x8(tb + ((i + 1) & 3) * 8, pk);
    }
    tmem_wait_ld();
    tmem_wait_st();
    acc = s;
  } else if (mode == 7) {
    float s = 0.0f;
    tmem_ld_32x16(tb + 32, a);
    for (int i = 0; i < iters; i += 2) {
      tmem_wait_ld();
      tmem_st_32x8(tb + ((i + 3) & 3) * 8, pk2);  // previous chunk's P
      tmem_ld_32x16(tb + 32 + ((i + 1) & 3) * 16, b);
      s = exp16(a, pk, s);
      tmem_wait_ld();
      tmem_st_32x8(tb + (i & 3) * 8, pk);
      tmem_ld_32x16(tb + 32 + ((i + 2) & 3) * 16, a);
      s = exp16(b, pk2, s);
    }
    tmem_wait_ld();
    tmem_wait_st();
    acc = s;
  } else if (mode == 8) {
    uint32_t x32[32];
    tmem_ld_32x32(tb, w32);
    for (int i = 0; i < iters; i += 2) {
      tmem_wait_ld();
      tmem_ld_32x32(tb + ((i + 1) & 3) * 32, x32);
      acc = red32(w32, acc);
      tmem_wait_ld();
      tmem_ld_32x32(tb + ((i + 2) & 3) * 32, w32);
      acc = red32(x32, acc);
    }
    tmem_wait_ld();
  } else if (mode == 9) {
    uint32_t c2[16], d2[16];
    for (int i = 0; i < iters; i += 4) {
      tmem_ld_32x16(tb, a);
      tmem_ld_32x16(tb + 16, b);
      tmem_ld_32x16(tb + 32, c2);
      tmem_ld_32x16(tb + 48, d2);
      tmem_wait_ld();
      acc = red16(a, acc);
      acc = red16(b, acc);
      acc = red16(c2, acc);
      acc = red16(d2, acc);
    }
  }
  const long long t1 = clock64();
  named_bar_sync(1, workers * 32);
  if (threadIdx.x == 0) mbar_arrive(&done);
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 32 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(pk[1]) + __uint_as_float(pk2[2]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

int main() {
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, 148 * 32 * sizeof(long long));
  cudaMalloc(&sink, 148 * 640 * sizeof(float));
  const char* names[] = {"ld x16 + wait",           "ld x32 + wait",          "ld x16, wait, st x8, wait::st", "st x8; ld x16; wait::ld",
                         "st x8 + wait::st",        "pass 1 (x16, prefetch)", "pass 2 (kernel order)",         "pass 2 (store delayed)",
                         "pass 1 (x32, prefetch)",  "4 x ld x16, one wait (per chunk)"};
  const int iters = 400;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int cfg = 0; cfg < 4; ++cfg)
  for (int mode : {0, 5, 6}) {
    const int spinners = cfg == 0 ? 0 : 1, sleep_ns = 0, mma_mode = cfg;
    printf("-- concurrent MMA stream: %s\n", cfg == 0 ? "none" : cfg == 1 ? "S-shaped (SS N=208)" : cfg == 2 ? "PV-shaped (TS N=64)" : "both");
    for (int warps : {4, 8, 16}) {
      probe<<<148, (warps + spinners) * 32, 100 * 1024>>>(mode, iters, cyc, sink, spinners, sleep_ns, mma_mode);
      cudaDeviceSynchronize();
      probe<<<148, (warps + spinners) * 32, 100 * 1024>>>(mode, iters, cyc, sink, spinners, sleep_ns, mma_mode);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[32];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("mode %d %-34s warps %2d: %7.1f cycles per chunk-iteration %s\n", mode, names[mode], warps, double(mx) / iters,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
