// tcgen05.mma issue-rate probe for the attention kernel's two MMA shapes on B200 (one CTA, one issuing thread):
// how long does a batch of dependent MMAs take from first issue to the commit's mbarrier arrival?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I proto-clip_b200/csrc -o tools/ubench_mma tools/ubench_mma.cu
#include <stdio.h>

#include "ptx.cuh"

using namespace pc;

struct Case {
  int ts;        // 1: A from TMEM, 0: A from smem (K-major SW128)
  int b_mn;      // 1: B MN-major SW128, 0: B K-major SW128
  int n;         // UMMA N
  int count;     // MMAs per batch
  int rotate_d;  // 1: each MMA writes a different accumulator (no D dependency)
  int k_adv;     // 1: operands advance per MMA as in the real kernels, 0: same operands every time
};

__global__ void __launch_bounds__(128) mma_probe(Case c, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {  // whole warp runs the loop (uniform control flow); one elected lane issues
    const uint32_t a_addr = smem_u32(smem);            // 16 KB: 128 rows x 128 B
    const uint32_t b_addr = smem_u32(smem + 16384);    // up to 64 KB
    const uint32_t idesc = umma_idesc_f16(128, c.n, 0, c.b_mn);
    uint32_t phase = 0;
    long long best = 1ll << 60, total = 0;
    for (int it = 0; it < iters; ++it) {
      const long long t0 = clock64();
      if (elect_one()) {
      for (int k = 0; k < c.count; ++k) {
        const int ka = c.k_adv ? k : 0;
        const uint32_t d = tmem + 256 + (c.rotate_d ? (k & 3) * 64 : 0);
        uint64_t db;
        if (c.b_mn) db = umma_desc_mnmajor_sw128(b_addr + ka * 2048, 1024);
        else db = umma_desc_kmajor_sw128(b_addr + (ka & 3) * 32 + (ka >> 2) * c.n * 128);
        if (c.ts) umma_f16_ts(d, tmem + 8 * ka, db, idesc, k != 0 ? 1u : 0u);
        else umma_f16_ss(d, umma_desc_kmajor_sw128(a_addr + (ka & 3) * 32), db, idesc, k != 0 ? 1u : 0u);
      }
      umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1;
      const long long dt = clock64() - t0;
      if (it > 0) {
        total += dt;
        best = dt < best ? dt : best;
      }
    }
    if (lane == 0) {
      out[0] = best;
      out[1] = total / (iters - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const Case cases[] = {
      {0, 0, 256, 4, 0, 1},  {0, 0, 256, 8, 0, 0},  {0, 0, 208, 4, 0, 1},  {0, 0, 208, 8, 0, 0},  {0, 0, 192, 4, 0, 1},
      {0, 0, 192, 8, 0, 0},  {0, 0, 224, 8, 0, 0},  {0, 0, 128, 4, 0, 1},  {0, 0, 128, 8, 0, 0},  {0, 0, 64, 13, 0, 1},
      {0, 0, 64, 26, 0, 0},  {1, 1, 64, 13, 0, 1},  {1, 1, 64, 26, 0, 0},  {1, 1, 64, 13, 1, 1},  {1, 1, 64, 26, 1, 0},
      {0, 1, 64, 13, 0, 1},  {0, 1, 64, 26, 0, 0},  {1, 0, 64, 13, 0, 1},  {1, 0, 64, 26, 0, 0},  {1, 1, 64, 1, 0, 1},
      {1, 1, 128, 13, 0, 0}, {1, 1, 128, 26, 0, 0}, {1, 1, 256, 13, 0, 0}, {0, 0, 208, 1, 0, 1},  {1, 1, 64, 4, 0, 1},
      {0, 0, 16, 4, 0, 1},   {0, 0, 80, 4, 0, 1},   {0, 0, 80, 8, 0, 0},
  };
  for (const Case& c : cases) {
    mma_probe<<<1, 128, 100 * 1024>>>(c, 50, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    const double ideal = double(c.count) * c.n / 2.0;  // 128 x N x 16 at 8192 flop/clk
    printf("%s A, %s B, N=%3d, %2d MMAs%s%s: best %6lld avg %6lld cycles per batch (%.0f per MMA; ideal %.0f per batch) %s\n",
           c.ts ? "tmem" : "smem", c.b_mn ? "MN-major" : "K-major ", c.n, c.count, c.rotate_d ? ", rotating D" : "",
           c.k_adv ? "" : ", same operands", h[0], h[1], double(h[0]) / c.count, ideal,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
