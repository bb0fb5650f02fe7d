// Multi-head self-attention core, softmax(Q K^T / sqrt(64) [+ causal mask]) V, for the CLIP towers
// (reference clip/model.py:173,183-185 -> nn.MultiheadAttention -> scaled_dot_product_attention; the text
// tower's additive mask clip/model.py:326-332 is exactly "j > i -> -inf", i.e. the causal flag here).
//
// Sequence lengths are tiny and fixed (50 / 77 / 197 / 257 tokens), so one CTA owns one
// (image, head, 128-query tile) and sees the WHOLE key axis at once: no online-softmax rescaling.
//   warp 4 (one lane)  TMA: Q tile [128,64], K [Lp,64], V [Lp,64] straight out of the packed qkv
//                      activation (128B swizzle); tcgen05.mma S = Q K^T into TMEM (N = Lp columns);
//                      later tcgen05.mma O = P V (A = P from smem K-major, B = V MN-major).
//   warps 0..3         one query row per thread (TMEM lane == row): two passes over the S row in TMEM
//                      (max, then exp2 / sum), P written as fp16 into the swizzled K-major smem tile,
//                      finally O * (1/sum) -> fp16 -> global.
// P aliases the Q and K staging buffers (dead once S is complete), which keeps a CTA at ~90 KB smem /
// 256 TMEM columns for L <= 256 so two CTAs share an SM and overlap each other's phases.
#include "kernels.cuh"
#include "ptx.cuh"

namespace pc {

namespace {

constexpr int ATT_THREADS = 160;
constexpr int HEAD_DIM = 64;

struct AttnParams {
  __half* out;     // [B*L, d]
  int L;           // tokens per sequence
  int Lp16;        // L rounded up to 16 (UMMA N / K extent)
  int kv_box;      // rows per K/V TMA box
  int kv_split;    // number of K/V boxes
  int heads;
  int d;           // heads * 64
  int causal;
  int m_tiles;     // ceil(L / 128)
  int off_v;       // smem offset of V (bytes, 1024-aligned); Q at 0, K at 16384, P at 0
  int off_bars;
  int tmem_cols;   // power of two >= max(Lp16, 64)
};

struct AttnBars {
  uint64_t qk, v, s, p, o;
  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sP = smem;
  uint8_t* sV = smem + p.off_v;
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + p.off_bars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mt = blockIdx.x % p.m_tiles;
  const int bh = blockIdx.x / p.m_tiles;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int m0 = mt * 128;
  const int row0 = b * p.L;  // first token row of this sequence in the [B*L, .] activations

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      mbar_init(&bars->qk, 1);
      mbar_init(&bars->v, 1);
      mbar_init(&bars->s, 1);
      mbar_init(&bars->p, 128);
      mbar_init(&bars->o, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 4) {
    if (lane == 0) {
      const int kv_bytes = p.kv_box * p.kv_split * 128;
      mbar_arrive_expect_tx(&bars->qk, 16384 + kv_bytes);
      tma_load_2d(sQ, &tmQ, &bars->qk, h * HEAD_DIM, row0 + m0);
      for (int s = 0; s < p.kv_split; ++s)
        tma_load_2d(sK + s * p.kv_box * 128, &tmKV, &bars->qk, p.d + h * HEAD_DIM, row0 + s * p.kv_box);
      mbar_arrive_expect_tx(&bars->v, kv_bytes);
      for (int s = 0; s < p.kv_split; ++s)
        tma_load_2d(sV + s * p.kv_box * 128, &tmKV, &bars->v, 2 * p.d + h * HEAD_DIM, row0 + s * p.kv_box);

      // S[128, Lp16] = Q K^T, in column chunks of <= 256 (UMMA N limit)
      mbar_wait(&bars->qk, 0);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t k_addr = smem_u32(sK);
      for (int c0 = 0; c0 < p.Lp16; c0 += 256) {
        const int nc = min(256, p.Lp16 - c0);
        const uint32_t idesc = umma_idesc_f16(128, nc, 0, 0);
#pragma unroll
        for (int k = 0; k < HEAD_DIM / 16; ++k) {
          umma_f16_ss(tmem + c0, umma_desc_kmajor_sw128(q_addr + k * 32),
                      umma_desc_kmajor_sw128(k_addr + c0 * 128 + k * 32), idesc, k != 0 ? 1u : 0u);
        }
      }
      umma_commit(&bars->s);

      // O[128, 64] = P V  (P: K-major 128x64 swizzled chunks; V: MN-major, 16 key rows per K step)
      mbar_wait(&bars->p, 0);
      mbar_wait(&bars->v, 0);
      tc_fence_after();
      const uint32_t p_addr = smem_u32(sP);
      const uint32_t v_addr = smem_u32(sV);
      const uint32_t idesc_o = umma_idesc_f16(128, HEAD_DIM, 0, 1);
      const int k_steps = p.Lp16 / 16;
      for (int kk = 0; kk < k_steps; ++kk) {
        umma_f16_ss(tmem, umma_desc_kmajor_sw128(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32),
                    umma_desc_mnmajor_sw128(v_addr + kk * 2048, 1024), idesc_o, kk != 0 ? 1u : 0u);
      }
      umma_commit(&bars->o);
    }
  } else {
    // ---------------------------------------------------------------- softmax + output (row per thread)
    const int r = threadIdx.x;  // 0..127 == TMEM lane
    const int i = m0 + r;       // query index inside the sequence
    const int jmax = p.causal ? min(i, p.L - 1) : p.L - 1;
    const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)

    mbar_wait(&bars->s, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c0 = 0; c0 < p.Lp16; c0 += 16) {
      uint32_t v[16];
      tmem_ld_32x16(t_row + c0, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c0 + j <= jmax) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float mxs = mx * sc;
    float sum = 0.0f;
    for (int c0 = 0; c0 < p.Lp16; c0 += 16) {
      uint32_t v[16];
      tmem_ld_32x16(t_row + c0, v);
      tmem_wait_ld();
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        float e0 = (c0 + j <= jmax) ? ex2_approx(fmaf(__uint_as_float(v[j]), sc, -mxs)) : 0.0f;
        float e1 = (c0 + j + 1 <= jmax) ? ex2_approx(fmaf(__uint_as_float(v[j + 1]), sc, -mxs)) : 0.0f;
        sum += e0 + e1;
        pk[j >> 1] = pack_half2(e0, e1);
      }
      const int u = c0 >> 3;  // 16-byte unit index along the key axis
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int uu = u + e;
        uint8_t* dst = sP + (uu >> 3) * 16384 + r * 128 + (((uu & 7) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * e], pk[4 * e + 1], pk[4 * e + 2], pk[4 * e + 3]);
      }
    }
    fence_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async proxy
    tc_fence_before();
    mbar_arrive(&bars->p);

    mbar_wait(&bars->o, 0);
    tc_fence_after();
    const float inv = __fdividef(1.0f, sum);
    uint32_t o[4][16];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld_32x16(t_row + c * 16, o[c]);
    tmem_wait_ld();
    if (i < p.L) {
      __half* dst = p.out + static_cast<size_t>(row0 + i) * p.d + h * HEAD_DIM;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          uint4 w;
          w.x = pack_half2(__uint_as_float(o[c][8 * e + 0]) * inv, __uint_as_float(o[c][8 * e + 1]) * inv);
          w.y = pack_half2(__uint_as_float(o[c][8 * e + 2]) * inv, __uint_as_float(o[c][8 * e + 3]) * inv);
          w.z = pack_half2(__uint_as_float(o[c][8 * e + 4]) * inv, __uint_as_float(o[c][8 * e + 5]) * inv);
          w.w = pack_half2(__uint_as_float(o[c][8 * e + 6]) * inv, __uint_as_float(o[c][8 * e + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 16 + e * 8) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

}  // namespace

int launch_attention(const __half* qkv, __half* out, int B, int L, int heads, int causal,
                     cudaStream_t stream) {
  PC_REQUIRE(qkv && out && B > 0 && L > 0 && heads > 0, PC_ERR_ARG, "attention: bad arguments");
  PC_REQUIRE(L <= 512, PC_ERR_ARG,
             "attention: L = %d > 512 needs the online-softmax variant (not built yet)", L);
  const int d = heads * HEAD_DIM;
  AttnParams p;
  p.out = out;
  p.L = L;
  p.Lp16 = (L + 15) / 16 * 16;
  p.kv_split = (p.Lp16 + 255) / 256;
  p.kv_box = ((p.Lp16 + p.kv_split - 1) / p.kv_split + 7) / 8 * 8;
  p.heads = heads;
  p.d = d;
  p.causal = causal;
  p.m_tiles = (L + 127) / 128;
  const int kv_bytes = p.kv_box * p.kv_split * 128;
  const int p_bytes = ((p.Lp16 + 63) / 64) * 16384;
  int region0 = 16384 + kv_bytes;
  if (p_bytes > region0) region0 = p_bytes;
  region0 = (region0 + 1023) / 1024 * 1024;
  p.off_v = region0;
  p.off_bars = region0 + (kv_bytes + 1023) / 1024 * 1024;
  int cols = 64;
  while (cols < p.Lp16) cols *= 2;
  p.tmem_cols = cols;
  const int smem_bytes = p.off_bars + 64 + 1024;
  PC_REQUIRE(smem_bytes <= 227 * 1024, PC_ERR_ARG, "attention: L = %d needs %d B smem", L, smem_bytes);

  static int configured_bytes = 0;
  if (smem_bytes > configured_bytes) {
    PC_CHECK_CUDA(
        cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured_bytes = smem_bytes;
  }
  CUtensorMap tmQ, tmKV;
  const uint64_t rows = static_cast<uint64_t>(B) * L;
  PC_TRY(make_tmap_f16_2d(&tmQ, qkv, 3 * d, rows, static_cast<uint64_t>(3 * d) * 2, 64, 128));
  PC_TRY(make_tmap_f16_2d(&tmKV, qkv, 3 * d, rows, static_cast<uint64_t>(3 * d) * 2, 64, p.kv_box));
  const int grid = B * heads * p.m_tiles;
  attention_kernel<<<grid, ATT_THREADS, smem_bytes, stream>>>(tmQ, tmKV, p);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

}  // namespace pc
