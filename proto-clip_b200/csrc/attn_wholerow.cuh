// Softmax building blocks of the whole-row / block-row attention kernels (attention6.cu, attention7.cu): a tile's score
// columns live in TMEM ([0,128) part a, from column 128 part b); every query row is shared by two threads that each own a
// contiguous range of 16-key chunks, overwrite the score columns they have read with their packed fp16 P, and meet
// through shared memory behind a 64-thread named barrier.
#pragma once
#include "attn_common.cuh"
#include "ptx.cuh"

namespace pc {
namespace wr {

constexpr int SB_COL = 128;  // first TMEM column of part b (keys >= 128 of the block)

__device__ __forceinline__ float lo_f(uint64_t v) { return __uint_as_float(static_cast<uint32_t>(v)); }
__device__ __forceinline__ float hi_f(uint64_t v) { return __uint_as_float(static_cast<uint32_t>(v >> 32)); }
__device__ __forceinline__ uint64_t pack_u32x2(uint32_t lo, uint32_t hi) {
  return static_cast<uint64_t>(lo) | (static_cast<uint64_t>(hi) << 32);
}
__device__ __forceinline__ void pair_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// TMEM column of the fp16 P of score chunk k (16 keys -> 8 columns): in place inside the chunk range of the thread that
// produced it. Thread 0 of a row owns chunks [0, h0); thread 1 owns [h0, nch): its part-a chunks [h0, 8) are packed from
// column 16 h0, its part-b chunks from column 128.
__device__ __forceinline__ int p_col(int k, int h0) {
  return k < h0 ? 8 * k : k < 8 ? 16 * h0 + 8 * (k - h0) : SB_COL + 8 * (k - 8);
}

template <int N>
__device__ __forceinline__ float max_full(const uint32_t (&v)[N], float mx) {
  float m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    mx = fmaxf(mx, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
    m1 = fmaxf(m1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
  }
  return fmaxf(mx, m1);
}

// Maximum of this thread's score chunks [k0, k1) (at most 7). Every array is defined unconditionally inside its own
// scope: a conditionally defined tcgen05.ld destination is live from the kernel entry for ptxas (it cost 80 registers).
template <bool CAUSAL>
__device__ __forceinline__ float row_max(uint32_t t_row, int k0, int k1, int L, int jmax) {
  float mx = -INFINITY;
  const int kf = CAUSAL ? k0 : max(k0, min(k1, L >> 4));  // chunks below kf hold 16 valid keys for every row
  int k = k0;
  if (k + 4 <= kf) {
    uint32_t A[32], B[32];
    tmem_ld_32x32(t_row + 16 * k, A);
    tmem_ld_32x32(t_row + 16 * k + 32, B);
    tmem_wait_ld();
    mx = max_full<32>(A, mx);
    mx = max_full<32>(B, mx);
    k += 4;
  }
  if (k + 3 <= kf) {  // 4 + 3 chunks: the first thread of a ViT-B/16 row in two rounds
    uint32_t A[32], B[16];
    tmem_ld_32x32(t_row + 16 * k, A);
    tmem_ld_32x16(t_row + 16 * k + 32, B);
    tmem_wait_ld();
    mx = max_full<32>(A, mx);
    mx = max_full<16>(B, mx);
    k += 3;
  }
  if (k + 2 <= kf) {
    uint32_t A[32];
    tmem_ld_32x32(t_row + 16 * k, A);
    tmem_wait_ld();
    mx = max_full<32>(A, mx);
    k += 2;
  }
  if (k < kf) {
    uint32_t A[16];
    tmem_ld_32x16(t_row + 16 * k, A);
    tmem_wait_ld();
    mx = max_full<16>(A, mx);
    ++k;
  }
#pragma unroll 1
  for (; k < k1; ++k) {  // masked chunks: the row's last one (or every chunk under the causal mask); also any chunk
    uint32_t A[16];      // the blocks above left over (they take at most 7)
    tmem_ld_32x16(t_row + 16 * k, A);
    tmem_wait_ld();
    mx = chunk_max<false>(A, jmax - 16 * k, mx);
  }
  return mx;
}

// exp2 on the FMA / ALU pipes for two values at once (MUFU: 16 exp2 / clk / SM): Cody-Waite split x = n + f,
// n = round(x), f in [-0.5, 0.5]; 2^f by a cubic minimax polynomial (max relative error 7.5e-5, below the fp16
// rounding of P); 2^n by adding n to the exponent field. x <= 0 here; clamped at -125.
__device__ __forceinline__ uint64_t exp2_poly_f32x2(uint64_t x2) {
  const float MAGIC = 12582912.0f;  // 1.5 * 2^23: x + MAGIC has round(x) in its low mantissa bits
  const uint64_t x = pack_f32x2(fmaxf(lo_f(x2), -125.0f), fmaxf(hi_f(x2), -125.0f));
  const uint64_t xr = add_f32x2(x, pack_f32x2(MAGIC, MAGIC));
  const uint64_t n = add_f32x2(xr, pack_f32x2(-MAGIC, -MAGIC));
  const uint64_t f = fma_f32x2(n, pack_f32x2(-1.0f, -1.0f), x);
  uint64_t q = fma_f32x2(f, pack_f32x2(0.0551716685f, 0.0551716685f), pack_f32x2(0.2426111251f, 0.2426111251f));
  q = fma_f32x2(q, f, pack_f32x2(0.6932609677f, 0.6932609677f));
  q = fma_f32x2(q, f, pack_f32x2(0.9999280572f, 0.9999280572f));
  const uint32_t r0 = static_cast<uint32_t>(q) + (static_cast<uint32_t>(xr) << 23);
  const uint32_t r1 = static_cast<uint32_t>(q >> 32) + (static_cast<uint32_t>(xr >> 32) << 23);
  return pack_u32x2(r0, r1);
}
// pairs of a 16-column chunk that take the polynomial (NP of 8), spread over the chunk
template <int NP>
__device__ __forceinline__ constexpr bool pair_is_poly(int j) {
  return NP >= 8 ? true : NP <= 0 ? false : ((j + 1) * NP) / 8 != (j * NP) / 8;
}

// N (16 or 32) unmasked score columns at s_addr -> N/2 columns of packed fp16 p = exp2(s * sc + nref) at p_addr.
// Per pair of scores: FFMA2, 2 x MUFU.EX2 (or the polynomial), FADD2 (row sum, two chains), F2FP.
template <int NP>
__device__ __forceinline__ void exp_unit32(uint32_t s_addr, uint32_t p_addr, uint64_t sc2, uint64_t nref2, uint64_t& acc_a,
                                           uint64_t& acc_b) {
  uint32_t A[32], pk[16];
  tmem_ld_32x32(s_addr, A);
  tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint64_t t = fma_f32x2(pack_u32x2(A[2 * j], A[2 * j + 1]), sc2, nref2);
    const uint64_t e = pair_is_poly<NP>(j & 7) ? exp2_poly_f32x2(t) : pack_f32x2(ex2_approx(lo_f(t)), ex2_approx(hi_f(t)));
    if (j & 1) acc_b = add_f32x2(acc_b, e);
    else acc_a = add_f32x2(acc_a, e);
    pk[j] = pack_half2(lo_f(e), hi_f(e));
  }
  tmem_st_32x16(p_addr, pk);
}
template <int NP>
__device__ __forceinline__ void exp_unit16(uint32_t s_addr, uint32_t p_addr, uint64_t sc2, uint64_t nref2, uint64_t& acc_a,
                                           uint64_t& acc_b) {
  uint32_t A[16], pk[8];
  tmem_ld_32x16(s_addr, A);
  tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint64_t t = fma_f32x2(pack_u32x2(A[2 * j], A[2 * j + 1]), sc2, nref2);
    const uint64_t e = pair_is_poly<NP>(j) ? exp2_poly_f32x2(t) : pack_f32x2(ex2_approx(lo_f(t)), ex2_approx(hi_f(t)));
    if (j & 1) acc_b = add_f32x2(acc_b, e);
    else acc_a = add_f32x2(acc_a, e);
    pk[j] = pack_half2(lo_f(e), hi_f(e));
  }
  tmem_st_32x8(p_addr, pk);
}

// score chunks [k0, k1) of this thread -> P at columns pcol + 8 (k - k0). (A 16-column loop with the next chunk's
// scores prefetched measured slower than these 32-column units: 33.7 vs 32.1 us at B = 96, L = 197.)
template <bool CAUSAL, int NP>
__device__ __forceinline__ void exp_chunks(uint32_t t_row, int k0, int k1, int pcol, int L, int jmax, uint64_t sc2,
                                           uint64_t nref2, uint64_t& acc_a, uint64_t& acc_b) {
  const int kf = CAUSAL ? k0 : max(k0, min(k1, L >> 4));  // chunks below kf hold 16 valid keys for every row
  int k = k0;
#pragma unroll 1
  for (; k + 2 <= kf; k += 2) exp_unit32<NP>(t_row + 16 * k, t_row + pcol + 8 * (k - k0), sc2, nref2, acc_a, acc_b);
  if (k < kf) {
    exp_unit16<NP>(t_row + 16 * k, t_row + pcol + 8 * (k - k0), sc2, nref2, acc_a, acc_b);
    ++k;
  }
#pragma unroll 1
  for (; k < k1; ++k) {  // masked chunks
    uint32_t A[16], pk[8];
    tmem_ld_32x16(t_row + 16 * k, A);
    tmem_wait_ld();
    acc_a = chunk_exp<false>(A, pk, jmax - 16 * k, sc2, nref2, acc_a);
    tmem_st_32x8(t_row + pcol + 8 * (k - k0), pk);
  }
}

}  // namespace wr
}  // namespace pc
