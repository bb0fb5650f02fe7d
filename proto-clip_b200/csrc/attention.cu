// Multi-head self-attention core, softmax(Q K^T / sqrt(64) [+ causal mask]) V, for the CLIP towers
// (reference clip/model.py:173,183-185 -> nn.MultiheadAttention -> scaled_dot_product_attention; the text
// tower's additive mask clip/model.py:326-332 is exactly "j > i -> -inf", i.e. the causal flag here).
//
// Persistent kernel, one CTA per SM (it owns all 512 TMEM columns), 19 warps:
//   warps 0-15   softmax: two PAIRS of warpgroups (pair = one 256-column TMEM region = one 128-row query tile in
//                flight). Inside a pair the two warpgroups split the KEY axis: thread (row r, half h) owns row r of
//                the tile (TMEM lane r) and the S columns of half h. S = Q K^T is read from TMEM twice (row max,
//                then exp2 / row sum; the two halves exchange max and sum through shared memory and a 64-thread
//                named barrier); P is written back IN PLACE over the thread's own S columns as packed fp16
//                (tcgen05.st) and consumed by the PV MMA straight from TMEM (A operand in TMEM, "ts" form): no
//                shared-memory round trip and no proxy fence for P. Four softmax warps per scheduler keep the MUFU
//                pipe busy across the TMEM-load latencies; the two pairs run half a period apart so one is in its
//                exponentials while the other waits on the tensor core.
//   warp 16 / 17 MMA issuer of pair 0 / 1: S = Q K^T (Q, K from 128B-swizzled smem), then O (+)= P V (V as the
//                MN-major B operand). Whole warp in the loop, one elected lane issues.
//   warp 18      TMA producer: Q tiles (one buffer per pair) and K/V blocks (two slots, shared by the pairs) cut
//                straight out of the packed qkv activation [B*L, 3d].
// Work decomposition: a GROUP is one K/V stream plus the (up to) two 128-row query tiles that use it:
//   L <= 128   ("split")  the two pairs take two different (image, head) items; a K/V slot holds both items' K/V
//   L  > 128              the two pairs take tiles 2t, 2t+1 of the same item and share its K/V
// Keys are processed in blocks of `kb` columns: the whole row in one block when L <= 256 (no rescaling at all:
// 50 / 77 / 197 tokens), 192-key blocks with an online-softmax rescale of O in TMEM otherwise (257 / 577).
// TMEM region of a pair (256 columns): S [0, n), P of half 0 at [0, cs/2), P of half 1 at [cs, cs + (n-cs)/2)
// (cs = the column where half 1 starts), O at [192, 256) (over dead S columns when n > 192).
#include <stdlib.h>

#include "kernels.cuh"
#include "ptx.cuh"

namespace pc {

namespace {

constexpr int HEAD_DIM = 64;
constexpr int MMA_WARP0 = 16;  // warps 0..15: softmax, [pair][key half][lane quarter]
constexpr int TMA_WARP = 18;
constexpr int ATT_THREADS = 19 * 32;
constexpr int O_COL = 192;     // O accumulator columns inside a pair's TMEM region
constexpr int Q_BYTES = 128 * 128;  // one query tile: 128 rows x 64 fp16
constexpr int KB_MULTI = 192;       // keys per block when the row does not fit one TMEM region

struct AttnParams {
  __half* out;     // [B*L, d]
  int L;           // tokens per sequence
  int lp16;        // L rounded up to 16
  int heads;
  int d;           // heads * 64
  int causal;
  int items;       // B * heads
  int m_tiles;     // ceil(L / 128)
  int split;       // 1: L <= 128, the WGs take different items
  int ppi;         // non-split: tile pairs per item = ceil(m_tiles / 2)
  int n_groups;
  int n_kvb;       // key blocks per row
  int kb;          // keys per block (multiple of 16)
  int sub_bytes;   // split: byte offset of WG 1's K (V) inside a slot's K (V) region
  int kreg_bytes;  // bytes of a slot's K region; the V region follows
  int off_kv;      // smem offset of slot 0 (slot 1 follows at + 2 * kreg_bytes)
  int off_stage;   // 16 x 2 KB output staging blocks (one per softmax warp: 32 rows x 64 B, 64B-swizzled)
  int off_xch;     // max (x2, by block parity) / sum exchange between the two key halves: 3 x [pair][half][128] floats
  int off_bars;
  int stagger;       // 1: pair 1 starts half a period after pair 0 (default); PC_ATTN_NO_STAGGER=1 clears it (A/B)
  long long* trace;  // bring-up only (env PC_ATTN_TRACE=1): [group iteration][WG][8] clock64 samples of CTA 0
};

#define ATRACE(slot, cond)                                                                              \
  do {                                                                                                  \
    if (p.trace != nullptr && blockIdx.x == 0 && (cond) && git < 32) p.trace[(git * 2 + w) * 8 + (slot)] = clock64(); \
  } while (0)

struct AttnBars {
  uint64_t kv_full[2];  // per slot: TMA bytes landed
  uint64_t kv_free[2];  // per slot: both WGs' PV MMAs retired (2 arrivals)
  uint64_t q_full[2];   // per WG
  uint64_t q_free[2];   // per WG: last S MMA of the group retired
  uint64_t s_full[2];   // per WG: S block in TMEM
  uint64_t p_full[2];   // per pair: P block in TMEM (8 warp arrivals)
  uint64_t o_full[2];   // per WG: last PV MMA of the group retired
  uint64_t o_free[2];   // per pair: O read out, region reusable (8 warp arrivals)
  uint32_t tmem_base;
};

// job of WG `w` inside group `g`
struct Job {
  bool active;
  int item;  // b * heads + h
  int tile;  // 128-row query tile inside the sequence
};
__device__ __forceinline__ Job job_of(const AttnParams& p, int g, int w) {
  Job j;
  if (p.split) {
    j.item = 2 * g + w;
    j.tile = 0;
    j.active = j.item < p.items;
  } else if (p.ppi == 1) {  // one tile pair per item (128 < L <= 256): no divisions on the per-group path
    j.item = g;
    j.tile = w;
    j.active = w < p.m_tiles;
  } else {
    j.item = g / p.ppi;
    j.tile = 2 * (g % p.ppi) + w;
    j.active = j.tile < p.m_tiles;
  }
  return j;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- softmax chunk helpers: 16 S columns [c0, c0+16) of this thread's row, already in registers ----------
// cmax = last valid column of THIS row (sequence end; causal rows differ). FULL chunks (every column valid for
// every row of the warp) take the branch-free, mask-free path.
template <bool FULL>
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[16], int c0, int cmax, float mx) {
  float m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    float a = __uint_as_float(v[j]), b = __uint_as_float(v[j + 1]);
    float c = __uint_as_float(v[j + 2]), d = __uint_as_float(v[j + 3]);
    if (!FULL) {
      a = (c0 + j <= cmax) ? a : -INFINITY;
      b = (c0 + j + 1 <= cmax) ? b : -INFINITY;
      c = (c0 + j + 2 <= cmax) ? c : -INFINITY;
      d = (c0 + j + 3 <= cmax) ? d : -INFINITY;
    }
    mx = fmaxf(mx, fmaxf(a, b));
    m1 = fmaxf(m1, fmaxf(c, d));
  }
  return fmaxf(mx, m1);
}
// packed fp32 pair math (one issue slot for two elements)
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
// p = exp2(s * sc - mxs) -> packed fp16 pairs; `acc` accumulates the fp32 (unrounded) p as an (even, odd) pair.
template <bool FULL>
__device__ __forceinline__ uint64_t chunk_exp(const uint32_t (&v)[16], uint32_t (&pk)[8], int c0, int cmax, uint64_t sc2,
                                              uint64_t nmxs2, uint64_t acc) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const uint64_t t = fma_f32x2(static_cast<uint64_t>(v[j]) | (static_cast<uint64_t>(v[j + 1]) << 32), sc2, nmxs2);
    float e0 = ex2_approx(__uint_as_float(static_cast<uint32_t>(t)));
    float e1 = ex2_approx(__uint_as_float(static_cast<uint32_t>(t >> 32)));
    if (!FULL) {
      e0 = (c0 + j <= cmax) ? e0 : 0.0f;
      e1 = (c0 + j + 1 <= cmax) ? e1 : 0.0f;
    }
    acc = add_f32x2(acc, pack_f32x2(e0, e1));
    pk[j >> 1] = pack_half2(e0, e1);
  }
  return acc;
}

// MULTI: more than one key block per row (L > 256): compiles the online-softmax rescale in.
// NCH: upper bound on the 16-column chunks of one key half (the softmax loops are fully unrolled over it: a lone
// warp issues ~0.2 instructions per clock through branchy loop code, so straight-line code is what makes the
// passes short). CAUSAL: the text tower's mask (every chunk takes the masked path).
template <bool MULTI, int NCH, bool CAUSAL>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + p.off_bars);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: keeps role-derived values in uniform registers
  const int lane = threadIdx.x & 31;

  if (warp == TMA_WARP) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      tma_prefetch_desc(&tmO);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->kv_full[i], 1);
        mbar_init(&bars->kv_free[i], 2);
        mbar_init(&bars->q_full[i], 1);
        mbar_init(&bars->q_free[i], 1);
        mbar_init(&bars->s_full[i], 1);
        mbar_init(&bars->p_full[i], 8);
        mbar_init(&bars->o_full[i], 1);
        mbar_init(&bars->o_free[i], 8);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  griddep_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  griddep_wait();  // qkv is the previous kernel's output

  if (warp == TMA_WARP) {
    // ---------------------------------------------------------------------------------- TMA producer
    // The whole warp walks the loops (warp-uniform values stay in uniform registers); one elected lane issues.
    uint32_t u = 0;              // K/V units loaded so far (slot = u & 1)
    uint32_t q_cnt0 = 0, q_cnt1 = 0;  // Q tiles loaded per WG
    const uint32_t kv_box_bytes = static_cast<uint32_t>(p.kb) * 128;
    for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
      const Job j0 = job_of(p, g, 0), j1 = job_of(p, g, 1);
      for (int jb = 0; jb < p.n_kvb; ++jb, ++u) {
        const int slot = u & 1;
        uint8_t* kbuf = smem + p.off_kv + slot * 2 * p.kreg_bytes;
        uint8_t* vbuf = kbuf + p.kreg_bytes;
        mbar_wait(&bars->kv_free[slot], ((u >> 1) & 1) ^ 1);
        if (jb == 0) {
          if (j0.active) mbar_wait(&bars->q_free[0], (q_cnt0 & 1) ^ 1);
          if (j1.active) mbar_wait(&bars->q_free[1], (q_cnt1 & 1) ^ 1);
        }
        if (elect_one()) {
          if (p.split) {
            mbar_arrive_expect_tx(&bars->kv_full[slot], 2 * kv_box_bytes * ((j0.active ? 1 : 0) + (j1.active ? 1 : 0)));
            if (j0.active) {
              const int b = j0.item / p.heads, h = j0.item % p.heads;
              const int row = b * p.L + jb * p.kb;
              tma_load_2d(kbuf, &tmKV, &bars->kv_full[slot], p.d + h * HEAD_DIM, row);
              tma_load_2d(vbuf, &tmKV, &bars->kv_full[slot], 2 * p.d + h * HEAD_DIM, row);
            }
            if (j1.active) {
              const int b = j1.item / p.heads, h = j1.item % p.heads;
              const int row = b * p.L + jb * p.kb;
              tma_load_2d(kbuf + p.sub_bytes, &tmKV, &bars->kv_full[slot], p.d + h * HEAD_DIM, row);
              tma_load_2d(vbuf + p.sub_bytes, &tmKV, &bars->kv_full[slot], 2 * p.d + h * HEAD_DIM, row);
            }
          } else {
            const int b = j0.item / p.heads, h = j0.item % p.heads;
            const int row = b * p.L + jb * p.kb;
            mbar_arrive_expect_tx(&bars->kv_full[slot], 2 * kv_box_bytes);
            tma_load_2d(kbuf, &tmKV, &bars->kv_full[slot], p.d + h * HEAD_DIM, row);
            tma_load_2d(vbuf, &tmKV, &bars->kv_full[slot], 2 * p.d + h * HEAD_DIM, row);
          }
          if (jb == 0) {
            if (j0.active) {
              const int b = j0.item / p.heads, h = j0.item % p.heads;
              mbar_arrive_expect_tx(&bars->q_full[0], Q_BYTES);
              tma_load_2d(smem, &tmQ, &bars->q_full[0], h * HEAD_DIM, b * p.L + j0.tile * 128);
            }
            if (j1.active) {
              const int b = j1.item / p.heads, h = j1.item % p.heads;
              mbar_arrive_expect_tx(&bars->q_full[1], Q_BYTES);
              tma_load_2d(smem + Q_BYTES, &tmQ, &bars->q_full[1], h * HEAD_DIM, b * p.L + j1.tile * 128);
            }
          }
        }
        __syncwarp();
        if (jb == 0) {
          q_cnt0 += j0.active ? 1 : 0;
          q_cnt1 += j1.active ? 1 : 0;
        }
      }
    }
  } else if (warp == MMA_WARP0 || warp == MMA_WARP0 + 1) {
    // ---------------------------------------------------------------------------------- MMA issuer of WG w
    // (whole warp in the loops, one elected lane issues: see the producer)
    const int w = warp - MMA_WARP0;
    const uint32_t region = tmem + w * 256;
    const uint32_t q_addr = smem_u32(smem + w * Q_BYTES);
    const uint64_t q_desc = umma_desc_kmajor_sw128(q_addr);
    const uint32_t idesc_o = umma_idesc_f16(128, HEAD_DIM, 0, 1);
    uint32_t u = 0, q_cnt = 0, st_cnt = 0;
    for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
      const Job j = job_of(p, g, w);
      for (int jb = 0; jb < p.n_kvb; ++jb, ++u) {
        const int slot = u & 1;
        mbar_wait(&bars->kv_full[slot], (u >> 1) & 1);
        if (!j.active) {  // this WG sits the unit out: release its share of the slot
          if (lane == 0) mbar_arrive(&bars->kv_free[slot]);
          __syncwarp();
          continue;
        }
        if (jb == 0) {
          mbar_wait(&bars->q_full[w], q_cnt & 1);
          mbar_wait(&bars->o_free[w], (q_cnt & 1) ^ 1);
        }
        // Stagger the warpgroups by half a period: WG 1 starts its first tile when WG 0 has finished its first
        // softmax, so from then on one WG computes exponentials while the other waits on the tensor core.
        if (p.stagger && w == 1 && u == 0) mbar_wait(&bars->p_full[0], 0);
        tc_fence_after();
        const uint32_t k_addr =
            smem_u32(smem + p.off_kv + slot * 2 * p.kreg_bytes + (p.split ? w * p.sub_bytes : 0));
        const uint32_t v_addr = k_addr + p.kreg_bytes;
        const int n_cols = min(p.kb, p.lp16 - jb * p.kb);
        if (elect_one()) {
          // S[128, n_cols] = Q K^T   (+2 in the descriptor's address field = 32 B = 16 fp16 along K)
          const uint32_t idesc_s = umma_idesc_f16(128, n_cols, 0, 0);
          const uint64_t k_desc = umma_desc_kmajor_sw128(k_addr);
#pragma unroll
          for (int k = 0; k < HEAD_DIM / 16; ++k) umma_f16_ss(region, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&bars->s_full[w]);
          if (jb == p.n_kvb - 1) umma_commit(&bars->q_free[w]);
        }
        __syncwarp();
        // O[128, 64] (+)= P V : P from TMEM (8 columns per 16 keys), V MN-major (16 key rows = 2048 B per K step)
        mbar_wait(&bars->p_full[w], st_cnt & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t v_desc = umma_desc_mnmajor_sw128(v_addr, 1024);
          const int k_steps = n_cols >> 4;
          const int k_half = (k_steps + 1) >> 1;  // k-steps whose P was written by key half 0 (at column 8 * kk)
          umma_f16_ts(region + O_COL, region, v_desc, idesc_o, jb != 0 ? 1u : 0u);
          for (int kk = 1; kk < k_steps; ++kk) {
            const uint32_t a_col = kk < k_half ? 8 * kk : 16 * k_half + 8 * (kk - k_half);
            umma_f16_ts(region + O_COL, region + a_col, v_desc + 128 * kk, idesc_o, 1u);
          }
          umma_commit(&bars->kv_free[slot]);
          if (jb == p.n_kvb - 1) umma_commit(&bars->o_full[w]);
        }
        __syncwarp();
        ++st_cnt;
      }
      if (j.active) ++q_cnt;
    }
  } else if (warp < MMA_WARP0) {
    // ---------------------------------------------------------------------------------- softmax, pair w / half hf
    const int w = warp >> 3;
    const int hf = (warp >> 2) & 1;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row of the tile == TMEM lane
    const uint32_t t_row = tmem + w * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t pair_bar = 1 + w * 4 + quarter;  // named barrier of the two warps that share these 32 rows
    float* xmax0 = reinterpret_cast<float*>(smem + p.off_xch);  // [2 (block parity)][pair][half][128]
    float* xsum = xmax0 + 1024;
    const int x_mine = (w * 2 + hf) * 128 + r, x_other = (w * 2 + (hf ^ 1)) * 128 + r;
    const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    uint32_t st_cnt = 0, o_cnt = 0;
    int git = -1;
    for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
      ++git;
      const Job j = job_of(p, g, w);
      if (!j.active) continue;
      const int i = j.tile * 128 + r;  // query index inside the sequence
      const bool warp_live = j.tile * 128 + quarter * 32 < p.L;  // some row of this warp is a real query
      const int jmax = CAUSAL ? min(i, p.L - 1) : p.L - 1;
      float m_run = -INFINITY, sum = 0.0f;
      for (int jb = 0; jb < p.n_kvb; ++jb, ++st_cnt) {
        const int n_cols = min(p.kb, p.lp16 - jb * p.kb);
        const int ncv = min(n_cols, p.L - jb * p.kb);  // valid key columns of this block
        const int cmax = jmax - jb * p.kb;             // last valid column of this row (may be < 0)
        const int cs = (((n_cols >> 4) + 1) >> 1) << 4;  // first column of key half 1
        const int c_lo = hf ? cs : 0;                    // this thread's S columns: [c_lo, c_lo + 16 * n16)
        const int n16 = ((hf ? n_cols : cs) - c_lo) >> 4;
        ATRACE(0, hf == 0 && quarter == 0 && lane == 0 && jb == 0);
        mbar_wait(&bars->s_full[w], st_cnt & 1);
        tc_fence_after();
        ATRACE(1, hf == 0 && quarter == 0 && lane == 0 && jb == 0);
        if (warp_live) {
          const uint32_t t_s = t_row + c_lo;
          // number of leading chunks of this half that need no masking (warp-uniform; 0 for causal rows)
          const int n_full = CAUSAL ? 0 : min(n16, max(0, (ncv - c_lo) >> 4));
          uint32_t R[2][16];
          // ---- pass 1: row maximum over this half (next chunk's load in flight during the reduction)
          float mx = -INFINITY;
          if (n16 > 0) tmem_ld_32x16(t_s, R[0]);
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            if (k < n16) {
              tmem_wait_ld();
              if (k + 1 < NCH && k + 1 < n16) tmem_ld_32x16(t_s + (k + 1) * 16, R[(k + 1) & 1]);
              if (k < n_full) mx = chunk_max<true>(R[k & 1], c_lo + k * 16, cmax, mx);
              else mx = chunk_max<false>(R[k & 1], c_lo + k * 16, cmax, mx);
            }
          }
          float* xmax = xmax0 + (st_cnt & 1) * 512;  // double-buffered: the partner may still read the previous block's
          xmax[x_mine] = mx;
          named_bar_sync(pair_bar, 64);
          mx = fmaxf(mx, xmax[x_other]);
          ATRACE(2, hf == 0 && quarter == 0 && lane == 0 && jb == 0);
          const float m_new = fmaxf(m_run, mx);
          if (MULTI && jb > 0) {
            // online softmax: bring this thread's 32 O columns and its partial sum to the new maximum. s_full(jb)
            // implies that the PV MMA of block jb-1 has retired (same issuing thread, in-order pipe): O is stable.
            const float alpha = ex2_approx((m_run - m_new) * sc);
            if (__any_sync(0xffffffffu, alpha != 1.0f)) {
              const uint32_t t_o = t_row + O_COL + 32 * hf;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                tmem_ld_32x16(t_o + 16 * hh, R[hh]);
              }
              tmem_wait_ld();
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                for (int e = 0; e < 16; ++e) R[hh][e] = __float_as_uint(__uint_as_float(R[hh][e]) * alpha);
                tmem_st_32x16(t_o + 16 * hh, R[hh]);
              }
            }
            sum *= alpha;
          }
          m_run = m_new;
          // ---- pass 2: p = exp2((s - m) / 8 * log2 e), fp16 P written over this thread's own S columns
          {
            const uint64_t sc2 = pack_f32x2(sc, sc), nmxs2 = pack_f32x2(-m_new * sc, -m_new * sc);
            uint64_t acc2 = 0;  // (0.0f, 0.0f)
            uint32_t pk[8];
            if (n16 > 0) tmem_ld_32x16(t_s, R[0]);
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
              if (k < n16) {
                tmem_wait_ld();
                if (k + 1 < NCH && k + 1 < n16) tmem_ld_32x16(t_s + (k + 1) * 16, R[(k + 1) & 1]);
                if (k < n_full) acc2 = chunk_exp<true>(R[k & 1], pk, c_lo + k * 16, cmax, sc2, nmxs2, acc2);
                else acc2 = chunk_exp<false>(R[k & 1], pk, c_lo + k * 16, cmax, sc2, nmxs2, acc2);
                tmem_st_32x8(t_s + k * 8, pk);
              }
            }
            sum += __uint_as_float(static_cast<uint32_t>(acc2)) + __uint_as_float(static_cast<uint32_t>(acc2 >> 32));
          }
          tmem_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_full[w]);
        ATRACE(3, hf == 0 && quarter == 0 && lane == 0 && jb == 0);
      }
      // ---- O / sum -> fp16 -> out[b, i, h*64 + 32*hf .. +31]
      if (warp_live) xsum[x_mine] = sum;  // published before the barrier below
      mbar_wait(&bars->o_full[w], o_cnt & 1);
      ++o_cnt;
      tc_fence_after();
      ATRACE(4, hf == 0 && quarter == 0 && lane == 0);
      uint32_t oa[32];
      if (warp_live) {
        tmem_ld_32x32(t_row + O_COL + 32 * hf, oa);
        tmem_wait_ld();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->o_free[w]);
      if (warp_live) {
        // fp16 half-rows (32 columns = 64 bytes) into this warp's own staging block, 64B-swizzled (16-byte chunk c
        // of row r at chunk c ^ ((r >> 1) & 3): conflict-free), then one TMA store of the 32 x 32 block through the
        // [B][L][d] map -- rows past the sequence end are clipped. No cross-warp hand-off: each warp owns its
        // staging block and its bulk groups.
        uint8_t* stg = smem + p.off_stage + warp * 2048;
        if (elect_one()) tma_store_wait_read<0>();  // this warp's previous store has drained the block
        named_bar_sync(pair_bar, 64);               // partner's partial row sum visible (and the wait above done)
        const float inv = __fdividef(1.0f, sum + xsum[x_other]);
        uint8_t* my_row = stg + lane * 64;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint4 x;
          x.x = pack_half2(__uint_as_float(oa[8 * e + 0]) * inv, __uint_as_float(oa[8 * e + 1]) * inv);
          x.y = pack_half2(__uint_as_float(oa[8 * e + 2]) * inv, __uint_as_float(oa[8 * e + 3]) * inv);
          x.z = pack_half2(__uint_as_float(oa[8 * e + 4]) * inv, __uint_as_float(oa[8 * e + 5]) * inv);
          x.w = pack_half2(__uint_as_float(oa[8 * e + 6]) * inv, __uint_as_float(oa[8 * e + 7]) * inv);
          *reinterpret_cast<uint4*>(my_row + ((e ^ ((lane >> 1) & 3)) << 4)) = x;
        }
        fence_async_smem();
        __syncwarp();
        if (elect_one()) {
          const int b = j.item / p.heads, h = j.item % p.heads;
          tma_store_3d(&tmO, stg, h * HEAD_DIM + 32 * hf, j.tile * 128 + quarter * 32, b);
          tma_store_commit();
        }
      }
      ATRACE(5, hf == 0 && quarter == 0 && lane == 0);
    }
    if (elect_one()) tma_store_wait_all<0>();  // output written before the CTA (and its staging smem) goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <bool MULTI, int NCH, bool CAUSAL>
int launch_variant(int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmKV,
                   const CUtensorMap& tmO, const AttnParams& p) {
  static int configured_bytes = 0;
  auto kern = attention_kernel<MULTI, NCH, CAUSAL>;
  if (smem_bytes > configured_bytes) {
    PC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured_bytes = smem_bytes;
  }
  PC_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(ATT_THREADS), smem_bytes, stream, 1, tmQ, tmKV, tmO, p));
  return PC_OK;
}

}  // namespace

int launch_attention(const __half* qkv, __half* out, int B, int L, int heads, int causal,
                     cudaStream_t stream) {
  PC_REQUIRE(qkv && out && B > 0 && L > 0 && heads > 0, PC_ERR_ARG, "attention: bad arguments");
  PC_REQUIRE(L <= 4096, PC_ERR_ARG, "attention: L = %d is beyond the supported sequence length (4096)", L);
  const int d = heads * HEAD_DIM;
  AttnParams p{};
  p.out = out;
  p.L = L;
  p.lp16 = (L + 15) / 16 * 16;
  p.heads = heads;
  p.d = d;
  p.causal = causal ? 1 : 0;
  p.items = B * heads;
  p.m_tiles = (L + 127) / 128;
  p.split = p.m_tiles == 1 ? 1 : 0;
  p.ppi = (p.m_tiles + 1) / 2;
  p.n_groups = p.split ? (p.items + 1) / 2 : p.items * p.ppi;
  if (p.lp16 <= 256) {
    p.n_kvb = 1;
    p.kb = p.lp16;
  } else {
    p.kb = KB_MULTI;
    p.n_kvb = (p.lp16 + p.kb - 1) / p.kb;
  }
  p.sub_bytes = p.kb * 128;
  p.kreg_bytes = (p.split ? 2 : 1) * p.kb * 128;  // kb % 8 == 0 -> 1024-byte multiples (swizzle atoms)
  p.off_kv = 2 * Q_BYTES;
  p.off_stage = p.off_kv + 4 * p.kreg_bytes;
  p.off_xch = p.off_stage + 8 * 4096;
  p.off_bars = p.off_xch + 3 * 512 * 4;
  int smem_bytes = p.off_bars + static_cast<int>(sizeof(AttnBars)) + 1024;
  // one CTA per SM by construction (each CTA owns all 512 TMEM columns): ask for more than half the SM's smem
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;
  PC_REQUIRE(smem_bytes <= 227 * 1024, PC_ERR_ARG, "attention: L = %d needs %d B smem", L, smem_bytes);

  CUtensorMap tmQ, tmKV, tmO;
  const uint64_t rows = static_cast<uint64_t>(B) * L;
  PC_TRY(make_tmap_f16_2d(&tmQ, qkv, 3 * d, rows, static_cast<uint64_t>(3 * d) * 2, 64, 128));
  PC_TRY(make_tmap_f16_2d(&tmKV, qkv, 3 * d, rows, static_cast<uint64_t>(3 * d) * 2, 64, p.kb));
  PC_TRY(make_tmap_f16_3d(&tmO, out, d, L, B, static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(L) * d * 2, 32, 32));
  const int sms = device_sm_count();
  const int grid = p.n_groups < sms ? p.n_groups : sms;
  static int tracing = -1;
  static long long* trace = nullptr;
  static int stagger = 1;
  if (tracing < 0) {
    const char* e = getenv("PC_ATTN_TRACE");
    tracing = (e && e[0] == '1') ? 1 : 0;
    const char* f = getenv("PC_ATTN_NO_STAGGER");
    stagger = (f && f[0] == '1') ? 0 : 1;
  }
  p.stagger = stagger;
  if (tracing) {
    if (!trace) PC_CHECK_CUDA(cudaMalloc(&trace, 32 * 2 * 8 * sizeof(long long)));
    PC_CHECK_CUDA(cudaMemsetAsync(trace, 0, 32 * 2 * 8 * sizeof(long long), stream));
    p.trace = trace;
  }
  // instantiation: smallest unrolled chunk count that covers one key half of a block
  const int need = ((p.kb >> 4) + 1) >> 1;
  int rc;
  if (p.n_kvb > 1) {
    rc = p.causal ? launch_variant<true, 6, true>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p)
                  : launch_variant<true, 6, false>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p);
  } else if (need <= 4) {
    rc = p.causal ? launch_variant<false, 4, true>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p)
                  : launch_variant<false, 4, false>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p);
  } else if (need <= 7) {
    rc = p.causal ? launch_variant<false, 7, true>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p)
                  : launch_variant<false, 7, false>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p);
  } else {
    rc = p.causal ? launch_variant<false, 8, true>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p)
                  : launch_variant<false, 8, false>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p);
  }
  PC_TRY(rc);
  if (tracing) {
    static int printed = 0;
    long long h[32 * 2 * 8];
    PC_CHECK_CUDA(cudaStreamSynchronize(stream));
    PC_CHECK_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    if (printed++ == 3) {
      const long long t0 = h[0] ? h[0] : h[8];
      fprintf(stderr, "[attn trace] B*heads=%d L=%d (cycles since first sample, CTA 0)\n", p.items, L);
      fprintf(stderr, "group WG  wait_s   s_full  pass1_end  p_arrive   o_full  stored\n");
      for (int g = 0; g < 32; ++g)
        for (int w = 0; w < 2; ++w) {
          const long long* r = h + (g * 2 + w) * 8;
          if (!r[1]) continue;
          fprintf(stderr, "%4d  %d %8lld %8lld %9lld %9lld %8lld %8lld\n", g, w, r[0] - t0, r[1] - t0, r[2] - t0, r[3] - t0,
                  r[4] - t0, r[5] - t0);
        }
    }
  }
  return PC_OK;
}

}  // namespace pc
