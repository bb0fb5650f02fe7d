// Multi-head self-attention core for long unmasked sequences (L > 257: ViT-L/14@336px has 577 tokens), round 2.
//
// attention6.cu keeps the WHOLE score row of a query tile in tensor memory, which ends at 256 keys. This kernel keeps
// that kernel's softmax machinery (attn_wholerow.cuh: two threads per query row, in-place packed fp16 P, part b first,
// lean exponential units) and streams the keys through it in BLOCKS OF 192: per 128-query tile the 256 TMEM columns hold
// S_a = Q K[0:128]^T at [0,128), S_b = Q K[128:192]^T at [128,192) and the O accumulator at [192,256) -- nothing
// aliases, O stays resident over the key blocks and the P V MMAs accumulate onto it. The softmax is the online one with
// a LAZY rescale: a row's reference maximum moves only when a block's maximum exceeds it by more than 2^8 (P stays far
// inside fp16, the row sum is fp32), and only then are the row's 64 O columns read, scaled and written back; for real
// score distributions that happens in the first block or two of a row and never again.
//
// 577 = 3 x 192 + 1: a fourth key block for one key would cost a whole S -> softmax -> P V hop, so the last key is folded
// in by the softmax threads themselves (the extra-key trick of attention6's L = 257 mode: k_x, v_x rows dropped into
// shared memory by the producer, q . k_x as a 64-long dot product split over the row's two threads, p_x v_x added to O
// in the epilogue). Any other remainder is a shorter, masked last block.
//
// What bounds it (measured at 64 images x 16 heads x 577, 204 us = 427 TFLOP/s against 273 us for the round-1 kernel):
// the P V instruction. M = 128, N = 64 (the head dimension), K = 16 with A from tensor memory costs ~ 140 cycles in the
// kernel, 12 of them per 192-key block and tile; with the four N = 192 score instructions a block costs 2.4 k tensor
// cycles, i.e. 126 us for the whole launch with the softmax switched off. Two variants that rearranged the chain did
// not beat this one and were not kept (git history): 96-key blocks with a double-buffered S slot per warpgroup (the
// softmax never waits for an MMA round trip, but the N = 96 score instructions sit on the 110-cycle shared-memory-operand
// floor: 220 us), and one tile per CTA with four threads per row, 192-key double-buffered S and two O buffers (all 16
// softmax warps in the same phase at the same time: 24.5 k exponentials per block are 1.5 k MUFU cycles that nothing
// else overlaps: 234 us).
//
// One CTA per SM, persistent over PASSES = (image, head, pair of query tiles); the two warpgroups take the two tiles and
// share the K / V stream (a ring of three 48 KB stages), alternating on the MUFU pipe through named barriers.
//   warps 0..15   softmax (WG w = warp >> 3, thread hf = (warp >> 2) & 1 of a row, TMEM lane quarter = warp & 3)
//   warps 16, 17  MMA issuer of WG 0 / 1: per block S_j (after P V_{j-1}: the tensor pipe runs in order, so S_j may
//                 overwrite P_{j-1}'s columns), then P V_j part b / part a as the softmax hands them over
//   warp 18       TMA producer (Q tiles of the pass, K / V blocks, the extra rows)
//   warp 19       idle
#include <stdlib.h>

#include "attn_common.cuh"
#include "attn_wholerow.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace pc {
namespace {

using namespace wr;

constexpr int HEAD_DIM = 64;
constexpr int MMA_WARP0 = 16;
constexpr int TMA_WARP = 18;
constexpr int THREADS7 = 20 * 32;
constexpr int KB = 192;                       // keys per block
constexpr int NST = 3;                        // K / V ring stages
constexpr int Q_BYTES = 128 * 128;            // one query tile, 128B-swizzled
constexpr int KV_BYTES = KB * 128;            // K or V of one block: 24 KB (a multiple of the 1024-byte swizzle atom)
constexpr int OFF_Q = 0;                      // Q tile of WG 0, then of WG 1
constexpr int OFF_KV = 2 * Q_BYTES;           // ring: stage s = K block | V block
constexpr int STAGE_BYTES = 2 * KV_BYTES;
constexpr int OFF_OUT = OFF_KV + NST * STAGE_BYTES;  // 16 x 2 KB output staging blocks
constexpr int OFF_XCH = OFF_OUT + 16 * 2048;         // row maximum / dot product / partial sum exchange
constexpr int OFF_XROW = OFF_XCH + 2 * 2 * 128 * 4;  // extra-key mode: k_x | v_x (128 B each, linear)
constexpr int OFF_BARS = OFF_XROW + 256;
constexpr int O_COL = 192;
constexpr float LAZY_LOG2 = 8.0f;  // a row's reference maximum moves only for blocks that exceed it by 2^8 in exp2 units

struct Params7 {
  int L;         // rows of a sequence in memory
  int Lk;        // keys streamed through the tensor cores (L, or L - 1 in extra-key mode)
  int xkey;      // 1: key L - 1 is folded in by the softmax threads
  int heads, d, items;
  int n_blk;     // key blocks per row = ceil(Lk / 192)
  int n_tiles;   // query tiles per sequence = ceil(L / 128)
  int ppi;       // passes per item = ceil(n_tiles / 2)
  int n_pass;    // items * ppi
  int pingpong;
  const __half* qkv;
  long long row_pitch, plane_pitch;  // elements
};

struct Bars7 {
  uint64_t q_full, q_free;             // Q tiles (+ extra rows) of the pass
  uint64_t k_full[NST], v_full[NST];   // per ring stage
  uint64_t k_free[NST], v_free[NST];   // both WGs' S / P V MMAs of the block retired (2 arrivals)
  uint64_t s_full[2];                  // per WG: S of the block in TMEM
  uint64_t pa_full[2], pb_full[2];     // per WG: P part a (8 warps) / part b (4 warps) in TMEM
  uint64_t pv_done[2];                 // per WG: last P V MMA of the pass retired
  uint64_t o_free[2];                  // per WG: O read out by the epilogue (8 warps)
  uint32_t tmem_base;
};

// chunks (16 keys) of block j and the split point between the row's two threads
__device__ __forceinline__ int block_chunks(const Params7& p, int j) {
  const int keys = min(KB, p.Lk - j * KB);
  return (keys + 15) >> 4;
}

template <bool XKEY>
__global__ void __launch_bounds__(THREADS7, 1)
attention7_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                  const __grid_constant__ CUtensorMap tmO, const Params7 p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars7* bars = reinterpret_cast<Bars7*>(smem + OFF_BARS);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == TMA_WARP) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      tma_prefetch_desc(&tmO);
      mbar_init(&bars->q_full, 1);
      mbar_init(&bars->q_free, XKEY ? 2 + 16 : 2);  // XKEY: the softmax warps read their q rows from the tile
      for (int i = 0; i < NST; ++i) {
        mbar_init(&bars->k_full[i], 1);
        mbar_init(&bars->v_full[i], 1);
        mbar_init(&bars->k_free[i], 2);
        mbar_init(&bars->v_free[i], 2);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->s_full[i], 1);
        mbar_init(&bars->pa_full[i], 8);
        mbar_init(&bars->pb_full[i], 4);
        mbar_init(&bars->pv_done[i], 1);
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
  griddep_wait();

  const int g_stride = gridDim.x;
  if (warp == TMA_WARP) {
    // ---------------------------------------------------------------------------------- producer
    uint32_t pass_no = 0, blk = 0;
    for (int g = blockIdx.x; g < p.n_pass; g += g_stride, ++pass_no) {
      const int item = g / p.ppi, tp = g - item * p.ppi;
      const int hd = item % p.heads, r0 = (item / p.heads) * p.L;
      mbar_wait(&bars->q_free, (pass_no & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->q_full, 2 * Q_BYTES + (XKEY ? 256 : 0));
        // rows past the sequence (last tile) belong to the next image or are zero-filled past the tensor: finite
        // either way, and the rows they produce are clipped by the output map
        tma_load_3d(smem + OFF_Q, &tmQ, &bars->q_full, 0, r0 + 256 * tp, hd);
        tma_load_3d(smem + OFF_Q + Q_BYTES, &tmQ, &bars->q_full, 0, r0 + 256 * tp + 128, hd);
        if (XKEY) {
          const __half* row = p.qkv + static_cast<long long>(r0 + p.Lk) * p.row_pitch;
          bulk_load(smem + OFF_XROW, row + (p.heads + hd) * p.plane_pitch, 128, &bars->q_full);
          bulk_load(smem + OFF_XROW + 128, row + (2 * p.heads + hd) * p.plane_pitch, 128, &bars->q_full);
        }
      }
      __syncwarp();
      for (int j = 0; j < p.n_blk; ++j, ++blk) {
        const int st = blk % NST;
        const uint32_t free_par = ((blk / NST) & 1) ^ 1;
        uint8_t* stage = smem + OFF_KV + st * STAGE_BYTES;
        mbar_wait(&bars->k_free[st], free_par);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->k_full[st], KV_BYTES);
          tma_load_3d(stage, &tmKV, &bars->k_full[st], 0, r0 + KB * j, p.heads + hd);
        }
        __syncwarp();
        mbar_wait(&bars->v_free[st], free_par);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->v_full[st], KV_BYTES);
          tma_load_3d(stage + KV_BYTES, &tmKV, &bars->v_full[st], 0, r0 + KB * j, 2 * p.heads + hd);
        }
        __syncwarp();
      }
    }
  } else if (warp == MMA_WARP0 || warp == MMA_WARP0 + 1) {
    // ---------------------------------------------------------------------------------- MMA issuer of WG w
    const int w = warp - MMA_WARP0;
    const uint32_t region = tmem + w * 256;
    const uint32_t idesc_o = umma_idesc_f16(128, HEAD_DIM, 0, 1);
    const uint32_t idesc_s = umma_idesc_f16(128, KB, 0, 0);  // one 192-key instruction per k-step: S_a | S_b are contiguous
    // blk: ring position (every pass); bcount / nb_seen / tcount: this WG's own blocks, blocks with a part b, passes
    uint32_t pass_no = 0, blk = 0, bcount = 0, nb_seen = 0, tcount = 0;
    for (int g = blockIdx.x; g < p.n_pass; g += g_stride, ++pass_no) {
      const int tp = g % p.ppi;
      if (2 * tp + w >= p.n_tiles) {  // odd tile count: no tile for this WG in the pass; release its share of the buffers
        mbar_wait(&bars->q_full, pass_no & 1);
        for (int j = 0; j < p.n_blk; ++j, ++blk) {
          const int st = blk % NST;
          const uint32_t full_par = (blk / NST) & 1;
          mbar_wait(&bars->k_full[st], full_par);
          mbar_wait(&bars->v_full[st], full_par);
          if (elect_one()) {
            mbar_arrive(&bars->k_free[st]);
            mbar_arrive(&bars->v_free[st]);
          }
          __syncwarp();
        }
        if (elect_one()) mbar_arrive(&bars->q_free);
        __syncwarp();
        continue;
      }
      const uint64_t q_desc = umma_desc_kmajor_sw128(smem_u32(smem + OFF_Q + w * Q_BYTES));
      mbar_wait(&bars->q_full, pass_no & 1);
      for (int j = 0; j < p.n_blk; ++j, ++blk, ++bcount) {
        const int st = blk % NST;
        const uint32_t full_par = (blk / NST) & 1;
        const int nch = block_chunks(p, j), h0 = (nch + 1) >> 1;
        const uint32_t kaddr = smem_u32(smem + OFF_KV + st * STAGE_BYTES);
        const uint64_t k_desc = umma_desc_kmajor_sw128(kaddr);
        const uint64_t v_desc = umma_desc_mnmajor_sw128(kaddr + KV_BYTES, 1024);
        mbar_wait(&bars->k_full[st], full_par);
        tc_fence_after();
        // S_j overwrites the columns P_{j-1} sat in: P V_{j-1} was issued before it and the tensor pipe runs in order.
        // Whole 128 / 64 key shapes also for a short last block: the rows past the keys are the next image's (or
        // zero-filled), their scores are masked in the softmax.
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HEAD_DIM / 16; ++k) umma_f16_ss(region, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&bars->s_full[w]);
          umma_commit(&bars->k_free[st]);
          if (j == p.n_blk - 1) umma_commit(&bars->q_free);
        }
        __syncwarp();
        mbar_wait(&bars->v_full[st], full_par);
        // the first P V of the pass overwrites O: the previous pass's epilogue must have read it
        if (j == 0) mbar_wait(&bars->o_free[w], (tcount & 1) ^ 1);
        if (nch > 8) {
          mbar_wait(&bars->pb_full[w], nb_seen & 1);
          ++nb_seen;
          tc_fence_after();
          if (elect_one()) {
            for (int k = 8; k < nch; ++k)
              umma_f16_ts(region + O_COL, region + p_col(k, h0), v_desc + 128 * k, idesc_o, (j > 0 || k != 8) ? 1u : 0u);
          }
          __syncwarp();
        }
        mbar_wait(&bars->pa_full[w], bcount & 1);
        tc_fence_after();
        if (elect_one()) {
          const int ka = nch < 8 ? nch : 8;
          for (int k = 0; k < ka; ++k)
            umma_f16_ts(region + O_COL, region + p_col(k, h0), v_desc + 128 * k, idesc_o, (j > 0 || nch > 8 || k != 0) ? 1u : 0u);
          umma_commit(&bars->v_free[st]);
          if (j == p.n_blk - 1) umma_commit(&bars->pv_done[w]);
        }
        __syncwarp();
      }
      ++tcount;
    }
  } else if (warp < MMA_WARP0) {
    // ---------------------------------------------------------------------------------- softmax WG w
    const int w = warp >> 3;
    const int hf = (warp >> 2) & 1;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t t_row = tmem + w * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = 0.125f * 1.4426950408889634f;
    const uint64_t sc2 = pack_f32x2(sc, sc);
    const int pair_bar = 3 + w * 4 + quarter;
    uint8_t* stg = smem + OFF_OUT + warp * 2048;
    float* my_x = reinterpret_cast<float*>(smem + OFF_XCH) + (w * 2 + hf) * 128 + r;
    const float* other_x = reinterpret_cast<float*>(smem + OFF_XCH) + (w * 2 + (hf ^ 1)) * 128 + r;
    uint32_t pass_no = 0, bcount = 0, tcount = 0;
    const int n_iter = (p.n_pass - static_cast<int>(blockIdx.x) + g_stride - 1) / g_stride;
    const int rounds = n_iter * p.n_blk;  // ping-pong rounds of this CTA (both WGs walk all of them)
    int round = 0;
    if (p.pingpong && w == 1) asm volatile("bar.arrive 1, 512;" ::: "memory");  // WG 0 goes first
    for (int g = blockIdx.x; g < p.n_pass; g += g_stride, ++pass_no) {
      const int item = g / p.ppi, tp = g - item * p.ppi;
      const int tile = 2 * tp + w;
      if (tile >= p.n_tiles) {  // keep the hand-over protocol (and the Q release count) balanced
        for (int j = 0; j < p.n_blk; ++j, ++round) {
          if (p.pingpong) {
            if (w == 0) {
              asm volatile("bar.sync 1, 512;" ::: "memory");
              asm volatile("bar.arrive 2, 512;" ::: "memory");
            } else {
              asm volatile("bar.sync 2, 512;" ::: "memory");
              if (round != rounds - 1) asm volatile("bar.arrive 1, 512;" ::: "memory");
            }
          }
        }
        if (XKEY) {
          mbar_wait(&bars->q_full, pass_no & 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->q_free);
        }
        continue;
      }
      const bool warp_live = tile * 128 + quarter * 32 < p.L;
      float m_run = -INFINITY, nref = 0.0f, s_x = 0.0f;
      uint64_t acc_a = 0, acc_b = 0;  // this thread's partial row sum, relative to m_run
      for (int j = 0; j < p.n_blk; ++j, ++bcount, ++round) {
        const int nch = block_chunks(p, j), h0 = (nch + 1) >> 1;
        const int c0 = hf ? h0 : 0, c1 = hf ? nch : h0;
        const int ca = min(c1, 8), cb = max(c0, 8);
        const int l_blk = min(KB, p.Lk - j * KB);  // valid keys of the block
        const int jmax = l_blk - 1;
        const bool last = j == p.n_blk - 1;
        if (XKEY && last) {  // q_i . k_x: this thread's 32 of the 64 dimensions (q from the Q tile in shared memory)
          mbar_wait(&bars->q_full, pass_no & 1);
          const uint8_t* qrow = smem + OFF_Q + w * Q_BYTES + r * 128;
          const uint8_t* xk = smem + OFF_XROW + 64 * hf;
          float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 a = *reinterpret_cast<const uint4*>(qrow + (((4 * hf + c) ^ (r & 7)) << 4));
            const uint4 b = *reinterpret_cast<const uint4*>(xk + 16 * c);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[e]));
              const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&bw[e]));
              d0 = fmaf(fa.x, fb.x, d0);
              d1 = fmaf(fa.y, fb.y, d1);
            }
          }
          s_x = d0 + d1;
        }
        mbar_wait(&bars->s_full[w], bcount & 1);
        tc_fence_after();
        if (XKEY && last) {  // slot writes are ordered behind s_full (the partner read its previous value before O was freed)
          *my_x = s_x;
          pair_bar_sync(pair_bar);
          s_x = hf ? *other_x + s_x : s_x + *other_x;
          pair_bar_sync(pair_bar);
        }
        // ---- block maximum: own chunks, the partner's through shared memory, the extra key
        float mx = -INFINITY;
        if (warp_live) mx = row_max<false>(t_row, c0, c1, l_blk, jmax);
        if (XKEY && last) mx = fmaxf(mx, s_x);
        *my_x = mx;
        pair_bar_sync(pair_bar);
        mx = fmaxf(mx, *other_x);
        pair_bar_sync(pair_bar);
        // ---- online softmax with a lazy reference: move it (and rescale sum and O) only for a clearly larger block
        const float m_new = (j == 0 || (mx - m_run) * sc > LAZY_LOG2) ? fmaxf(mx, m_run) : m_run;
        if (j > 0 && __any_sync(0xffffffffu, m_new != m_run)) {
          const float alpha = (m_new == m_run) ? 1.0f : ex2_approx((m_run - m_new) * sc);
          const uint64_t al2 = pack_f32x2(alpha, alpha);
          acc_a = fma_f32x2(acc_a, al2, 0ull);
          acc_b = fma_f32x2(acc_b, al2, 0ull);
          if (warp_live) {  // P V_{j-1} has retired (S_j was issued behind it): O is quiescent until this block's P V
            uint32_t O2[32];
            tmem_ld_32x32(t_row + O_COL + 32 * hf, O2);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) O2[e] = __float_as_uint(__uint_as_float(O2[e]) * alpha);
            tmem_st_32x32(t_row + O_COL + 32 * hf, O2);
            tmem_wait_st();
          }
        }
        m_run = m_new;
        nref = (m_run == -INFINITY) ? 0.0f : -m_run * sc;
        const uint64_t nref2 = pack_f32x2(nref, nref);
        if (p.pingpong) {
          if (w == 0) asm volatile("bar.sync 1, 512;" ::: "memory");
          else asm volatile("bar.sync 2, 512;" ::: "memory");
        }
        // ---- exponentials: part b first, then part a (P in place, attn_wholerow.cuh)
        if (hf == 1 && nch > 8) {
          if (warp_live) {
            exp_chunks<false, 0>(t_row, cb, c1, p_col(cb, h0), l_blk, jmax, sc2, nref2, acc_a, acc_b);
            tmem_wait_st();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->pb_full[w]);
        }
        if (warp_live) {
          exp_chunks<false, 0>(t_row, c0, ca, p_col(c0, h0), l_blk, jmax, sc2, nref2, acc_a, acc_b);
          tmem_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->pa_full[w]);
        if (p.pingpong) {
          if (w == 0) asm volatile("bar.arrive 2, 512;" ::: "memory");
          else if (round != rounds - 1) asm volatile("bar.arrive 1, 512;" ::: "memory");
        }
      }
      uint4 vx[4];
      if (XKEY) {  // this warp's reads of the Q tile and of the extra rows are done once v_x is in registers
        const uint8_t* xv = smem + OFF_XROW + 128 + 64 * hf;
#pragma unroll
        for (int c = 0; c < 4; ++c) vx[c] = *reinterpret_cast<const uint4*>(xv + 16 * c);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->q_free);
      }
      // ---- row sum = the two threads' partial sums (+ the extra key)
      const float part = (lo_f(acc_a) + hi_f(acc_a)) + (lo_f(acc_b) + hi_f(acc_b));
      *my_x = part;
      pair_bar_sync(pair_bar);
      float sum = hf ? *other_x + part : part + *other_x;
      pair_bar_sync(pair_bar);  // both partial sums read: the slots may be rewritten (the next pass's S is issued early,
                                // so unlike in attention6 nothing else orders the partner's read before the next write)
      float p_x = 0.0f;
      if (XKEY) {
        const float e = ex2_approx(fmaf(s_x, sc, nref));
        sum += e;
        p_x = __half2float(__float2half_rn(e));
      }
      mbar_wait(&bars->pv_done[w], tcount & 1);
      tc_fence_after();
      if (!warp_live) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->o_free[w]);
      } else {
        uint32_t O2[32];
        tmem_ld_32x32(t_row + O_COL + 32 * hf, O2);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->o_free[w]);
        if (XKEY) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t vw[4] = {vx[c].x, vx[c].y, vx[c].z, vx[c].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fv = __half22float2(*reinterpret_cast<const __half2*>(&vw[e]));
              O2[8 * c + 2 * e] = __float_as_uint(fmaf(p_x, fv.x, __uint_as_float(O2[8 * c + 2 * e])));
              O2[8 * c + 2 * e + 1] = __float_as_uint(fmaf(p_x, fv.y, __uint_as_float(O2[8 * c + 2 * e + 1])));
            }
          }
        }
        if (elect_one()) tma_store_wait_read<0>();  // the previous pass's store has drained this block
        __syncwarp();
        const float inv = __fdividef(1.0f, sum);
        uint8_t* my_row = stg + lane * 64;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int e = cc * 8;
          uint4 x;
          x.x = pack_half2(__uint_as_float(O2[e + 0]) * inv, __uint_as_float(O2[e + 1]) * inv);
          x.y = pack_half2(__uint_as_float(O2[e + 2]) * inv, __uint_as_float(O2[e + 3]) * inv);
          x.z = pack_half2(__uint_as_float(O2[e + 4]) * inv, __uint_as_float(O2[e + 5]) * inv);
          x.w = pack_half2(__uint_as_float(O2[e + 6]) * inv, __uint_as_float(O2[e + 7]) * inv);
          *reinterpret_cast<uint4*>(my_row + ((cc ^ ((lane >> 1) & 3)) << 4)) = x;
        }
        fence_async_smem();
        __syncwarp();
        if (elect_one()) {
          const int b = item / p.heads, h = item % p.heads;
          tma_store_3d(&tmO, stg, h * HEAD_DIM + 32 * hf, tile * 128 + quarter * 32, b);
          tma_store_commit();
        }
      }
      ++tcount;
    }
    if (elect_one()) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <bool XKEY>
int launch_variant7(int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmKV,
                    const CUtensorMap& tmO, const Params7& p) {
  static int configured[kMaxDevices];
  auto kern = attention7_kernel<XKEY>;
  PC_CHECK_CUDA(ensure_dynamic_smem(kern, smem_bytes, configured));
  PC_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(THREADS7), smem_bytes, stream, 1, tmQ, tmKV, tmO, p));
  return PC_OK;
}

int env_int7(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

// Unmasked sequences longer than attention6 takes (L > 257); PC_ATTN7=0 or PC_ATTN_IMPL=2 keep them on attention.cu's
// round-1 streaming kernel (A/B).
bool attention7_supports(int L, int causal) {
  static int on = -1;
  if (on < 0) on = env_int7("PC_ATTN7", 1) && env_int7("PC_ATTN_IMPL", 6) == 6;
  return on && !causal && L > 257;
}

int launch_attention7(const __half* qkv, __half* out, int B, int L, int heads, cudaStream_t stream) {
  const int d = heads * HEAD_DIM;
  Params7 p{};
  p.L = L;
  p.xkey = (L % KB == 1) ? 1 : 0;  // 577 = 3 x 192 + 1: no fourth block for one key
  p.Lk = L - p.xkey;
  p.heads = heads;
  p.d = d;
  p.items = B * heads;
  p.n_blk = (p.Lk + KB - 1) / KB;
  p.n_tiles = (L + 127) / 128;
  p.ppi = (p.n_tiles + 1) / 2;
  p.n_pass = p.items * p.ppi;
  static int pingpong = -1;
  if (pingpong < 0) pingpong = env_int7("PC_ATTN7_PINGPONG", 1);
  p.pingpong = pingpong;
  p.qkv = qkv;
  p.row_pitch = 3 * d;
  p.plane_pitch = HEAD_DIM;
  const int smem_bytes = OFF_BARS + static_cast<int>(sizeof(Bars7));
  static_assert(OFF_BARS + sizeof(Bars7) <= 227 * 1024, "attention7: shared memory budget");
  CUtensorMap tmQ, tmKV, tmO;
  const uint64_t rows = static_cast<uint64_t>(B) * L;
  const uint64_t row_pitch = static_cast<uint64_t>(3 * d) * 2;
  PC_TRY(make_tmap_f16_3d(&tmQ, qkv, 64, rows, 3 * heads, row_pitch, 128, 64, 128));
  PC_TRY(make_tmap_f16_3d(&tmKV, qkv, 64, rows, 3 * heads, row_pitch, 128, 64, KB));
  PC_TRY(make_tmap_f16_3d(&tmO, out, d, L, B, static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(L) * d * 2, 32, 32));
  const int sms = device_sm_count();
  const int grid = p.n_pass < sms ? p.n_pass : sms;
  if (p.xkey) PC_TRY((launch_variant7<true>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p)));
  else PC_TRY((launch_variant7<false>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p)));
  return PC_OK;
}

}  // namespace pc
