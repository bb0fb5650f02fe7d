// Multi-head self-attention core, softmax(Q K^T / sqrt(64) [+ causal mask]) V, for the CLIP towers
// (reference clip/model.py:173,183-185 -> nn.MultiheadAttention -> scaled_dot_product_attention; the text
// tower's additive mask clip/model.py:326-332 is exactly "j > i -> -inf", i.e. the causal flag here).
//
// Persistent kernel, one CTA per SM (it owns all 512 TMEM columns), 12 warps:
//   warps 0-3 / 4-7   softmax warpgroup ("WG") 0 / 1: one 128-row query tile in flight each, one query row per
//                     thread (row == TMEM lane). The key axis is streamed in blocks of 64 keys through a two-slot
//                     S ring in TMEM: while the threads run the softmax of block j, the tensor core already
//                     computes S of block j+1 and the PV product of block j-1, so the softmax warps only ever wait
//                     at tile boundaries. A block's 64 scores are read from TMEM once, kept in registers for the
//                     max and the exp2 pass, and P goes back IN PLACE as packed fp16 (tcgen05.st) to be consumed
//                     by the PV MMA straight from TMEM (A operand in TMEM, "ts" form): no shared-memory round trip
//                     and no proxy fence for P. Online softmax with a LAZY reference: the running reference
//                     maximum only moves (and O, accumulated in TMEM, is only rescaled) when a block's maximum
//                     exceeds it by more than 8 in the exp2 domain, so P stays below 2^8 and rescales are rare.
//   warp 8 / 9        MMA issuer of WG 0 / 1 (whole warp in the loop, one elected lane issues):
//                     S_0, S_1, PV_0, S_2, PV_1, ... (S_j into ring slot j & 1: the in-order tensor pipe guarantees
//                     that PV_{j-2} has consumed the slot). Q, K from 128B-swizzled smem, V as the MN-major B operand.
//   warp 10 / 11      TMA producers: the K / V rows of an item (slots shared by the two WGs) and WG 0's query tiles /
//                     WG 1's query tiles, cut straight out of the packed qkv activation [B*L, 3d].
// Work decomposition: a GROUP is one K/V set plus the (up to) two 128-row query tiles that use it:
//   L <= 128   ("split")  the two WGs take two different (image, head) items; a K/V slot holds both items' K/V
//   L  > 128              the two WGs take tiles 2t, 2t+1 of the same item and share its K/V
// TMEM region of a WG (256 columns): S / P ring slots at [0, 64) and [64, 128), O at [128, 192).
#include <stdlib.h>

#include "attn_common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace pc {

namespace {

constexpr int HEAD_DIM = 64;
constexpr int KVB = 64;          // keys per block (one S ring slot)
constexpr int MMA_WARP0 = 8;     // warps 0..7: the two softmax warpgroups
constexpr int TMA_WARP = 10;     // K/V of every group + Q tiles of WG 0
constexpr int TMA_WARP_Q1 = 11;  // Q tiles of WG 1 (its own warp: the two WGs' tile boundaries stay decoupled)
constexpr int ATT_THREADS = 12 * 32;
constexpr int Q_BYTES = 128 * 128;  // one query tile: 128 rows x 64 fp16
constexpr int O_COL = 128;          // O accumulator columns inside a WG's TMEM region
constexpr float RESCALE_LOG2 = 8.0f;  // move the reference maximum only when it is exceeded by 2^8

struct AttnParams {
  int L;           // tokens per sequence
  int lp16;        // L rounded up to 16
  int heads;
  int d;           // heads * 64
  int items;       // B * heads
  int m_tiles;     // ceil(L / 128)
  int split;       // 1: L <= 128, the WGs take different items
  int ppi;         // non-split: tile pairs per item = ceil(m_tiles / 2)
  int n_groups;
  int n_blk;       // key blocks per row = ceil(lp16 / 64)
  int kv_rows;     // rows per K/V TMA box
  int kv_boxes;    // boxes per K (V) load
  int n_slots;     // K/V slots (2 when they fit, else 1)
  int sub_bytes;   // split: byte offset of WG 1's K (V) inside a slot's K (V) region
  int kreg_bytes;  // bytes of a slot's K region; the V region follows
  int off_kv;      // smem offset of slot 0
  int off_stage;   // 8 x 4 KB output staging blocks (one per softmax warp: 32 rows x 128 B, 128B-swizzled)
  int off_bars;
  long long* trace;  // bring-up only (env PC_ATTN_TRACE=1): [group iteration][WG][8] clock64 samples of CTA 0
};

#define ATRACE(slot, cond)                                                                              \
  do {                                                                                                  \
    if (p.trace != nullptr && blockIdx.x == 0 && (cond) && git < 32) p.trace[(git * 2 + w) * 8 + (slot)] = clock64(); \
  } while (0)

struct AttnBars {
  uint64_t kv_full[2];     // per K/V slot: TMA bytes landed
  uint64_t kv_free[2];     // per K/V slot: both WGs' last PV MMA of the group retired (2 arrivals)
  uint64_t q_full[2];      // per WG
  uint64_t q_free[2];      // per WG: last S MMA of the tile retired
  uint64_t s_full[2][2];   // [WG][ring slot]: S block in TMEM
  uint64_t p_full[2][2];   // [WG][ring slot]: P block in TMEM (4 warp arrivals)
  uint64_t pv_done[2][2];  // [WG][block parity]: the PV MMA of a block retired
  uint64_t o_free[2];      // per WG: O read out, accumulator reusable (4 warp arrivals)
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

template <bool CAUSAL>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + p.off_bars);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform (uniform registers)
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
        mbar_init(&bars->o_free[i], 4);
        for (int s = 0; s < 2; ++s) {
          mbar_init(&bars->s_full[i][s], 1);
          mbar_init(&bars->p_full[i][s], 4);
          mbar_init(&bars->pv_done[i][s], 1);
        }
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
    uint32_t u = 0;                   // K/V units loaded so far
    uint32_t q_cnt0 = 0;              // Q tiles of WG 0 loaded so far
    const uint32_t kv_box_bytes = static_cast<uint32_t>(p.kv_rows) * 128;
    const uint32_t kv_bytes = kv_box_bytes * p.kv_boxes;  // one K (or V) of one item
    for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x, ++u) {
      const Job j0 = job_of(p, g, 0), j1 = job_of(p, g, 1);
      const int slot = (p.n_slots == 2) ? (u & 1) : 0;
      const uint32_t use = (p.n_slots == 2) ? (u >> 1) : u;  // how many times this slot was filled before
      uint8_t* kbuf = smem + p.off_kv + slot * 2 * p.kreg_bytes;
      uint8_t* vbuf = kbuf + p.kreg_bytes;
      mbar_wait(&bars->kv_free[slot], (use & 1) ^ 1);
      if (elect_one()) {
        const int n_items = p.split ? ((j0.active ? 1 : 0) + (j1.active ? 1 : 0)) : 1;
        mbar_arrive_expect_tx(&bars->kv_full[slot], 2 * kv_bytes * n_items);
        for (int w = 0; w < 2; ++w) {
          const Job& j = w ? j1 : j0;
          if (p.split ? !j.active : (w == 1)) continue;  // shared K/V: loaded once (j0 is always active)
          const int b = j.item / p.heads, h = j.item % p.heads;
          uint8_t* kd = kbuf + (p.split ? w * p.sub_bytes : 0);
          uint8_t* vd = vbuf + (p.split ? w * p.sub_bytes : 0);
          for (int x = 0; x < p.kv_boxes; ++x) {
            const int row = b * p.L + x * p.kv_rows;
            tma_load_2d(kd + x * kv_box_bytes, &tmKV, &bars->kv_full[slot], p.d + h * HEAD_DIM, row);
            tma_load_2d(vd + x * kv_box_bytes, &tmKV, &bars->kv_full[slot], 2 * p.d + h * HEAD_DIM, row);
          }
        }
      }
      __syncwarp();
      if (j0.active) {  // WG 0's query tile (after the K/V of the group is on its way)
        mbar_wait(&bars->q_free[0], (q_cnt0 & 1) ^ 1);
        if (elect_one()) {
          const int b = j0.item / p.heads, h = j0.item % p.heads;
          mbar_arrive_expect_tx(&bars->q_full[0], Q_BYTES);
          tma_load_2d(smem, &tmQ, &bars->q_full[0], h * HEAD_DIM, b * p.L + j0.tile * 128);
        }
        __syncwarp();
        ++q_cnt0;
      }
    }
  } else if (warp == TMA_WARP_Q1) {
    // ---------------------------------------------------------------------------------- Q producer of WG 1
    uint32_t q_cnt1 = 0;
    for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
      const Job j1 = job_of(p, g, 1);
      if (!j1.active) continue;
      mbar_wait(&bars->q_free[1], (q_cnt1 & 1) ^ 1);
      if (elect_one()) {
        const int b = j1.item / p.heads, h = j1.item % p.heads;
        mbar_arrive_expect_tx(&bars->q_full[1], Q_BYTES);
        tma_load_2d(smem + Q_BYTES, &tmQ, &bars->q_full[1], h * HEAD_DIM, b * p.L + j1.tile * 128);
      }
      __syncwarp();
      ++q_cnt1;
    }
  } else if (warp == MMA_WARP0 || warp == MMA_WARP0 + 1) {
    // ---------------------------------------------------------------------------------- MMA issuer of WG w
    const int w = warp - MMA_WARP0;
    const uint32_t region = tmem + w * 256;
    const uint64_t q_desc = umma_desc_kmajor_sw128(smem_u32(smem + w * Q_BYTES));
    const uint32_t idesc_o = umma_idesc_f16(128, HEAD_DIM, 0, 1);
    uint32_t u = 0, q_cnt = 0;
    uint32_t pf_cnt0 = 0, pf_cnt1 = 0;  // p_full phases consumed per ring slot
    for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x, ++u) {
      const Job j = job_of(p, g, w);
      const int slot = (p.n_slots == 2) ? (u & 1) : 0;
      const uint32_t use = (p.n_slots == 2) ? (u >> 1) : u;
      mbar_wait(&bars->kv_full[slot], use & 1);
      if (!j.active) {  // this WG sits the group out: release its share of the slot
        if (lane == 0) mbar_arrive(&bars->kv_free[slot]);
        __syncwarp();
        continue;
      }
      mbar_wait(&bars->q_full[w], q_cnt & 1);
      // Stagger the warpgroups: WG 1 starts its first tile when WG 0 has delivered its first P block, so that one WG
      // tends to be in its exponentials while the other is in its max / hand-off phase.
      if (w == 1 && q_cnt == 0) mbar_wait(&bars->p_full[0][0], 0);
      tc_fence_after();
      const uint32_t k_addr = smem_u32(smem + p.off_kv + slot * 2 * p.kreg_bytes + (p.split ? w * p.sub_bytes : 0));
      const uint32_t v_addr = k_addr + p.kreg_bytes;
      const uint64_t k_desc = umma_desc_kmajor_sw128(k_addr);
      const uint64_t v_desc = umma_desc_mnmajor_sw128(v_addr, 1024);
      // S_0; then for t = 1 .. n_blk: S_t (if any) followed by PV_{t-1}
      for (int t = 0; t <= p.n_blk; ++t) {
        if (t < p.n_blk) {
          const int n_cols = min(KVB, p.lp16 - t * KVB);
          if (elect_one()) {
            // S_t[128, n_cols] = Q K_t^T into ring slot t & 1 (key row r of K at byte r * 128: >> 4 in the descriptor)
            const uint32_t idesc_s = umma_idesc_f16(128, n_cols, 0, 0);
            const uint64_t kd = k_desc + static_cast<uint64_t>(t) * (KVB * 128 / 16);
#pragma unroll
            for (int k = 0; k < HEAD_DIM / 16; ++k)
              umma_f16_ss(region + (t & 1) * KVB, q_desc + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
            umma_commit(&bars->s_full[w][t & 1]);
            if (t == p.n_blk - 1) umma_commit(&bars->q_free[w]);
          }
          __syncwarp();
        }
        if (t >= 1) {
          const int jb = t - 1;
          const int n_cols = min(KVB, p.lp16 - jb * KVB);
          if (jb == 0) mbar_wait(&bars->o_free[w], (q_cnt & 1) ^ 1);  // the previous tile's O has been read out
          if (jb & 1) {
            mbar_wait(&bars->p_full[w][1], pf_cnt1 & 1);
            ++pf_cnt1;
          } else {
            mbar_wait(&bars->p_full[w][0], pf_cnt0 & 1);
            ++pf_cnt0;
          }
          tc_fence_after();
          if (elect_one()) {
            // O[128, 64] (+)= P_jb V_jb : P from TMEM (8 columns per 16 keys), V MN-major (16 key rows = 2048 B per step)
            const uint64_t vd = v_desc + static_cast<uint64_t>(jb) * (KVB * 128 / 16);
            const int k_steps = n_cols >> 4;
            for (int kk = 0; kk < k_steps; ++kk)
              umma_f16_ts(region + O_COL, region + (jb & 1) * KVB + 8 * kk, vd + 128 * kk, idesc_o, (jb | kk) != 0 ? 1u : 0u);
            umma_commit(&bars->pv_done[w][jb & 1]);
            if (jb == p.n_blk - 1) umma_commit(&bars->kv_free[slot]);
          }
          __syncwarp();
        }
      }
      ++q_cnt;
    }
  } else if (warp < MMA_WARP0) {
    // ---------------------------------------------------------------------------------- softmax WG w
    const int w = warp >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row of the tile == TMEM lane
    const uint32_t t_row = tmem + w * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint64_t sc2 = pack_f32x2(sc, sc);
    uint32_t sf_cnt0 = 0, sf_cnt1 = 0;  // s_full phases consumed per ring slot
    uint32_t pv_obs0 = 0, pv_obs1 = 0;  // pv_done phases observed per block parity
    uint32_t pv_exp0 = 0, pv_exp1 = 0;  // PV MMAs that follow the P blocks delivered so far, per block parity
    int git = -1;
    for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
      ++git;
      const Job j = job_of(p, g, w);
      if (!j.active) continue;
      const int i = j.tile * 128 + r;  // query index inside the sequence
      const bool warp_live = j.tile * 128 + quarter * 32 < p.L;  // some row of this warp is a real query
      const int jmax = CAUSAL ? min(i, p.L - 1) : p.L - 1;       // last key this row attends to
      float m_ref = -INFINITY;  // reference maximum of the row (raw score units)
      uint64_t acc2 = 0;        // running (even, odd) sums of p
      for (int jb = 0; jb < p.n_blk; ++jb) {
        const int s = jb & 1;
        const int n16 = min(KVB, p.lp16 - jb * KVB) >> 4;  // 16-key chunks of this block
        const int kbase = jb * KVB;
        if (jb == 0) ATRACE(0, quarter == 0 && lane == 0);
        if (s) {
          mbar_wait(&bars->s_full[w][1], sf_cnt1 & 1);
          ++sf_cnt1;
        } else {
          mbar_wait(&bars->s_full[w][0], sf_cnt0 & 1);
          ++sf_cnt0;
        }
        tc_fence_after();
        if (jb == 0) ATRACE(1, quarter == 0 && lane == 0);
        // S_jb complete => every PV of this parity delivered so far (PV_{jb-2}, ...) has retired, because it was issued
        // before S_jb on the in-order pipe: keep the phase bookkeeping of that barrier in step (no real waiting)
        if (s) {
          while (pv_obs1 < pv_exp1) {
            mbar_wait(&bars->pv_done[w][1], pv_obs1 & 1);
            ++pv_obs1;
          }
        } else {
          while (pv_obs0 < pv_exp0) {
            mbar_wait(&bars->pv_done[w][0], pv_obs0 & 1);
            ++pv_obs0;
          }
        }
        if (warp_live) {
          const uint32_t t_s = t_row + s * KVB;
          uint32_t R[4][16];
          // all four chunks are loaded (columns past the block's end hold stale data that is never used): straight-
          // line loads keep the 64 scores in registers
#pragma unroll
          for (int k = 0; k < 4; ++k) tmem_ld_32x16(t_s + k * 16, R[k]);
          tmem_wait_ld();
          // blocks whose every key is valid for every row of the warp need no masking (never for causal rows)
          const bool full = !CAUSAL && kbase + n16 * 16 <= p.L;
          // ---- block maximum
          float mx = -INFINITY;
          if (full) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < n16) mx = chunk_max<true>(R[k], 0, mx);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < n16) mx = chunk_max<false>(R[k], jmax - kbase - k * 16, mx);
          }
          // ---- lazy reference update: only when the block exceeds the reference by more than 2^8
          if (jb == 0) {
            m_ref = mx;
          } else {
            const bool grow = (mx - m_ref) * sc > RESCALE_LOG2;  // false for NaN / (-inf) - (-inf)
            if (__any_sync(0xffffffffu, grow)) {
              // O and the running sums follow the new reference. PV_{jb-1} (other parity) must have retired first.
              if (s) {
                while (pv_obs0 < pv_exp0) {
                  mbar_wait(&bars->pv_done[w][0], pv_obs0 & 1);
                  ++pv_obs0;
                }
              } else {
                while (pv_obs1 < pv_exp1) {
                  mbar_wait(&bars->pv_done[w][1], pv_obs1 & 1);
                  ++pv_obs1;
                }
              }
              tc_fence_after();
              const float alpha = grow ? ex2_approx((m_ref - mx) * sc) : 1.0f;  // m_ref = -inf -> 0
              if (grow) m_ref = mx;
              uint32_t o[16];
#pragma unroll
              for (int hh = 0; hh < 4; ++hh) {
                tmem_ld_32x16(t_row + O_COL + 16 * hh, o);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                tmem_st_32x16(t_row + O_COL + 16 * hh, o);
              }
              acc2 = fma_f32x2(acc2, pack_f32x2(alpha, alpha), 0);
            }
          }
          // ---- p = exp2((s - ref) / 8 * log2 e), fp16 P written over the slot's own S columns
          const float nref = (m_ref == -INFINITY) ? 0.0f : -m_ref * sc;
          const uint64_t nref2 = pack_f32x2(nref, nref);
          uint32_t pk[8];
          if (full) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < n16) {
                acc2 = chunk_exp<true>(R[k], pk, 0, sc2, nref2, acc2);
                tmem_st_32x8(t_s + k * 8, pk);
              }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < n16) {
                acc2 = chunk_exp<false>(R[k], pk, jmax - kbase - k * 16, sc2, nref2, acc2);
                tmem_st_32x8(t_s + k * 8, pk);
              }
          }
          tmem_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_full[w][s]);
        if (s) ++pv_exp1;  // the issuer follows this arrival with PV_jb
        else ++pv_exp0;
      }
      ATRACE(3, quarter == 0 && lane == 0);
      // ---- all PV MMAs of the tile retired -> O / sum -> fp16 -> out[b, i, h*64 .. h*64+63]
      while (pv_obs0 < pv_exp0) {
        mbar_wait(&bars->pv_done[w][0], pv_obs0 & 1);
        ++pv_obs0;
      }
      while (pv_obs1 < pv_exp1) {
        mbar_wait(&bars->pv_done[w][1], pv_obs1 & 1);
        ++pv_obs1;
      }
      tc_fence_after();
      ATRACE(4, quarter == 0 && lane == 0);
      uint32_t O4[4][16];
      if (warp_live) {
#pragma unroll
        for (int k = 0; k < 4; ++k) tmem_ld_32x16(t_row + O_COL + 16 * k, O4[k]);
        tmem_wait_ld();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->o_free[w]);
      if (warp_live) {
        // fp16 rows into this warp's swizzled staging block (16-byte chunk c of row r at chunk c ^ (r & 7)), then
        // one TMA store of the 32 x 64 block through the [B][L][d] map: rows past the sequence end are clipped.
        uint8_t* stg = smem + p.off_stage + warp * 4096;
        if (elect_one()) tma_store_wait_read<0>();  // the previous tile's store has drained this block
        __syncwarp();
        const float sum = __uint_as_float(static_cast<uint32_t>(acc2)) + __uint_as_float(static_cast<uint32_t>(acc2 >> 32));
        const float inv = __fdividef(1.0f, sum);
        uint8_t* my_row = stg + lane * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            uint4 x;
            x.x = pack_half2(__uint_as_float(O4[k][8 * e + 0]) * inv, __uint_as_float(O4[k][8 * e + 1]) * inv);
            x.y = pack_half2(__uint_as_float(O4[k][8 * e + 2]) * inv, __uint_as_float(O4[k][8 * e + 3]) * inv);
            x.z = pack_half2(__uint_as_float(O4[k][8 * e + 4]) * inv, __uint_as_float(O4[k][8 * e + 5]) * inv);
            x.w = pack_half2(__uint_as_float(O4[k][8 * e + 6]) * inv, __uint_as_float(O4[k][8 * e + 7]) * inv);
            *reinterpret_cast<uint4*>(my_row + (((2 * k + e) ^ (lane & 7)) << 4)) = x;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (elect_one()) {
          const int b = j.item / p.heads, h = j.item % p.heads;
          tma_store_3d(&tmO, stg, h * HEAD_DIM, j.tile * 128 + quarter * 32, b);
          tma_store_commit();
        }
      }
      ATRACE(5, quarter == 0 && lane == 0);
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

// The last L % 128 query rows of every sequence when that remainder is tiny (ViT-L/14: L = 257 = 2 x 128 + 1): a third
// 128-row tile would cost a whole pass of its warpgroup over every key block for one live row (measured: 192 TFLOP/s
// at L = 257 against 320 at L = 577). One CTA per (sequence, head, row), one warp per 32 keys: a lane scores one key
// (a 64-wide dot product of a 128-byte row), the warp forms its partial (max, sum, P V) with the 32 V-row loads in
// flight at once, the partials meet in shared memory (online-softmax merge). fp32 throughout.
constexpr int TAIL_MAX_ROWS = 8;
constexpr int TAIL_MAX_WARPS = 32;
__global__ void __launch_bounds__(TAIL_MAX_WARPS * 32)
attention_tail_rows_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int B, int L, int heads, int row0,
                           int nrows, int causal, size_t plane_stride, int row_pitch, int out_seq_rows) {
  __shared__ float qs[64];
  __shared__ float part[TAIL_MAX_WARPS][66];  // per warp: max, sum, o[64]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // q / k / v of head h: plane (w * heads + h) at plane_stride elements, row r at r * row_pitch inside it. Packed
  // [B*L, 3d]: plane_stride 64, row_pitch 3d; planar [3 * heads][B*L][64]: plane_stride B*L*64, row_pitch 64.
  // Output row of (sequence b, query i): b * out_seq_rows + (i - row0) when out_seq_rows < L (compact), else b*L + i.
  const int d = heads * HEAD_DIM;
  griddep_launch_dependents();
  griddep_wait();
  const int item = blockIdx.x;
  const int i = row0 + item % nrows, bh = item / nrows, b = bh / heads, h = bh % heads;
  const int nkeys = causal ? i + 1 : L;
  const size_t seq0 = static_cast<size_t>(b) * L;
  if (w == 0) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(
        qkv + static_cast<size_t>(h) * plane_stride + (seq0 + i) * row_pitch + 2 * lane));
    qs[2 * lane] = f.x * 0.125f;
    qs[2 * lane + 1] = f.y * 0.125f;
  }
  __syncthreads();
  const __half* kbase = qkv + static_cast<size_t>(heads + h) * plane_stride + seq0 * row_pitch;
  const __half* vbase = qkv + static_cast<size_t>(2 * heads + h) * plane_stride + seq0 * row_pitch;
  // scores of keys [32 w, 32 w + 32): eight lanes share a key row (16 bytes each: every load instruction reads four
  // whole 128-byte rows), eight rounds of four keys; the 8-lane partial dot products meet by shuffles
  __shared__ float sc_s[TAIL_MAX_WARPS][32];
  const int part8 = lane & 7, kq = lane >> 3;
  float qv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) qv[e] = qs[8 * part8 + e];
  uint4 kk[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int jj = min(32 * w + 4 * t + kq, nkeys - 1);
    kk[t] = *(reinterpret_cast<const uint4*>(kbase + static_cast<size_t>(jj) * row_pitch) + part8);
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const __half2* h2 = reinterpret_cast<const __half2*>(&kk[t]);
    float acc = 0.0f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(h2[e]);
      acc = fmaf(f.x, qv[2 * e], fmaf(f.y, qv[2 * e + 1], acc));
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (part8 == 0) sc_s[w][4 * t + kq] = acc;
  }
  __syncwarp();
  const int j = 32 * w + lane;
  const float sc = (j < nkeys) ? sc_s[w][lane] : -INFINITY;
  const float mw = warp_max(sc);
  const float pj = (j < nkeys) ? __expf(sc - mw) : 0.0f;
  const float lw = warp_sum(pj);
  float o0 = 0.0f, o1 = 0.0f;
  if (32 * w < nkeys) {
    __half2 vv[32];
#pragma unroll
    for (int u = 0; u < 32; ++u)
      vv[u] = *reinterpret_cast<const __half2*>(vbase + static_cast<size_t>(min(32 * w + u, nkeys - 1)) * row_pitch + 2 * lane);
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      const float pu = __shfl_sync(0xffffffffu, pj, u);  // 0 for keys past nkeys
      const float2 f = __half22float2(vv[u]);
      o0 = fmaf(pu, f.x, o0);
      o1 = fmaf(pu, f.y, o1);
    }
  }
  if (lane == 0) {
    part[w][0] = mw;
    part[w][1] = lw;
  }
  part[w][2 + 2 * lane] = o0;
  part[w][3 + 2 * lane] = o1;
  __syncthreads();
  if (w == 0) {
    float m = -INFINITY;
    for (int k = 0; k < nwarps; ++k) m = fmaxf(m, part[k][0]);
    float l = 0.0f, a0 = 0.0f, a1 = 0.0f;
    for (int k = 0; k < nwarps; ++k) {
      const float sck = (part[k][0] == -INFINITY) ? 0.0f : __expf(part[k][0] - m);
      l = fmaf(part[k][1], sck, l);
      a0 = fmaf(part[k][2 + 2 * lane], sck, a0);
      a1 = fmaf(part[k][3 + 2 * lane], sck, a1);
    }
    const float inv = 1.0f / l;
    const size_t orow = out_seq_rows < L ? static_cast<size_t>(b) * out_seq_rows + (i - row0) : seq0 + i;
    *reinterpret_cast<__half2*>(out + orow * d + h * HEAD_DIM + 2 * lane) = __floats2half2_rn(a0 * inv, a1 * inv);
  }
}

template <bool CAUSAL>
int launch_variant(int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmKV,
                   const CUtensorMap& tmO, const AttnParams& p) {
  static int configured[kMaxDevices];
  auto kern = attention_kernel<CAUSAL>;
  PC_CHECK_CUDA(ensure_dynamic_smem(kern, smem_bytes, configured));
  PC_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(ATT_THREADS), smem_bytes, stream, 1, tmQ, tmKV, tmO, p));
  return PC_OK;
}

}  // namespace

// Attention of query rows [row0, row0 + nrows) of every sequence only (nrows small), against all L keys; out is compact
// [B * nrows, d]. The last ViT block needs the CLS row alone (clip/model.py:233 keeps x[:, 0, :]).
int launch_attention_rows(const __half* qkv, int qkv_planar, __half* out, int B, int L, int heads, int row0, int nrows,
                          int causal, cudaStream_t stream) {
  const int warps = (L + 31) / 32;
  PC_REQUIRE(qkv && out && B > 0 && L > 0 && heads > 0 && nrows >= 1 && row0 >= 0 && row0 + nrows <= L && nrows < L &&
                 warps <= TAIL_MAX_WARPS,
             PC_ERR_ARG, "attention_rows: bad arguments (L = %d, rows [%d, %d))", L, row0, row0 + nrows);
  const int d = heads * HEAD_DIM;
  const size_t plane_stride = qkv_planar ? static_cast<size_t>(B) * L * HEAD_DIM : HEAD_DIM;
  const int row_pitch = qkv_planar ? HEAD_DIM : 3 * d;
  PC_CHECK_CUDA(launch_pdl(attention_tail_rows_kernel, dim3(B * heads * nrows), dim3(warps * 32), 0, stream, 1, qkv, out,
                           B, L, heads, row0, nrows, causal, plane_stride, row_pitch, nrows));
  return PC_OK;
}

int launch_attention_layout(const __half* qkv, int qkv_planar, __half* out, int B, int L, int heads, int causal,
                            cudaStream_t stream) {
  PC_REQUIRE(qkv && out && B > 0 && L > 0 && heads > 0, PC_ERR_ARG, "attention: bad arguments");
  if (attention6_supports(L)) return launch_attention6(qkv, qkv_planar, out, B, L, heads, causal, stream);
  PC_REQUIRE(!qkv_planar, PC_ERR_ARG, "attention: the planar qkv layout is only read by the L <= 208 kernel (L = %d)", L);
  return launch_attention(qkv, out, B, L, heads, causal, stream);
}

int launch_attention(const __half* qkv, __half* out, int B, int L, int heads, int causal,
                     cudaStream_t stream) {
  PC_REQUIRE(qkv && out && B > 0 && L > 0 && heads > 0, PC_ERR_ARG, "attention: bad arguments");
  if (attention6_supports(L)) return launch_attention6(qkv, 0, out, B, L, heads, causal, stream);  // whole-row S in TMEM
  // ViT-L/14 (L = 257): 256 x 256 whole-row + key 256 in the softmax threads + query row 256 in the kernel's spare warp
  if (attention6_supports_xkey(L, causal)) return launch_attention6(qkv, 0, out, B, L, heads, causal, stream);
  if (attention7_supports(L, causal)) return launch_attention7(qkv, out, B, L, heads, stream);  // 192-key blocks, O in TMEM
  if (attention5_supports(L)) return launch_attention5(qkv, out, B, L, heads, causal, stream);  // round-1 kernel (A/B)
  const int d = heads * HEAD_DIM;
  AttnParams p{};
  p.L = L;
  p.lp16 = (L + 15) / 16 * 16;
  p.heads = heads;
  p.d = d;
  p.items = B * heads;
  p.m_tiles = (L + 127) / 128;
  // a tiny last tile goes to attention_tail_rows_kernel instead (see there); PC_ATTN_NO_TAIL=1 switches that off
  static int no_tail = -1;
  if (no_tail < 0) {
    const char* e = getenv("PC_ATTN_NO_TAIL");
    no_tail = e ? atoi(e) : 0;  // 1: third tile in the main kernel; 2 (timing only, wrong results): tail rows skipped
  }
  const int tail_rows = (L > 256 && L % 128 >= 1 && L % 128 <= TAIL_MAX_ROWS && no_tail != 1) ? L % 128 : 0;
  const int tail_warps = (L + 31) / 32;  // one warp per 32 keys
  const bool use_tail = tail_rows > 0 && tail_warps <= TAIL_MAX_WARPS;
  if (use_tail) p.m_tiles -= 1;
  p.split = p.m_tiles == 1 ? 1 : 0;
  p.ppi = (p.m_tiles + 1) / 2;
  p.n_groups = p.split ? (p.items + 1) / 2 : p.items * p.ppi;
  p.n_blk = (p.lp16 + KVB - 1) / KVB;
  // K / V of one item: lp16 rows of 128 bytes, fetched in boxes of at most 256 rows
  p.kv_boxes = (p.lp16 + 255) / 256;
  p.kv_rows = ((p.lp16 + p.kv_boxes - 1) / p.kv_boxes + 7) / 8 * 8;
  const int item_bytes = p.kv_rows * p.kv_boxes * 128;  // multiple of 1024 (swizzle atoms)
  p.sub_bytes = item_bytes;
  p.kreg_bytes = (p.split ? 2 : 1) * item_bytes;
  p.off_kv = 2 * Q_BYTES;
  const int fixed = 2 * Q_BYTES + 8 * 4096 + static_cast<int>(sizeof(AttnBars)) + 256;
  p.n_slots = (fixed + 4 * p.kreg_bytes <= 227 * 1024) ? 2 : 1;
  p.off_stage = p.off_kv + p.n_slots * 2 * p.kreg_bytes;
  p.off_bars = p.off_stage + 8 * 4096;
  int smem_bytes = p.off_bars + static_cast<int>(sizeof(AttnBars));
  PC_REQUIRE(smem_bytes <= 227 * 1024, PC_ERR_ARG,
             "attention: L = %d needs %d B of shared memory for one item's K and V (limit 227 KB)", L, smem_bytes);
  // one CTA per SM by construction (each CTA owns all 512 TMEM columns): ask for more than half the SM's smem
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;

  CUtensorMap tmQ, tmKV, tmO;
  const uint64_t rows = static_cast<uint64_t>(B) * L;
  PC_TRY(make_tmap_f16_2d(&tmQ, qkv, 3 * d, rows, static_cast<uint64_t>(3 * d) * 2, 64, 128));
  PC_TRY(make_tmap_f16_2d(&tmKV, qkv, 3 * d, rows, static_cast<uint64_t>(3 * d) * 2, 64, p.kv_rows));
  PC_TRY(make_tmap_f16_3d(&tmO, out, d, L, B, static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(L) * d * 2, 64, 32));
  const int sms = device_sm_count();
  const int grid = p.n_groups < sms ? p.n_groups : sms;
  static int tracing = -1;
  static long long* trace = nullptr;
  if (tracing < 0) {
    const char* e = getenv("PC_ATTN_TRACE");
    tracing = (e && e[0] == '1') ? 1 : 0;
  }
  if (tracing) {
    if (!trace) PC_CHECK_CUDA(cudaMalloc(&trace, 32 * 2 * 8 * sizeof(long long)));
    PC_CHECK_CUDA(cudaMemsetAsync(trace, 0, 32 * 2 * 8 * sizeof(long long), stream));
    p.trace = trace;
  }
  PC_TRY(causal ? launch_variant<true>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p)
                : launch_variant<false>(grid, smem_bytes, stream, tmQ, tmKV, tmO, p));
  if (use_tail && no_tail != 2) {
    PC_CHECK_CUDA(launch_pdl(attention_tail_rows_kernel, dim3(B * heads * tail_rows), dim3(tail_warps * 32), 0, stream, 1,
                             qkv, out, B, L, heads, L - tail_rows, tail_rows, causal, static_cast<size_t>(HEAD_DIM), 3 * d, L));
  }
  if (tracing) {
    static int printed = 0;
    long long h[32 * 2 * 8];
    PC_CHECK_CUDA(cudaStreamSynchronize(stream));
    PC_CHECK_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    if (printed++ == 3) {
      const long long t0 = h[0] ? h[0] : h[8];
      fprintf(stderr, "[attn trace] B*heads=%d L=%d (cycles since first sample, CTA 0)\n", p.items, L);
      fprintf(stderr, "group WG  wait_s0  s0_full  last_p_arrive  pv_all_done  stored\n");
      for (int g = 0; g < 32; ++g)
        for (int w = 0; w < 2; ++w) {
          const long long* r = h + (g * 2 + w) * 8;
          if (!r[1]) continue;
          fprintf(stderr, "%4d  %d %8lld %8lld %14lld %12lld %7lld\n", g, w, r[0] - t0, r[1] - t0, r[3] - t0, r[4] - t0,
                  r[5] - t0);
        }
    }
  }
  return PC_OK;
}

}  // namespace pc
