// Multi-head self-attention core for sequences of at most 256 tokens (ViT-B/32, ViT-B/16, the text tower, the
// ModifiedResNet attention pool): FOUR query tiles in flight per SM. Same contract as attention.cu's kernel
// (reference clip/model.py:173,183-185; causal flag = the text mask of clip/model.py:326-332).
//
// Why a second kernel: with two query tiles per SM (attention.cu) every scheduler runs two softmax warps whose
// dependent chain (S ready -> tcgen05.ld -> row max -> exp2 -> tcgen05.st -> P ready -> PV -> next S) leaves the MUFU
// pipe idle about half of the time, and the tile epilogue is not overlapped at all. Here each of the four
// schedulers runs FOUR independent softmax warps (four warpgroups, one 128-row query tile each), so that one
// warpgroup's MMA round trips and epilogue are covered by the other three.
//
// Persistent kernel, one CTA per SM, 24 warps:
//   warps 0-15   softmax warpgroups ("WG") 0..3, one query row per thread (row == TMEM lane). A WG owns 128 TMEM
//                columns: one 64-key S / P slot at [0, 64) and the O accumulator at [64, 128). Per 64-key block:
//                pass 1 reads the scores for the row maximum, pass 2 re-reads them 16 at a time, p = exp2(...) goes back
//                IN PLACE as packed fp16 and is consumed by the PV MMA straight from TMEM. Online softmax with a lazy
//                reference (moves only when exceeded by 2^8; O is rescaled in TMEM then). The WG also issues the TMA
//                load of its own next query tile as soon as the last S MMA of the current one has retired; output rows
//                leave through a 2 KB per-warp staging block and TMA stores.
//   warps 16-19  MMA issuer of WG 0..3: S_0, then per block PV_j followed by S_{j+1} (same slot: the in-order tensor
//                pipe guarantees PV_j has consumed P_j). Whole warp in the loop, one elected lane issues.
//   warps 20-21  K / V producer of CHANNEL 0 / 1. A channel is a pair of WGs (2c, 2c+1) sharing one K/V buffer:
//                the two query tiles of an item (128 < L <= 256) or two different items (L <= 128, "split"). K and V
//                are staged per 64-key block, each block slot with its own full / free barriers: block j of the NEXT
//                item is fetched as soon as the S (K) or PV (V) MMAs of both WGs on block j of the current item have
//                retired, so the loads run almost one item ahead of the math (the kernel is otherwise bound by
//                the latency of its own K/V fetches: with the math switched off it runs barely faster).
#include <stdlib.h>

#include "attn_common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace pc {
namespace {

constexpr int HEAD_DIM = 64;
constexpr int KVB = 64;        // keys per block
constexpr int MMA_WARP0 = 16;  // warps 0..15: four softmax warpgroups
constexpr int KV_WARP0 = 20;   // warps 20, 21: K/V producers of channel 0 / 1 (22, 23 idle: warpgroup padding)
constexpr int THREADS5 = 24 * 32;
constexpr int Q_BYTES = 128 * 128;  // one query tile: 128 rows x 64 fp16
constexpr int O_COL = 64;           // O accumulator columns inside a WG's 128-column TMEM region
constexpr float RESCALE_LOG2 = 8.0f;
// K (and V) of a channel live in four 64-key block slots of 8 KB: the four blocks of the shared item (128 < L <= 256)
// or two blocks for each of the channel's two items (L <= 128). Every slot has its own full / free barriers, so the
// next item's block j is fetched as soon as both users of the current block j are done with it: the loads run
// almost a whole item ahead of the math.
constexpr int SLOT_BYTES = KVB * 128;
constexpr int KREG_BYTES = 4 * SLOT_BYTES;

struct Params5 {
  int L, lp16, heads, d, items;
  int split;       // 1: L <= 128, the two WGs of a channel take different items
  int n_groups;    // split: ceil(items / 2), else items
  int n_blk;       // key blocks per row
  int tail_rows;   // key rows of the last block (multiple of 16, <= 64)
  int off_kv;      // channel c at off_kv + c * 2 * KREG_BYTES: K region, then V region
  int off_stage;   // 16 x 2 KB output staging blocks (one per softmax warp: 32 rows x 64 B, 64B-swizzled)
  int off_bars;
  long long* trace;  // bring-up only (env PC_ATTN_TRACE=1): [tile][32] clock64 samples of CTA 0, WG 0, warp 0
  int debug;       // bring-up only (env PC_ATTN5_DEBUG): 1 = no exp2 (MUFU off), 2 = no MMA issue (wrong results: timing A/B only)
};

#define TR5(slot)                                                                                       \
  do {                                                                                                  \
    if (p.trace != nullptr && blockIdx.x == 0 && warp == 0 && lane == 0 && tcount < 8) p.trace[tcount * 32 + (slot)] = clock64(); \
  } while (0)

struct Bars5 {
  uint64_t k_full[2][4], v_full[2][4];  // [channel][block slot]: TMA bytes landed
  uint64_t k_free[2][4], v_free[2][4];  // [channel][block slot]: the S / PV MMAs reading the slot retired
                                        // (2 arrivals when the WGs of the channel share the item, else 1)
  uint64_t q_full[4];             // per WG
  uint64_t s_full[4];             // per WG: S block in TMEM (and every earlier MMA of the WG retired)
  uint64_t p_full[4];             // per WG: P block in TMEM (4 warp arrivals)
  uint64_t pv_done[4];            // per WG: last PV MMA of the tile retired
  uint64_t o_free[4];             // per WG: O read out (4 warp arrivals)
  uint32_t tmem_base;
};

struct Job {
  bool active;
  int item;  // b * heads + h
  int tile;
};
__device__ __forceinline__ Job job_of(const Params5& p, int g, int s) {
  Job j;
  if (p.split) {
    j.item = 2 * g + s;
    j.tile = 0;
    j.active = j.item < p.items;
  } else {
    j.item = g;
    j.tile = s;
    j.active = true;
  }
  return j;
}
// first group >= g (stepping by `stride`) in which WG sub-index s has work; -1 if none
__device__ __forceinline__ int next_active(const Params5& p, int g, int stride, int s) {
  for (; g < p.n_groups; g += stride)
    if (job_of(p, g, s).active) return g;
  return -1;
}

template <bool CAUSAL>
__global__ void __launch_bounds__(THREADS5, 1)
attention5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                  const __grid_constant__ CUtensorMap tmKVt, const __grid_constant__ CUtensorMap tmO, const Params5 p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars5* bars = reinterpret_cast<Bars5*>(smem + p.off_bars);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;

  if (warp == KV_WARP0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      tma_prefetch_desc(&tmKVt);
      tma_prefetch_desc(&tmO);
      for (int i = 0; i < 2; ++i)
        for (int sl = 0; sl < 4; ++sl) {
          mbar_init(&bars->k_full[i][sl], 1);
          mbar_init(&bars->v_full[i][sl], 1);
          mbar_init(&bars->k_free[i][sl], p.split ? 1 : 2);
          mbar_init(&bars->v_free[i][sl], p.split ? 1 : 2);
        }
      for (int i = 0; i < 4; ++i) {
        mbar_init(&bars->q_full[i], 1);
        mbar_init(&bars->s_full[i], 1);
        mbar_init(&bars->p_full[i], 4);
        mbar_init(&bars->pv_done[i], 1);
        mbar_init(&bars->o_free[i], 4);
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

  const int g_stride = 2 * gridDim.x;
  if (warp == KV_WARP0 || warp == KV_WARP0 + 1) {
    // ---------------------------------------------------------------------------------- K / V producer of channel c
    const int c = warp - KV_WARP0;
    uint8_t* kreg = smem + p.off_kv + c * 2 * KREG_BYTES;
    uint32_t u = 0;
    for (int g = 2 * blockIdx.x + c; g < p.n_groups; g += g_stride, ++u) {
      for (int jb = 0; jb < p.n_blk; ++jb) {
        const bool tail = jb == p.n_blk - 1;
        const uint32_t bytes = static_cast<uint32_t>(tail ? p.tail_rows : KVB) * 128;
#pragma unroll 1
        for (int kv = 0; kv < 2; ++kv) {      // K_jb (released first), then V_jb
          for (int s = 0; s < 2; ++s) {
            if (!p.split && s == 1) break;    // shared item: loaded once
            const Job j = job_of(p, g, s);
            if (!j.active) continue;
            const int slot = p.split ? s * 2 + jb : jb;
            mbar_wait(kv ? &bars->v_free[c][slot] : &bars->k_free[c][slot], (u & 1) ^ 1);
            if (elect_one()) {
              uint64_t* full = kv ? &bars->v_full[c][slot] : &bars->k_full[c][slot];
              mbar_arrive_expect_tx(full, bytes);
              tma_load_2d(kreg + kv * KREG_BYTES + slot * SLOT_BYTES, tail ? &tmKVt : &tmKV, full,
                          (1 + kv) * p.d + (j.item % p.heads) * HEAD_DIM, (j.item / p.heads) * p.L + jb * KVB);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= MMA_WARP0 && warp < MMA_WARP0 + 4) {
    // ---------------------------------------------------------------------------------- MMA issuer of WG w
    const int w = warp - MMA_WARP0, c = w >> 1, s = w & 1;
    const uint32_t region = tmem + w * 128;
    const uint64_t q_desc = umma_desc_kmajor_sw128(smem_u32(smem + w * Q_BYTES));
    const uint32_t idesc_o = umma_idesc_f16(128, HEAD_DIM, 0, 1);
    const int slot0 = p.split ? s * 2 : 0;  // first block slot of this WG's item
    const uint32_t k_addr = smem_u32(smem + p.off_kv + c * 2 * KREG_BYTES + slot0 * SLOT_BYTES);
    const uint64_t k_desc = umma_desc_kmajor_sw128(k_addr);
    const uint64_t v_desc = umma_desc_mnmajor_sw128(k_addr + KREG_BYTES, 1024);
    uint32_t u = 0, tcount = 0, pcount = 0;
    for (int g = 2 * blockIdx.x + c; g < p.n_groups; g += g_stride, ++u) {
      const Job j = job_of(p, g, s);
      if (!j.active) continue;  // split mode only: its slots are neither loaded nor awaited
      mbar_wait(&bars->k_full[c][slot0], u & 1);
      mbar_wait(&bars->q_full[w], tcount & 1);
      tc_fence_after();
      if (elect_one()) {  // S_0[128, n_cols] = Q K_0^T
        const int n_cols = min(KVB, p.lp16);
        const uint32_t idesc_s = umma_idesc_f16(128, n_cols, 0, 0);
#pragma unroll
        for (int k = 0; k < HEAD_DIM / 16; ++k)
          if (!(p.debug & 2)) umma_f16_ss(region, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&bars->s_full[w]);
        umma_commit(&bars->k_free[c][slot0]);
      }
      __syncwarp();
      for (int jb = 0; jb < p.n_blk; ++jb) {
        if (jb == 0) mbar_wait(&bars->o_free[w], (tcount & 1) ^ 1);  // the previous tile's O has been read out
        mbar_wait(&bars->v_full[c][slot0 + jb], u & 1);
        if (jb + 1 < p.n_blk) mbar_wait(&bars->k_full[c][slot0 + jb + 1], u & 1);
        mbar_wait(&bars->p_full[w], pcount & 1);
        ++pcount;
        tc_fence_after();
        if (elect_one()) {
          // O[128, 64] (+)= P_jb V_jb : P from TMEM (8 columns per 16 keys), V MN-major (16 key rows = 2048 B per step)
          const int k_steps = min(KVB, p.lp16 - jb * KVB) >> 4;
          const uint64_t vd = v_desc + static_cast<uint64_t>(jb) * (SLOT_BYTES / 16);
          for (int kk = 0; kk < k_steps; ++kk)
            if (!(p.debug & 2)) umma_f16_ts(region + O_COL, region + 8 * kk, vd + 128 * kk, idesc_o, (jb | kk) != 0 ? 1u : 0u);
          umma_commit(&bars->v_free[c][slot0 + jb]);
          if (jb + 1 < p.n_blk) {
            // S_{jb+1} into the same slot: the in-order pipe runs it after PV_jb has consumed P_jb
            const int n_cols = min(KVB, p.lp16 - (jb + 1) * KVB);
            const uint32_t idesc_s = umma_idesc_f16(128, n_cols, 0, 0);
            const uint64_t kd = k_desc + static_cast<uint64_t>(jb + 1) * (SLOT_BYTES / 16);
#pragma unroll
            for (int k = 0; k < HEAD_DIM / 16; ++k)
              if (!(p.debug & 2)) umma_f16_ss(region, q_desc + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
            umma_commit(&bars->s_full[w]);
            umma_commit(&bars->k_free[c][slot0 + jb + 1]);
          } else {
            umma_commit(&bars->pv_done[w]);
          }
        }
        __syncwarp();
      }
      ++tcount;
    }
  } else if (warp < MMA_WARP0) {
    // ---------------------------------------------------------------------------------- softmax WG w
    const int w = warp >> 2, c = w >> 1, s = w & 1;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row of the tile == TMEM lane
    const uint32_t t_row = tmem + w * 128 + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint64_t sc2 = pack_f32x2(sc, sc);
    uint8_t* q_smem = smem + w * Q_BYTES;
    uint32_t scount = 0, tcount = 0;
    if (quarter == 0) {  // this WG's first query tile
      const int g0 = next_active(p, 2 * blockIdx.x + c, g_stride, s);
      if (g0 >= 0 && elect_one()) {
        const Job j = job_of(p, g0, s);
        mbar_arrive_expect_tx(&bars->q_full[w], Q_BYTES);
        tma_load_2d(q_smem, &tmQ, &bars->q_full[w], (j.item % p.heads) * HEAD_DIM, (j.item / p.heads) * p.L + j.tile * 128);
      }
      __syncwarp();
    }
    for (int g = 2 * blockIdx.x + c; g < p.n_groups; g += g_stride) {
      const Job j = job_of(p, g, s);
      if (!j.active) continue;
      const int i = j.tile * 128 + r;  // query index inside the sequence
      const bool warp_live = j.tile * 128 + quarter * 32 < p.L;
      const int jmax = CAUSAL ? min(i, p.L - 1) : p.L - 1;  // last key this row attends to
      float m_ref = -INFINITY;
      uint64_t acc2 = 0;
      TR5(0);
      for (int jb = 0; jb < p.n_blk; ++jb) {
        const int n16 = min(KVB, p.lp16 - jb * KVB) >> 4;
        const int kbase = jb * KVB;
        TR5(1 + jb * 5);
        mbar_wait(&bars->s_full[w], scount & 1);
        ++scount;
        tc_fence_after();
        TR5(2 + jb * 5);
        if (jb == p.n_blk - 1 && quarter == 0) {
          // every S MMA of this tile has retired: the query buffer is free -> fetch the WG's next tile
          const int gn = next_active(p, g + g_stride, g_stride, s);
          if (gn >= 0 && elect_one()) {
            const Job jn = job_of(p, gn, s);
            mbar_arrive_expect_tx(&bars->q_full[w], Q_BYTES);
            tma_load_2d(q_smem, &tmQ, &bars->q_full[w], (jn.item % p.heads) * HEAD_DIM,
                        (jn.item / p.heads) * p.L + jn.tile * 128);
          }
          __syncwarp();
        }
        if (warp_live) {
          const bool full = !CAUSAL && kbase + n16 * 16 <= p.L;
          uint32_t A[16], B[16];
          // ---- pass 1: block maximum
          float mx = -INFINITY;
#pragma unroll
          for (int k = 0; k < 4; k += 2) {
            if (k < n16) {
              tmem_ld_32x16(t_row + k * 16, A);
              if (k + 1 < n16) tmem_ld_32x16(t_row + (k + 1) * 16, B);
              tmem_wait_ld();
              if (full) {
                mx = chunk_max<true>(A, 0, mx);
                if (k + 1 < n16) mx = chunk_max<true>(B, 0, mx);
              } else {
                mx = chunk_max<false>(A, jmax - kbase - k * 16, mx);
                if (k + 1 < n16) mx = chunk_max<false>(B, jmax - kbase - (k + 1) * 16, mx);
              }
            }
          }
          TR5(3 + jb * 5);
          // ---- lazy reference update (S_jb ready => PV_{jb-1} retired: O may be rescaled right away)
          if (jb == 0) {
            m_ref = mx;
          } else {
            const bool grow = (mx - m_ref) * sc > RESCALE_LOG2;  // false for NaN / (-inf) - (-inf)
            if (__any_sync(0xffffffffu, grow)) {
              const float alpha = grow ? ex2_approx((m_ref - mx) * sc) : 1.0f;  // m_ref = -inf -> 0
              if (grow) m_ref = mx;
#pragma unroll 1
              for (int hh = 0; hh < 4; ++hh) {
                tmem_ld_32x16(t_row + O_COL + 16 * hh, A);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) A[e] = __float_as_uint(__uint_as_float(A[e]) * alpha);
                tmem_st_32x16(t_row + O_COL + 16 * hh, A);
              }
              acc2 = fma_f32x2(acc2, pack_f32x2(alpha, alpha), 0);
            }
          }
          // ---- pass 2: p = exp2((s - ref) / 8 * log2 e), fp16 P written over the slot's own S columns
          const float nref = (m_ref == -INFINITY) ? 0.0f : -m_ref * sc;
          const uint64_t nref2 = pack_f32x2(nref, nref);
          uint32_t pk[8];
          tmem_ld_32x16(t_row, A);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < n16) {
              tmem_wait_ld();
              uint32_t(&cur)[16] = (k & 1) ? B : A;
              uint32_t(&nxt)[16] = (k & 1) ? A : B;
              if (k + 1 < n16) tmem_ld_32x16(t_row + (k + 1) * 16, nxt);  // in flight during this chunk's exponentials
              if (p.debug & 1) acc2 = chunk_exp<true, true>(cur, pk, 0, sc2, nref2, acc2);
              else if (full) acc2 = chunk_exp<true>(cur, pk, 0, sc2, nref2, acc2);
              else acc2 = chunk_exp<false>(cur, pk, jmax - kbase - k * 16, sc2, nref2, acc2);
              tmem_st_32x8(t_row + k * 8, pk);  // columns [8k, 8k+8): below every S chunk still to be read
            }
          }
          TR5(4 + jb * 5);
          tmem_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_full[w]);
        TR5(5 + jb * 5);
      }
      // ---- last PV MMA of the tile retired -> O / sum -> fp16 -> out[b, i, h*64 .. h*64+63]
      mbar_wait(&bars->pv_done[w], tcount & 1);
      tc_fence_after();
      TR5(22);
      const float sum = __uint_as_float(static_cast<uint32_t>(acc2)) + __uint_as_float(static_cast<uint32_t>(acc2 >> 32));
      const float inv = __fdividef(1.0f, sum);
      if (warp_live) {
        // fp16 rows go out through this warp's 2 KB staging block, 32 head-dim columns at a time: 32 rows x 64 B,
        // 64B-swizzled (16-byte chunk c of row r at chunk c ^ ((r >> 1) & 3): conflict-free 128-bit stores), then one
        // TMA store through the [B][L][d] map, which clips the rows past the sequence end. (One 128-byte row per
        // thread straight to global memory costs 32 store wavefronts per instruction: 9 us of the launch.)
        uint8_t* stg = smem + p.off_stage + warp * 2048;
        uint8_t* my_row = stg + lane * 64;
        const int b = j.item / p.heads, h = j.item % p.heads;
        uint32_t A[16], B[16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld_32x16(t_row + O_COL + 32 * half, A);
          tmem_ld_32x16(t_row + O_COL + 32 * half + 16, B);
          if (elect_one()) tma_store_wait_read<0>();  // the previous store has drained the staging block
          __syncwarp();
          tmem_wait_ld();
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint32_t(&src)[16] = (cc & 2) ? B : A;
            const int e = (cc & 1) * 8;
            uint4 x;
            x.x = pack_half2(__uint_as_float(src[e + 0]) * inv, __uint_as_float(src[e + 1]) * inv);
            x.y = pack_half2(__uint_as_float(src[e + 2]) * inv, __uint_as_float(src[e + 3]) * inv);
            x.z = pack_half2(__uint_as_float(src[e + 4]) * inv, __uint_as_float(src[e + 5]) * inv);
            x.w = pack_half2(__uint_as_float(src[e + 6]) * inv, __uint_as_float(src[e + 7]) * inv);
            *reinterpret_cast<uint4*>(my_row + ((cc ^ ((lane >> 1) & 3)) << 4)) = x;
          }
          fence_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_store_3d(&tmO, stg, h * HEAD_DIM + 32 * half, j.tile * 128 + quarter * 32, b);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->o_free[w]);
      TR5(23);
      ++tcount;
    }
    if (elect_one()) tma_store_wait_all<0>();  // output written before the CTA (and its staging smem) goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == KV_WARP0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <bool CAUSAL>
int launch_variant5(int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmKV,
                    const CUtensorMap& tmKVt, const CUtensorMap& tmO, const Params5& p) {
  static int configured[kMaxDevices];
  auto kern = attention5_kernel<CAUSAL>;
  PC_CHECK_CUDA(ensure_dynamic_smem(kern, smem_bytes, configured));
  PC_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(THREADS5), smem_bytes, stream, 1, tmQ, tmKV, tmKVt, tmO, p));
  return PC_OK;
}

}  // namespace

// L <= 256 and the two channels' K/V fit next to four query tiles
bool attention5_supports(int L) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("PC_ATTN_IMPL");  // A/B switch: 5 = this kernel, 2 = attention.cu for every shape
    on = (e && atoi(e) == 5) ? 1 : 0;
  }
  return on && L <= 256;
}

int launch_attention5(const __half* qkv, __half* out, int B, int L, int heads, int causal, cudaStream_t stream) {
  const int d = heads * HEAD_DIM;
  Params5 p{};
  p.L = L;
  p.lp16 = (L + 15) / 16 * 16;
  p.heads = heads;
  p.d = d;
  p.items = B * heads;
  p.split = L <= 128 ? 1 : 0;
  p.n_groups = p.split ? (p.items + 1) / 2 : p.items;
  p.n_blk = (p.lp16 + KVB - 1) / KVB;
  p.tail_rows = p.lp16 - KVB * (p.n_blk - 1);
  p.off_kv = 4 * Q_BYTES;
  p.off_stage = p.off_kv + 4 * KREG_BYTES;
  p.off_bars = p.off_stage + 16 * 2048;
  int smem_bytes = p.off_bars + static_cast<int>(sizeof(Bars5));
  PC_REQUIRE(smem_bytes <= 227 * 1024, PC_ERR_ARG, "attention5: L = %d needs %d B of shared memory", L, smem_bytes);
  if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;  // one CTA per SM (it owns all 512 TMEM columns)
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("PC_ATTN5_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  p.debug = dbg;
  static int tracing = -1;
  static long long* trace = nullptr;
  if (tracing < 0) {
    const char* e = getenv("PC_ATTN_TRACE");
    tracing = (e && e[0] == '1') ? 1 : 0;
  }
  if (tracing) {
    if (!trace) PC_CHECK_CUDA(cudaMalloc(&trace, 8 * 32 * sizeof(long long)));
    PC_CHECK_CUDA(cudaMemsetAsync(trace, 0, 8 * 32 * sizeof(long long), stream));
    p.trace = trace;
  }
  CUtensorMap tmQ, tmKV, tmKVt, tmO;
  const uint64_t rows = static_cast<uint64_t>(B) * L;
  PC_TRY(make_tmap_f16_2d(&tmQ, qkv, 3 * d, rows, static_cast<uint64_t>(3 * d) * 2, 64, 128));
  PC_TRY(make_tmap_f16_2d(&tmKV, qkv, 3 * d, rows, static_cast<uint64_t>(3 * d) * 2, 64, KVB));
  PC_TRY(make_tmap_f16_2d(&tmKVt, qkv, 3 * d, rows, static_cast<uint64_t>(3 * d) * 2, 64, p.tail_rows));
  PC_TRY(make_tmap_f16_3d(&tmO, out, d, L, B, static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(L) * d * 2, 32, 32));
  const int sms = device_sm_count();
  const int want = (p.n_groups + 1) / 2;
  const int grid = want < sms ? want : sms;
  PC_TRY(causal ? launch_variant5<true>(grid, smem_bytes, stream, tmQ, tmKV, tmKVt, tmO, p)
                : launch_variant5<false>(grid, smem_bytes, stream, tmQ, tmKV, tmKVt, tmO, p));
  if (tracing) {
    static int printed = 0;
    static long long h[8 * 32];
    PC_CHECK_CUDA(cudaStreamSynchronize(stream));
    PC_CHECK_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    if (printed++ == 3) {
      const long long t0 = h[0];
      fprintf(stderr, "[attn5 trace] items=%d L=%d (cycles since the first tile's start; CTA 0, WG 0, warp 0)\n", p.items, L);
      for (int t = 0; t < 8 && h[t * 32]; ++t) {
        const long long* r = h + t * 32;
        fprintf(stderr, "tile %d: start %7lld  pv_done %7lld  stored %7lld\n", t, r[0] - t0, r[22] - t0, r[23] - t0);
        for (int jb = 0; jb < 4 && r[1 + jb * 5]; ++jb) {
          const long long* b = r + 1 + jb * 5;
          fprintf(stderr, "   blk %d: at %7lld | s_full +%5lld  pass1 +%5lld  pass2 +%5lld  st/arrive +%5lld\n", jb, b[0] - t0,
                  b[1] - b[0], b[2] - b[1], b[3] - b[2], b[4] - b[3]);
        }
      }
    }
  }
  return PC_OK;
}

}  // namespace pc
