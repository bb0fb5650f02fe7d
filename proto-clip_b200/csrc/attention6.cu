// Multi-head self-attention core for sequences of at most 208 tokens (ViT-B/32, ViT-B/16, the text tower, the
// ModifiedResNet attention pools): WHOLE-ROW scores in TMEM, exact (single-pass) softmax.
// Same contract as attention.cu (reference clip/model.py:173,183-185; causal flag = the text mask of
// clip/model.py:326-332).
//
// Why: the round-1 kernel (attention5.cu) streamed 64-key blocks through an online softmax. Per block it paid one
// S commit -> mbarrier -> tcgen05.ld hop, one P -> mbarrier -> PV hop and a possible O rescale, and ncu showed it
// bound by instruction issue (ALU pipe 40 %, MUFU 36 %, tensor 17 %), not by a pipe it could saturate. For L <= 208
// the whole score row fits in TMEM, so a query tile costs TWO hops in total and a minimal instruction stream:
//   S[128, lp16] = Q K^T          two UMMA batches (N = 128 and N = lp16 - 128), one commit
//   row maximum over all keys     tcgen05.ld + 3-input max only, then p = exp2((s - max) * scale) written back
//                                 IN PLACE as packed fp16: per pair of scores one FFMA2, two MUFU.EX2 (or the FMA-pipe
//                                 polynomial), one FADD2 (row sum) and one F2FP -- no running maximum, no rescale of O
//   O[128, 64] = P V              two UMMA batches with P as the TMEM A operand; the first starts while the
//                                 exponentials of the other part are still running
//
// TMEM: 256 columns per query tile. S_a (keys [0,128)) at [0,128), S_b at [128, 128 + nb), nb <= 80; the O accumulator
// at [192,256). Every softmax thread writes its fp16 P IN PLACE over score columns it has already read itself (so
// there is no hazard between threads): P is not contiguous, each UMMA k-step of P V addresses its own 8 columns.
// Part b (keys >= 128) is exponentiated FIRST: once it is done its last score chunk [192,208) is dead and P_b V_b
// may start accumulating O while part a's exponentials are still running; S_a of the NEXT tile does not overlap O,
// so it is issued right behind the tile's last P V MMA without waiting for the epilogue.
//
// Persistent kernel, one CTA per SM, 20 warps:
//   warps 0-7 / 8-15  softmax warpgroup 0 / 1 (8 warps = one 128-row query tile): the two tiles of one (image, head)
//                    item (128 < L <= 208) or two different items (L <= 128, "split"). TWO threads per query row
//                    (row == TMEM lane; warps q and q + 4 of a group share lane quadrant q): thread 0 owns key chunks
//                    [0, h0), thread 1 owns [h0, nch), h0 = ceil(nch / 2). Each takes the maximum of its own chunks,
//                    the two meet through shared memory (one 64-thread named barrier), each exponentiates its own
//                    chunks, the partial row sums meet the same way, and in the epilogue each thread scales and
//                    stores 32 of the row's 64 output columns. Four softmax warps per scheduler keep the MUFU pipe fed
//                    (a single warp reaches 75 % of it, profiles/r01_ubench2.log); NP of every 8 score pairs take
//                    a cubic polynomial on the FMA pipe instead. The two warpgroups can hand the MUFU pipe to each
//                    other through a pair of named barriers (ping-pong, PC_ATTN6_PINGPONG).
//   warp 16 / 17     MMA issuer of WG 0 / 1 (whole warp in the loop, one elected lane issues)
//   warp 18          TMA producer: a STAGE holds everything an item needs (Q tile 0, Q tile 1, K, V: 96 KB); two
//                    stages, so the next item's operands land while the current one is computed
//   warp 19          idle (warpgroup padding); in extra-key mode (L = 257) it computes query row 256 of every item on the
//                    CUDA cores from the K / V tiles in shared memory
#include <stdlib.h>

#include "attn_common.cuh"
#include "attn_wholerow.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace pc {
namespace {

constexpr int HEAD_DIM = 64;
constexpr int MMA_WARP0 = 16;
constexpr int TMA_WARP = 18;
constexpr int THREADS6 = 20 * 32;
constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 fp16, 128B-swizzled
constexpr int OFF_Q = 0;               // Q tile of WG 0, then of WG 1
constexpr int OFF_K = 2 * TILE_BYTES;  // 256 key rows (split: 128 per WG)
constexpr int OFF_V = 4 * TILE_BYTES;
constexpr int STAGE_BYTES = 6 * TILE_BYTES;   // 96 KB
constexpr int OFF_OUT = 2 * STAGE_BYTES;      // 16 x 2 KB output staging blocks (one per softmax warp: 32 rows x 64 B)
constexpr int OFF_XCH = OFF_OUT + 16 * 2048;  // row maximum / partial row sum exchange [tile][thread of the row][row] fp32
constexpr int OFF_XROW = OFF_XCH + 2 * 2 * 128 * 4;  // extra-key mode, per stage: k_256 | v_256 | q_256 (128 B each, linear)
constexpr int XROW_BYTES = 3 * 128;
constexpr int OFF_BARS = OFF_XROW + 2 * XROW_BYTES;
constexpr int HELPER_WARP = 19;
constexpr int O_COL = 192;   // O accumulator inside a tile's 256-column TMEM region

struct Params6 {
  int L, lp16, heads, d, items;  // L = rows of a sequence in memory
  int Lm;        // query rows / keys this kernel runs through the tensor cores (= L, or 256 in extra-key mode)
  const __half* qkv;            // extra-key mode: element pointer + pitches of the [plane][row][64] view
  long long row_pitch, plane_pitch;
  __half* out;                  // extra-key mode: [B][L][d], row 256 of every sequence is written by the helper warp
  int split;     // 1: L <= 128, the two WGs take different items
  int n_groups;  // split: ceil(items / 2), else items
  int na;        // keys of part a = min(lp16, 128)
  int nb;        // keys of part b = lp16 - na
  int pingpong;  // 1: the WGs alternate on the MUFU pipe (named barriers 1 / 2)
  int debug;     // bring-up only (env PC_ATTN6_DEBUG): 2 = no MMA issue (wrong results: timing A/B only)
  long long* trace;  // bring-up only (env PC_ATTN_TRACE=1): [tile][WG][16] clock64 samples of CTA 0, warp quarter 0, thread 0 of the row
};

#define TR6(slot)                                                                                              \
  do {                                                                                                         \
    if (p.trace != nullptr && blockIdx.x == 0 && quarter == 0 && hf == 0 && lane == 0 && tcount < 8)           \
      p.trace[(tcount * 2 + w) * 16 + (slot)] = clock64();                                                     \
  } while (0)

struct Bars6 {
  uint64_t qk_full[2];     // per stage: Q tiles + K landed
  uint64_t v_full[2];      // per stage: V landed
  uint64_t qk_free[2];     // per stage: both WGs' S MMAs retired -> Q and K may be overwritten (2 arrivals)
  uint64_t v_free[2];      // per stage: both WGs' last PV MMA retired -> V may be overwritten (2 arrivals)
  uint64_t s_full[2];      // per WG: S of the tile in TMEM
  uint64_t pa_full[2];     // per WG: P part a in TMEM (8 warp arrivals)
  uint64_t pb_full[2];     // per WG: P part b in TMEM (4 warp arrivals: it belongs to the rows' second threads)
  uint64_t pv_done[2];     // per WG: last PV MMA of the tile retired
  uint64_t o_free[2];      // per WG: O read out (8 warp arrivals)
  uint32_t tmem_base;
};

struct Job6 {
  bool active;
  int item;  // b * heads + h
  int tile;
};
__device__ __forceinline__ Job6 job_of(const Params6& p, int g, int w) {
  Job6 j;
  if (p.split) {
    j.item = 2 * g + w;
    j.tile = 0;
    j.active = j.item < p.items;
  } else {
    j.item = g;
    j.tile = w;
    j.active = true;
  }
  return j;
}

using namespace wr;


// XKEY (L = 257 = 2 x 128 + 1, ViT-L/14): the tensor cores see 256 queries x 256 keys. The producer also drops rows
// k_256, v_256 and q_256 of the item into shared memory (three 128-byte bulk copies). Key 256 is folded in by the softmax
// threads themselves: its score is a 64-long dot product per row (half per thread of the row, q from the Q tile in
// shared memory), it joins the row maximum and the row sum, and its p * v row is added to O in the epilogue. Query row
// 256 is computed by the spare warp on the CUDA cores from the K / V tiles in shared memory (257 dot products, one
// softmax, 257 axpys per item): a separate pass over K and V for that one row cost 57 us of 175 at 192 images.
template <bool CAUSAL, int NP, bool XKEY>
__global__ void __launch_bounds__(THREADS6, 1)
attention6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmT,
                  const __grid_constant__ CUtensorMap tmO, const Params6 p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Bars6* bars = reinterpret_cast<Bars6*>(smem + OFF_BARS);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;

  if (warp == TMA_WARP) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmT);
      tma_prefetch_desc(&tmO);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->qk_full[i], 1);
        mbar_init(&bars->v_full[i], 1);
        mbar_init(&bars->qk_free[i], XKEY ? 2 + 16 + 1 : 2);  // XKEY: the softmax warps and the helper read the stage too
        mbar_init(&bars->v_free[i], XKEY ? 2 + 16 + 1 : 2);
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
  griddep_wait();  // qkv is the previous kernel's output

  const int g_stride = gridDim.x;
  const int nch = p.lp16 >> 4;    // 16-key score chunks per row
  const int h0 = (nch + 1) >> 1;  // chunks [0, h0) belong to thread 0 of a row, [h0, nch) to thread 1 (h0 <= 7)
  if (warp == TMA_WARP) {
    // ---------------------------------------------------------------------------------- producer
    // Q and K of a stage are released as soon as the item's S MMAs have retired, V only after its last P V MMA: the
    // Q / K of item i + 2 are on their way while item i is still in its softmax (the kernel is HBM-fed at large batches).
    // Boxes of lp16 (split) / nb (second tile) rows: the rows of a 128-row tile past the sequence stay whatever the
    // buffer held (finite or not, they only feed score rows >= L, which are never stored).
    uint32_t u = 0;
    for (int g = blockIdx.x; g < p.n_groups; g += g_stride, ++u) {
      const int st = u & 1;
      const uint32_t free_par = ((u >> 1) & 1) ^ 1;
      uint8_t* stage = smem + st * STAGE_BYTES;
      const int n_act = (p.split && !job_of(p, g, 1).active) ? 1 : 2;
      mbar_wait(&bars->qk_free[st], free_par);
      if (elect_one()) {
        uint64_t* qk = &bars->qk_full[st];
        if (p.split) {
          mbar_arrive_expect_tx(qk, static_cast<uint32_t>(n_act) * (2 * p.na * 128));
          for (int s = 0; s < n_act; ++s) {
            const int item = 2 * g + s;
            const int hd = item % p.heads, r0 = (item / p.heads) * p.L;
            tma_load_3d(stage + OFF_Q + s * TILE_BYTES, &tmT, qk, 0, r0, hd);
            tma_load_3d(stage + OFF_K + s * TILE_BYTES, &tmT, qk, 0, r0, p.heads + hd);
          }
        } else {
          const int hd = g % p.heads, r0 = (g / p.heads) * p.L;
          mbar_arrive_expect_tx(qk, 2 * TILE_BYTES + 2 * p.nb * 128 + (XKEY ? 256 : 0));
          if (XKEY) {
            uint8_t* xr = smem + OFF_XROW + st * XROW_BYTES;
            const __half* row = p.qkv + static_cast<long long>(r0 + p.Lm) * p.row_pitch;
            bulk_load(xr, row + (p.heads + hd) * p.plane_pitch, 128, qk);
            bulk_load(xr + 256, row + hd * p.plane_pitch, 128, qk);
          }
          tma_load_3d(stage + OFF_K, &tmQ, qk, 0, r0, p.heads + hd);
          tma_load_3d(stage + OFF_Q, &tmQ, qk, 0, r0, hd);
          tma_load_3d(stage + OFF_K + TILE_BYTES, &tmT, qk, 0, r0 + 128, p.heads + hd);
          tma_load_3d(stage + OFF_Q + TILE_BYTES, &tmT, qk, 0, r0 + 128, hd);
        }
      }
      __syncwarp();
      mbar_wait(&bars->v_free[st], free_par);
      if (elect_one()) {
        uint64_t* vf = &bars->v_full[st];
        if (p.split) {
          mbar_arrive_expect_tx(vf, static_cast<uint32_t>(n_act) * (p.na * 128));
          for (int s = 0; s < n_act; ++s) {
            const int item = 2 * g + s;
            tma_load_3d(stage + OFF_V + s * TILE_BYTES, &tmT, vf, 0, (item / p.heads) * p.L, 2 * p.heads + item % p.heads);
          }
        } else {
          const int hd = g % p.heads, r0 = (g / p.heads) * p.L;
          mbar_arrive_expect_tx(vf, TILE_BYTES + p.nb * 128 + (XKEY ? 128 : 0));
          if (XKEY)
            bulk_load(smem + OFF_XROW + st * XROW_BYTES + 128,
                      p.qkv + static_cast<long long>(r0 + p.Lm) * p.row_pitch + (2 * p.heads + hd) * p.plane_pitch, 128, vf);
          tma_load_3d(stage + OFF_V, &tmQ, vf, 0, r0, 2 * p.heads + hd);
          tma_load_3d(stage + OFF_V + TILE_BYTES, &tmT, vf, 0, r0 + 128, 2 * p.heads + hd);
        }
      }
      __syncwarp();
    }
  } else if (warp == MMA_WARP0 || warp == MMA_WARP0 + 1) {
    // ---------------------------------------------------------------------------------- MMA issuer of WG w
    const int w = warp - MMA_WARP0;
    const uint32_t region = tmem + w * 256;
    const uint32_t idesc_o = umma_idesc_f16(128, HEAD_DIM, 0, 1);
    const uint32_t idesc_sa = umma_idesc_f16(128, p.na, 0, 0);
    const uint32_t idesc_sb = umma_idesc_f16(128, p.nb > 0 ? p.nb : 16, 0, 0);
    const int kv_off = p.split ? w * TILE_BYTES : 0;
    const int ka_steps = p.na >> 4;
    uint32_t u = 0, tcount = 0;
    for (int g = blockIdx.x; g < p.n_groups; g += g_stride, ++u) {
      const int st = u & 1;
      const uint32_t full_par = (u >> 1) & 1;
      const Job6 j = job_of(p, g, w);
      if (!j.active) {  // split mode, odd item count: nothing was loaded for this WG; release its share of the stage
        if (elect_one()) {
          mbar_arrive(&bars->qk_free[st]);
          mbar_arrive(&bars->v_free[st]);
        }
        __syncwarp();
        continue;
      }
      const uint32_t sbase = smem_u32(smem + st * STAGE_BYTES);
      const uint64_t q_desc = umma_desc_kmajor_sw128(sbase + OFF_Q + w * TILE_BYTES);
      const uint64_t k_desc = umma_desc_kmajor_sw128(sbase + OFF_K + kv_off);
      const uint64_t v_desc = umma_desc_mnmajor_sw128(sbase + OFF_V + kv_off, 1024);
      mbar_wait(&bars->qk_full[st], full_par);
      tc_fence_after();
      // S_a overlaps neither O nor anything the previous tile still reads (its P was consumed by the MMAs issued
      // before this one: the tensor pipe runs them in order)
      if (elect_one()) {
        if (!(p.debug & 2)) {
#pragma unroll
          for (int k = 0; k < HEAD_DIM / 16; ++k)
            umma_f16_ss(region, q_desc + 2 * k, k_desc + 2 * k, idesc_sa, k != 0 ? 1u : 0u);
        }
      }
      __syncwarp();
      // S_b's last chunk and the first P V MMA overwrite O: the previous tile's O must have been read out. (s_full is
      // committed behind this wait in split mode too: it orders every softmax thread's read of its partner's
      // exchange slot before the partner's next write.)
      mbar_wait(&bars->o_free[w], (tcount & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        if (nch > 8 && !(p.debug & 2)) {
#pragma unroll
          for (int k = 0; k < HEAD_DIM / 16; ++k)
            umma_f16_ss(region + SB_COL, q_desc + 2 * k, k_desc + (TILE_BYTES >> 4) + 2 * k, idesc_sb, k != 0 ? 1u : 0u);
        }
        umma_commit(&bars->s_full[w]);
        umma_commit(&bars->qk_free[st]);  // this WG's reads of the stage's Q and K have retired
      }
      __syncwarp();
      mbar_wait(&bars->v_full[st], full_par);
      if (nch > 8) {
        mbar_wait(&bars->pb_full[w], tcount & 1);
        tc_fence_after();
        if (elect_one()) {
          // O[128, 64] = P_b V_b : P from TMEM (8 columns per 16 keys), V MN-major (16 key rows = 2048 B per step)
          if (!(p.debug & 2))
            for (int k = 8; k < nch; ++k)
              umma_f16_ts(region + O_COL, region + p_col(k, h0), v_desc + 128 * k, idesc_o, k != 8 ? 1u : 0u);
        }
        __syncwarp();
      }
      mbar_wait(&bars->pa_full[w], tcount & 1);
      tc_fence_after();
      if (elect_one()) {
        if (!(p.debug & 2))
          for (int k = 0; k < ka_steps; ++k)
            umma_f16_ts(region + O_COL, region + p_col(k, h0), v_desc + 128 * k, idesc_o, (nch > 8 || k != 0) ? 1u : 0u);
        umma_commit(&bars->pv_done[w]);
        umma_commit(&bars->v_free[st]);
      }
      __syncwarp();
      ++tcount;
    }
  } else if (warp < MMA_WARP0) {
    // ---------------------------------------------------------------------------------- softmax WG w
    const int w = warp >> 3;
    const int hf = (warp >> 2) & 1;  // which of the two threads of a row
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row of the tile == TMEM lane
    const uint32_t t_row = tmem + w * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint64_t sc2 = pack_f32x2(sc, sc);
    // this thread's chunks [c0, c1); [c0, ca) belong to part a, [cb, c1) to part b
    const int c0 = hf ? h0 : 0, c1 = hf ? nch : h0;
    const int ca = min(c1, 8), cb = max(c0, 8);
    const int pair_bar = 3 + w * 4 + quarter;  // named barrier of the row's two threads (two warps)
    uint8_t* stg = smem + OFF_OUT + warp * 2048;
    float* my_x = reinterpret_cast<float*>(smem + OFF_XCH) + (w * 2 + hf) * 128 + r;
    const float* other_x = reinterpret_cast<float*>(smem + OFF_XCH) + (w * 2 + (hf ^ 1)) * 128 + r;
    uint32_t tcount = 0;
    const int n_iter = (p.n_groups - static_cast<int>(blockIdx.x) + g_stride - 1) / g_stride;
    if (p.pingpong && w == 1) asm volatile("bar.arrive 1, 512;" ::: "memory");  // WG 0 goes first
    int it = 0;
    for (int g = blockIdx.x; g < p.n_groups; g += g_stride, ++it) {
      const Job6 j = job_of(p, g, w);
      const bool last_iter = it == n_iter - 1;
      if (!j.active) {  // keep the hand-over protocol balanced
        if (p.pingpong) {
          if (w == 0) {
            asm volatile("bar.sync 1, 512;" ::: "memory");
            asm volatile("bar.arrive 2, 512;" ::: "memory");
          } else {
            asm volatile("bar.sync 2, 512;" ::: "memory");
            if (!last_iter) asm volatile("bar.arrive 1, 512;" ::: "memory");
          }
        }
        continue;
      }
      const int i = j.tile * 128 + r;  // query index inside the sequence
      const bool warp_live = j.tile * 128 + quarter * 32 < p.Lm;
      const int jmax = CAUSAL ? min(i, p.Lm - 1) : p.Lm - 1;  // last key this row attends to (through the MMAs)
      TR6(0);
      float s_x = 0.0f;  // XKEY: q_i . k_256 (this thread's 32 of the 64 dimensions until the exchange below)
      const int xst = it & 1;
      if (XKEY) {
        mbar_wait(&bars->qk_full[xst], (it >> 1) & 1);
        const uint8_t* qrow = smem + xst * STAGE_BYTES + OFF_Q + w * TILE_BYTES + r * 128;
        const uint8_t* xk = smem + OFF_XROW + xst * XROW_BYTES + 64 * hf;
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
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->qk_free[xst]);  // this warp's reads of the stage's Q tile and k_256 are done
      }
      mbar_wait(&bars->s_full[w], tcount & 1);
      tc_fence_after();
      TR6(1);
      if (XKEY) {  // the two halves of the dot product meet (slot writes are ordered behind s_full, see the MMA warp)
        *my_x = s_x;
        pair_bar_sync(pair_bar);
        s_x = hf ? *other_x + s_x : s_x + *other_x;  // same order in both threads
        pair_bar_sync(pair_bar);
      }
      // ---- row maximum: own chunks, then the partner's through shared memory
      float mx = -INFINITY;
      if (warp_live) mx = row_max<CAUSAL>(t_row, c0, c1, p.Lm, jmax);
      if (XKEY) mx = fmaxf(mx, s_x);
      *my_x = mx;
      pair_bar_sync(pair_bar);
      mx = fmaxf(mx, *other_x);
      pair_bar_sync(pair_bar);  // both maxima read: the slots may take the partial sums
      const float nref = (mx == -INFINITY) ? 0.0f : -mx * sc;
      const uint64_t nref2 = pack_f32x2(nref, nref);
      TR6(2);
      if (p.pingpong) {
        if (w == 0) asm volatile("bar.sync 1, 512;" ::: "memory");
        else asm volatile("bar.sync 2, 512;" ::: "memory");
      }
      TR6(3);
      // ---- exponentials: part b first (it frees the columns O overlaps), then part a
      uint64_t acc_a = 0, acc_b = 0;
      if (hf == 1 && nch > 8) {
        if (warp_live) {
          exp_chunks<CAUSAL, NP>(t_row, cb, c1, p_col(cb, h0), p.Lm, jmax, sc2, nref2, acc_a, acc_b);
          tmem_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->pb_full[w]);
      }
      TR6(4);
      if (warp_live) {
        exp_chunks<CAUSAL, NP>(t_row, c0, ca, p_col(c0, h0), p.Lm, jmax, sc2, nref2, acc_a, acc_b);
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->pa_full[w]);
      if (p.pingpong) {
        if (w == 0) asm volatile("bar.arrive 2, 512;" ::: "memory");
        else if (!last_iter) asm volatile("bar.arrive 1, 512;" ::: "memory");
      }
      TR6(5);
      // ---- the row sum is the sum of its two threads' partial sums
      const float part = (lo_f(acc_a) + hi_f(acc_a)) + (lo_f(acc_b) + hi_f(acc_b));
      *my_x = part;
      pair_bar_sync(pair_bar);
      float sum = hf ? *other_x + part : part + *other_x;  // same order in both threads
      float p_x = 0.0f;
      uint4 vx[4];
      if (XKEY) {
        const float e = ex2_approx(fmaf(s_x, sc, nref));
        sum += e;
        p_x = __half2float(__float2half_rn(e));  // P is fp16 for the tensor cores: the same rounding for this key
      }
      // ---- last PV MMA of the tile retired -> O / sum -> fp16 -> out[b, i, h*64 + 32*hf .. +31]
      mbar_wait(&bars->pv_done[w], tcount & 1);
      tc_fence_after();
      TR6(6);
      if (XKEY) {
        mbar_wait(&bars->v_full[xst], (it >> 1) & 1);  // long complete (the P V MMAs waited for it): orders the reads below
        const uint8_t* xv = smem + OFF_XROW + xst * XROW_BYTES + 128 + 64 * hf;
#pragma unroll
        for (int c = 0; c < 4; ++c) vx[c] = *reinterpret_cast<const uint4*>(xv + 16 * c);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->v_free[xst]);
      }
      if (!warp_live) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->o_free[w]);
      }
      if (warp_live) {
        uint32_t O2[32];
        tmem_ld_32x32(t_row + O_COL + 32 * hf, O2);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->o_free[w]);  // O may be overwritten by the next tile
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
        TR6(7);
        // fp16 rows into this warp's 2 KB staging block: 32 rows x 64 B, 64B-swizzled (16-byte chunk c of row r at
        // chunk c ^ ((r >> 1) & 3): conflict-free 128-bit stores), then one TMA store through the [B][L][d] map,
        // which clips the rows past the sequence end.
        if (elect_one()) tma_store_wait_read<0>();  // the previous tile's store has drained this block
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
          const int b = j.item / p.heads, h = j.item % p.heads;
          tma_store_3d(&tmO, stg, h * HEAD_DIM + 32 * hf, j.tile * 128 + quarter * 32, b);
          tma_store_commit();
        }
      }
      TR6(8);
      ++tcount;
    }
    if (elect_one()) tma_store_wait_all<0>();  // output written before the CTA (and its staging smem) goes away
  }
  if (XKEY && warp == HELPER_WARP) {
    // ---------------------------------------------------------------------------------- query row 256 of every item
    // One warp, warp-level tensor-core instructions (mma.sync m16n8k16) in the transposed form, so that the operand that
    // carries a single useful vector is the 8-wide B side: S^T = K q^T (16 keys per instruction, A through ldmatrix from
    // the swizzled K tiles), key 256 by hand, one softmax across the 8 lanes that hold column 0, O^T = V^T P^T (A through
    // ldmatrix.trans from the V tiles, P moved into B-fragment position by shuffles). 128 mma + 128 ldmatrix per item.
    const float sc = 0.125f * 1.4426950408889634f;
    const int g4 = lane >> 2, t4 = lane & 3;
    int it = 0;
    for (int g = blockIdx.x; g < p.n_groups; g += g_stride, ++it) {
      const int st = it & 1;
      const uint32_t par = (it >> 1) & 1;
      const uint8_t* stage = smem + st * STAGE_BYTES;
      const uint8_t* xr = smem + OFF_XROW + st * XROW_BYTES;
      mbar_wait(&bars->qk_full[st], par);
      if (p.debug & 4) {  // bring-up (wrong row 256): barrier traffic only
        mbar_wait(&bars->v_full[st], par);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->qk_free[st]);
          mbar_arrive(&bars->v_free[st]);
        }
        continue;
      }
      // ---- S^T = K q^T: A fragments = 16 keys x 16 dims of the K tile (ldmatrix), B = q in column 0 (lanes 0..3)
      uint32_t qb0[4], qb1[4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        qb0[ks] = g4 == 0 ? *reinterpret_cast<const uint32_t*>(xr + 256 + (16 * ks + 2 * t4) * 2) : 0u;
        qb1[ks] = g4 == 0 ? *reinterpret_cast<const uint32_t*>(xr + 256 + (16 * ks + 8 + 2 * t4) * 2) : 0u;
      }
      // ldmatrix row of this lane: matrix i = lane >> 3 -> keys + 8 (i & 1), dim chunk + (i >> 1); row lane & 7
      const int lk = (lane & 7) + 8 * ((lane >> 3) & 1), lc = lane >> 4;
      float sv[32];  // lanes with t4 == 0: scores of keys 16 mt + g4 (sv[2 mt]) and 16 mt + 8 + g4 (sv[2 mt + 1])
#pragma unroll
      for (int mt = 0; mt < 16; ++mt) {
        const int key = 16 * mt + lk;
        const uint32_t krow = smem_u32(stage + OFF_K + (key >> 7) * TILE_BYTES + (key & 127) * 128);
        float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4(a0, a1, a2, a3, krow + (((2 * ks + lc) ^ (key & 7)) << 4));
          mma_m16n8k16_f16(c0, c1, c2, c3, a0, a1, a2, a3, qb0[ks], qb1[ks]);
        }
        sv[2 * mt] = c0;
        sv[2 * mt + 1] = c2;
      }
      // key 256: lanes 0..3 hold q; the four partial dot products meet, lane 0 broadcasts
      float sx = 0.0f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float2 q0 = __half22float2(*reinterpret_cast<const __half2*>(&qb0[ks]));
        const float2 q1 = __half22float2(*reinterpret_cast<const __half2*>(&qb1[ks]));
        const float2 k0 = __half22float2(*reinterpret_cast<const __half2*>(xr + (16 * ks + 2 * t4) * 2));
        const float2 k1 = __half22float2(*reinterpret_cast<const __half2*>(xr + (16 * ks + 8 + 2 * t4) * 2));
        sx = fmaf(q0.x, k0.x, fmaf(q0.y, k0.y, fmaf(q1.x, k1.x, fmaf(q1.y, k1.y, sx))));
      }
      sx += __shfl_xor_sync(0xffffffffu, sx, 1);
      sx += __shfl_xor_sync(0xffffffffu, sx, 2);
      sx = __shfl_sync(0xffffffffu, sx, 0);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->qk_free[st]);
      // ---- softmax over the 8 lanes with t4 == 0 (the other lanes hold columns 1..7 of S^T: zeros, harmless)
      float mx = sx;
#pragma unroll
      for (int e = 0; e < 32; ++e) mx = fmaxf(mx, sv[e]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
      float sum = 0.0f;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        sv[e] = ex2_approx((sv[e] - mx) * sc);
        sum += sv[e];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      sum += __shfl_xor_sync(0xffffffffu, sum, 8);
      sum += __shfl_xor_sync(0xffffffffu, sum, 16);
      const float ex = ex2_approx((sx - mx) * sc);
      sum += ex;
      const float px = __half2float(__float2half_rn(ex));
      // ---- O^T = V^T P^T: A fragments = 16 dims x 16 keys of the V tile (ldmatrix.trans), B = P in column 0
      mbar_wait(&bars->v_full[st], par);
      float o[4][4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.0f;
      // ldmatrix.trans row of this lane: matrix i -> dim chunk + (i & 1), keys + 8 (i >> 1); row (= key) lane & 7
      const int vk = (lane & 7) + 8 * (lane >> 4), vc = (lane >> 3) & 1;
#pragma unroll
      for (int kt = 0; kt < 16; ++kt) {
        // P of keys 16 kt + 2 t4 (+1) and + 8: they live in lanes 4 g (t4 == 0) with g = key % 8
        const float p00 = __shfl_sync(0xffffffffu, sv[2 * kt], 8 * t4), p01 = __shfl_sync(0xffffffffu, sv[2 * kt], 8 * t4 + 4);
        const float p10 = __shfl_sync(0xffffffffu, sv[2 * kt + 1], 8 * t4),
                    p11 = __shfl_sync(0xffffffffu, sv[2 * kt + 1], 8 * t4 + 4);
        const uint32_t b0 = g4 == 0 ? pack_half2(p00, p01) : 0u, b1 = g4 == 0 ? pack_half2(p10, p11) : 0u;
        const int key = 16 * kt + vk;
        const uint32_t vrow = smem_u32(stage + OFF_V + (key >> 7) * TILE_BYTES + (key & 127) * 128);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4_trans(a0, a1, a2, a3, vrow + (((2 * mt + vc) ^ (key & 7)) << 4));
          mma_m16n8k16_f16(o[mt][0], o[mt][1], o[mt][2], o[mt][3], a0, a1, a2, a3, b0, b1);
        }
      }
      const float inv = __fdividef(1.0f, sum);
      const long long b = g / p.heads, h = g % p.heads;
      __half* orow = p.out + (b * p.L + p.Lm) * p.d + h * HEAD_DIM;
      if (t4 == 0) {  // column 0 of O^T: dims 16 mt + g4 (c0) and 16 mt + 8 + g4 (c2)
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const int d0 = 16 * mt + g4, d1 = d0 + 8;
          const float v0 = __half2float(*reinterpret_cast<const __half*>(xr + 128 + 2 * d0));
          const float v1 = __half2float(*reinterpret_cast<const __half*>(xr + 128 + 2 * d1));
          orow[d0] = __float2half_rn(fmaf(px, v0, o[mt][0]) * inv);
          orow[d1] = __float2half_rn(fmaf(px, v1, o[mt][2]) * inv);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->v_free[st]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <bool CAUSAL, int NP, bool XKEY = false>
int launch_variant6(int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmT,
                    const CUtensorMap& tmO, const Params6& p) {
  static int configured[kMaxDevices];
  auto kern = attention6_kernel<CAUSAL, NP, XKEY>;
  PC_CHECK_CUDA(ensure_dynamic_smem(kern, smem_bytes, configured));
  PC_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(THREADS6), smem_bytes, stream, 1, tmQ, tmT, tmO, p));
  return PC_OK;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

// L <= 208: the whole score row of a 128-query tile (<= 208 fp32 columns) and its O accumulator (64) fit in half of
// the SM's tensor memory with S_a clear of O
bool attention6_supports(int L) {
  static int impl = -1, max_l = -1;
  if (impl < 0) impl = env_int("PC_ATTN_IMPL", 6);  // A/B switch: 5 = round-1 kernel (attention5.cu), 2 = attention.cu
  if (max_l < 0) max_l = env_int("PC_ATTN6_MAX_L", 256);
  return impl == 6 && L <= max_l && L <= 256;
}

// qkv: packed [B*L, 3d] (nn.MultiheadAttention in-proj order) or planar [3 * heads][B*L][64] (GemmArgs::c_planar).
// Either way the kernel sees a 3-D tensor [plane = 3 * heads][row = B*L][64]: only the strides differ.
// L = 257 (ViT-L/14), no mask: 256 x 256 on the tensor cores + key 256 in the softmax threads (XKEY); the caller adds
// query row 256 (attention_tail_rows_kernel). PC_ATTN6_XKEY=0 keeps that length on the streaming kernel (A/B).
bool attention6_supports_xkey(int L, int causal) {
  static int on = -1;
  if (on < 0) on = env_int("PC_ATTN6_XKEY", 1);
  return on && attention6_supports(256) && L == 257 && !causal;
}

int launch_attention6(const __half* qkv, int qkv_planar, __half* out, int B, int L, int heads, int causal,
                      cudaStream_t stream) {
  const int d = heads * HEAD_DIM;
  const bool xkey = L == 257;
  PC_REQUIRE(!xkey || (!causal && !qkv_planar), PC_ERR_ARG, "attention6: the extra-key mode takes packed, unmasked qkv");
  Params6 p{};
  p.L = L;
  p.Lm = xkey ? 256 : L;
  p.qkv = qkv;
  p.out = out;
  p.row_pitch = 3 * d;
  p.plane_pitch = HEAD_DIM;
  p.lp16 = (p.Lm + 15) / 16 * 16;
  p.heads = heads;
  p.d = d;
  p.items = B * heads;
  p.split = p.Lm <= 128 ? 1 : 0;
  p.n_groups = p.split ? (p.items + 1) / 2 : p.items;
  p.na = p.lp16 < 128 ? p.lp16 : 128;
  p.nb = p.lp16 - p.na;
  static int pingpong = -1, dbg = -1, tracing = -1, npoly = -1;
  if (pingpong < 0) pingpong = env_int("PC_ATTN6_PINGPONG", 1);
  if (dbg < 0) dbg = env_int("PC_ATTN6_DEBUG", 0);
  if (tracing < 0) tracing = env_int("PC_ATTN_TRACE", 0);
  if (npoly < 0) npoly = env_int("PC_ATTN6_POLY", 0);  // pairs (of 8 per 16-key chunk) on the FMA-pipe polynomial
  p.pingpong = pingpong;
  p.debug = dbg;
  const int smem_bytes = OFF_BARS + static_cast<int>(sizeof(Bars6));
  static_assert(OFF_BARS + sizeof(Bars6) <= 227 * 1024, "attention6: shared memory budget");
  static long long* trace = nullptr;
  if (tracing) {
    if (!trace) PC_CHECK_CUDA(cudaMalloc(&trace, 8 * 2 * 16 * sizeof(long long)));
    PC_CHECK_CUDA(cudaMemsetAsync(trace, 0, 8 * 2 * 16 * sizeof(long long), stream));
    p.trace = trace;
  }
  CUtensorMap tmQ, tmT, tmO;
  const uint64_t rows = static_cast<uint64_t>(B) * L;
  const int tail_rows = p.split ? p.na : p.nb;  // split: the K / V box of one item; else part b of the shared item
  const uint64_t row_pitch = qkv_planar ? 128 : static_cast<uint64_t>(3 * d) * 2;
  const uint64_t plane_pitch = qkv_planar ? rows * 128 : 128;
  PC_TRY(make_tmap_f16_3d(&tmQ, qkv, 64, rows, 3 * heads, row_pitch, plane_pitch, 64, 128));
  PC_TRY(make_tmap_f16_3d(&tmT, qkv, 64, rows, 3 * heads, row_pitch, plane_pitch, 64, tail_rows));
  PC_TRY(make_tmap_f16_3d(&tmO, out, d, L, B, static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(L) * d * 2, 32, 32));
  const int sms = device_sm_count();
  const int grid = p.n_groups < sms ? p.n_groups : sms;
  if (xkey) PC_TRY((launch_variant6<false, 0, true>(grid, smem_bytes, stream, tmQ, tmT, tmO, p)));
  else if (causal) PC_TRY((launch_variant6<true, 0>(grid, smem_bytes, stream, tmQ, tmT, tmO, p)));
  else if (npoly == 0) PC_TRY((launch_variant6<false, 0>(grid, smem_bytes, stream, tmQ, tmT, tmO, p)));
  else if (npoly == 1) PC_TRY((launch_variant6<false, 1>(grid, smem_bytes, stream, tmQ, tmT, tmO, p)));
  else if (npoly == 2) PC_TRY((launch_variant6<false, 2>(grid, smem_bytes, stream, tmQ, tmT, tmO, p)));
  else if (npoly == 3) PC_TRY((launch_variant6<false, 3>(grid, smem_bytes, stream, tmQ, tmT, tmO, p)));
  else PC_TRY((launch_variant6<false, 4>(grid, smem_bytes, stream, tmQ, tmT, tmO, p)));
  if (tracing) {
    static int printed = 0;
    static long long h[8 * 2 * 16];
    PC_CHECK_CUDA(cudaStreamSynchronize(stream));
    PC_CHECK_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    if (printed++ == 3) {
      const long long t0 = h[0];
      fprintf(stderr, "[attn6 trace] items=%d L=%d (cycles since WG 0's first tile; CTA 0, warp quarter 0, thread 0 of the row)\n",
              p.items, L);
      fprintf(stderr, "tile WG    start | s_full  rowmax  turn    exp_b   exp_a   pv_done o_read  stored\n");
      for (int t = 0; t < 8; ++t)
        for (int w = 0; w < 2; ++w) {
          const long long* r = h + (t * 2 + w) * 16;
          if (!r[0]) continue;
          fprintf(stderr, "%3d  %d %8lld | +%5lld  +%5lld  +%5lld  +%5lld  +%5lld  +%5lld  +%5lld  +%5lld\n", t, w, r[0] - t0,
                  r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[7] - r[6], r[8] - r[7]);
        }
    }
  }
  return PC_OK;
}

}  // namespace pc
