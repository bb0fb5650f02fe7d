// C[M,N] = A[M,K] * W[N,K]^T on the 5th-gen tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators
// in TMEM), every operand and result tile moved by TMA through 128B-swizzled shared memory.
//
// This is the kernel behind every Linear on the hot path: MHA in_proj / out_proj and the MLP
// c_fc / c_proj of ResidualAttentionBlock (reference clip/model.py:169-190), the patch-embedding conv
// restated as a GEMM (clip/model.py:209,222), the visual / text projections (:235-236, :352),
// Adapter_FC's two bias-free Linears (model.py:84-87) and the query x prototype contraction inside
// P() (utils.py:230-233).
//
// Two instantiations of one persistent, warp-specialised kernel:
//   PAIR    two CTAs on the two SMs of a TPC (cluster 2x1x1) share a 256 x 256 tile: tcgen05.mma.cta_group::2
//           (UMMA 256x256x16). Each CTA stages its own 128 rows of A and HALF of the W tile (the tensor core
//           reads the other half from the peer's shared memory): 32 KB per k-block per SM instead of 48 KB.
//           The large encoder Linears run here (M >= 256 and N >= 256).
//   SINGLE  one CTA, 128 x 128 tile (UMMA 128x128x16): small problems (projections, adapter, tails).
// Warp roles (10 warps; the issue arbiter favours high warp ids, so the two single-lane drivers sit on top):
//   warps 0..7  epilogue: tcgen05.ld the accumulator (lane quarter = warp % 4, column half = warp / 4), add
//               bias / QuickGELU / residual in the reference's rounding order, write the 32-row x 128-byte
//               block into a swizzled staging buffer and hand it to a TMA store (bulk async group); the
//               residual block is TMA-loaded into the same buffer one block ahead. Two staging buffers per
//               warp, so the store of block i overlaps the math of block i+1.
//   warp 8      TMA producer: A and W tiles into a 5-deep mbarrier ring (PAIR: bytes of both CTAs are
//               credited to the leader's `full` barrier)
//   warp 9      MMA issuer (PAIR: leader CTA only; commits are multicast to both CTAs); owns TMEM
// Two accumulator stages (2 x BN TMEM columns) let the epilogue of tile i overlap the main loop of tile i+1.
// Tiles are walked n-fastest so concurrently running CTAs share A row-blocks in L2.
#include <stdlib.h>

#include "kernels.cuh"
#include "ptx.cuh"

namespace pc {

namespace {

constexpr int BM = 128;  // accumulator rows per CTA (TMEM lanes)
constexpr int BK = 64;   // 64 fp16 = 128 B = one swizzle row
constexpr int MAX_STAGES = 6;
constexpr int EPI_WARPS = 8;
constexpr int TMA_WARP = EPI_WARPS;
constexpr int MMA_WARP = EPI_WARPS + 1;
constexpr int GEMM_THREADS = (EPI_WARPS + 2) * 32;
constexpr int STG_BYTES = 32 * 128;  // one staging block: 32 rows x 128 B, 16-byte chunks XOR-swizzled by (row & 7)

// Shared-memory plan per epilogue kind. The residual epilogue needs two staging blocks per warp (the residual tile
// of block i+1 is TMA-prefetched while block i is stored) and gets a 5-deep operand ring; every other epilogue
// stores out of ONE staging block per warp, which buys a sixth ring stage (the ring is latency-bound: both the
// producer and the MMA issuer wait on each other at 5 stages).
template <int EPI>
struct SmemT {
  static constexpr int NSTG = (EPI == EPI_BIAS_RES) ? 2 : 1;  // staging blocks per epilogue warp
  static constexpr int STAGES = (EPI == EPI_BIAS_RES) ? 5 : 6;
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KB
  static constexpr int B_BYTES = 128 * BK * 2;  // 16 KB: 128 W rows per CTA in both modes
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OFF_STAGING = STAGES * STAGE_BYTES;
  static constexpr int STAGING_BYTES = EPI_WARPS * NSTG * STG_BYTES;
  static constexpr int OFF_BIAS = OFF_STAGING + STAGING_BYTES;
  static constexpr int OFF_SCALE = OFF_BIAS + 256 * 4;  // s_n of the LayerNorm-folded epilogues
  static constexpr int OFF_BARS = OFF_SCALE + 256 * 4;
  static constexpr int TOTAL = OFF_BARS + 512;  // + barrier block; the buffer is declared 1024-byte aligned
  static_assert(TOTAL <= 227 * 1024, "GEMM shared memory exceeds the 227 KB a CTA may use");
};

struct Bars {
  uint64_t full[MAX_STAGES];        // PAIR: used in the leader only (TMA bytes of both CTAs)
  uint64_t empty[MAX_STAGES];       // per CTA
  uint64_t tmem_full[2];            // per CTA
  uint64_t tmem_empty[2];           // PAIR: leader only, 16 arrivals; SINGLE: 8 arrivals
  uint64_t res_full[EPI_WARPS][2];  // residual block landed in staging buffer [warp][buf]
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 512, "barrier block");

// QuickGELU x * sigmoid(1.702 x) (clip/model.py:164-166) on a packed half2, with
// sigmoid(t) = 0.5 * tanh(t / 2) + 0.5 so one MUFU.TANH serves two elements. fp16 math throughout, as the
// reference's eager fp16 ops; tanh.approx.f16x2 is accurate to ~2^-11, i.e. about one fp16 ulp of the result.
__device__ __forceinline__ uint32_t quick_gelu_f16x2(uint32_t xb) {
  const __half2 x = *reinterpret_cast<const __half2*>(&xb);
  const __half2 t = __hmul2(x, __float2half2_rn(0.851f));
  const uint32_t tb = *reinterpret_cast<const uint32_t*>(&t);
  uint32_t thb;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(thb) : "r"(tb));
  const __half2 th = *reinterpret_cast<const __half2*>(&thb);
  const __half2 half = __float2half2_rn(0.5f);
  const __half2 y = __hmul2(x, __hfma2(th, half, half));
  return *reinterpret_cast<const uint32_t*>(&y);
}

__device__ __forceinline__ uint32_t hadd2_u32(uint32_t a, uint32_t b) {
  const __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// nn.ReLU on a packed fp16 pair (NaN propagates like torch.relu: max.NaN)
__device__ __forceinline__ uint32_t relu_f16x2(uint32_t a) {
  uint32_t r;
  asm("max.NaN.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(0u));
  return r;
}

#define TRACE(slot, val)                                                                   \
  do {                                                                                     \
    if ((g.debug & 8) && blockIdx.x == 0 && tidx < 64) g.trace[tidx * 16 + (slot)] = (val); \
  } while (0)

// Columns the MMA of the tile at column n0 computes: the tile width clipped to N, rounded up to the instruction's
// granularity (16; at least 32 for the 256-row pair shape). W rows past N are zero-filled by TMA. A launch that also
// accumulates row statistics over whole 64-column groups keeps full tiles (stale accumulator columns must not enter).
__device__ __forceinline__ int mma_cols(const GemmArgs& g, int n0, int bn, bool pair) {
  if (g.stats_out != nullptr) return bn;
  int n = (g.N - n0 + 15) & ~15;
  if (pair && n < 32) n = 32;
  return n < bn ? n : bn;
}

template <bool PAIR, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
               const GemmArgs g) {
  using Smem = SmemT<EPI>;
  constexpr int STAGES = Smem::STAGES;
  constexpr int NSTG = Smem::NSTG;
  constexpr int BN = PAIR ? 256 : 128;                  // tile columns (TMEM columns per accumulator stage)
  constexpr int TILE_M = PAIR ? 256 : 128;              // tile rows (both CTAs of a pair)
  constexpr int GRP_COLS = (EPI == EPI_F32) ? 32 : 64;  // columns per 128-byte staging row
  constexpr int NG = (BN / 2) / GRP_COLS;               // staging blocks per epilogue warp per tile
  constexpr uint32_t TX_BYTES = (PAIR ? 2 : 1) * Smem::STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // identical carve-up in both CTAs of a pair (UMMA / commit / TMA-barrier addressing relies on equal offsets)
  uint8_t* smem = smem_raw;  // 1024-byte aligned (128B-swizzle atoms)
  Bars* bars = reinterpret_cast<Bars*>(smem + Smem::OFF_BARS);
  float* sbias = reinterpret_cast<float*>(smem + Smem::OFF_BIAS);
  float* sscale = reinterpret_cast<float*>(smem + Smem::OFF_SCALE);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: keeps role-derived values in uniform registers
  const int lane = threadIdx.x & 31;
  const int rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = rank == 0;
  const int worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x;  // tile-loop index of this CTA (pair)
  const int num_workers = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int m_blks = (g.M + TILE_M - 1) / TILE_M;
  const int n_blks = (g.N + BN - 1) / BN;
  const int kb_per_tap = (g.K + BK - 1) / BK;
  const int k_blks = (g.conv_taps > 0 ? g.conv_taps : 1) * kb_per_tap;
  const int num_tiles = m_blks * n_blks;

  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmC);
    if (EPI == EPI_BIAS_RES) tma_prefetch_desc(&tmR);
  }
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&bars->full[i], 1);
        mbar_init(&bars->empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->tmem_full[i], 1);
        mbar_init(&bars->tmem_empty[i], (PAIR ? 2 : 1) * EPI_WARPS);
      }
      for (int i = 0; i < EPI_WARPS; ++i) {
        mbar_init(&bars->res_full[i][0], 1);
        mbar_init(&bars->res_full[i][1], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    if (PAIR) {
      tmem_alloc_pair(&bars->tmem_base, 2 * BN);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(&bars->tmem_base, 2 * BN);
      tmem_relinquish();
    }
  }
  griddep_launch_dependents();  // the next kernel of the stream may start its own prologue from here on
  const long long t_entry = ((g.debug & 16) && threadIdx.x == 0) ? static_cast<long long>(globaltimer_ns()) : 0;
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // barriers of BOTH CTAs initialised before any remote arrive / multicast commit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  griddep_wait();  // operands (and the residual / output buffers) belong to the previous kernel until here
  if ((g.debug & 16) && threadIdx.x == 0) {  // bring-up: per-CTA wall-clock timeline of the whole grid
    g.trace[3 * blockIdx.x + 0] = t_entry;
    g.trace[3 * blockIdx.x + 1] = static_cast<long long>(globaltimer_ns());
  }

  if (warp == TMA_WARP) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp walks the loop (warp-uniform control flow keeps coordinates and barrier addresses in
    // uniform registers); one elected lane issues. A single-lane `if (lane == 0)` region would make ptxas
    // wrap every UTMALDG / UTCHMMA in a value-uniformity waterfall (ELECT / R2UR / BRA.U.ANY).
    {
      int stage = 0;
      uint32_t phase = 0;
      int tidx = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++tidx) {
        const int m0 = (tile / n_blks) * TILE_M + rank * BM;
        // the pair's B operand is split evenly: this CTA stages rows [rank * n_mma / 2, +n_mma / 2) of the W tile
        const int n0 = (tile % n_blks) * BN + (PAIR ? rank * (mma_cols(g, (tile % n_blks) * BN, BN, PAIR) >> 1) : 0);
        int px0 = 0, py0 = 0, pn0 = 0;  // patch mode: first pixel / image of this CTA's 128 rows
        if (g.conv_w > 0) {
          const int patch = m0 >> 7;
          const int t2 = patch / g.px;
          px0 = (patch - t2 * g.px) * g.pw;
          pn0 = t2 / g.py;
          py0 = (t2 - pn0 * g.py) * g.ph;
          pn0 *= g.pn;
        }
        long long w_empty = 0;
        for (int kb = 0; kb < k_blks; ++kb) {
          const long long tw = (g.debug & 8) ? clock64() : 0;
          mbar_wait(&bars->empty[stage], phase ^ 1);
          if (g.debug & 8) w_empty += clock64() - tw;
          uint8_t* sA = smem + stage * Smem::STAGE_BYTES;
          uint8_t* sB = sA + Smem::A_BYTES;
          int a_col = kb * BK, a_row = m0, w_col = kb * BK;
          if (g.conv_taps > 0) {  // implicit 3x3 convolution: k-block = (tap, channel block), A rows shifted per tap
            const int tap = kb / kb_per_tap, kc = (kb - tap * kb_per_tap) * BK;
            a_col = kc;
            w_col = tap * g.K + kc;
            a_row = m0 + (tap / 3 - 1) * g.conv_pitch + (tap % 3 - 1);
          }
          if (g.conv_w > 0) {
            if (elect_one()) {
              const int tap = kb / kb_per_tap;
              if (PAIR) {
                if (leader) mbar_arrive_expect_tx(&bars->full[stage], TX_BYTES);
                tma_load_4d_pair(sA, &tmA, &bars->full[stage], a_col, px0 + tap % 3 - 1, py0 + tap / 3 - 1, pn0, kEvictNormal);
                tma_load_2d_pair(sB, &tmW, &bars->full[stage], w_col, n0, kEvictLast);
              } else {
                mbar_arrive_expect_tx(&bars->full[stage], TX_BYTES);
                tma_load_4d_hint(sA, &tmA, &bars->full[stage], a_col, px0 + tap % 3 - 1, py0 + tap / 3 - 1, pn0, kEvictNormal);
                tma_load_2d_hint(sB, &tmW, &bars->full[stage], w_col, n0, kEvictLast);
              }
            }
          } else if (elect_one()) {
            if (g.debug & 2) {
              if (leader) mbar_arrive(&bars->full[stage]);
            } else if (PAIR) {
              if (leader) mbar_arrive_expect_tx(&bars->full[stage], TX_BYTES);
              tma_load_2d_pair(sA, &tmA, &bars->full[stage], a_col, a_row, kEvictNormal);
              tma_load_2d_pair(sB, &tmW, &bars->full[stage], w_col, n0, kEvictLast);
            } else {
              mbar_arrive_expect_tx(&bars->full[stage], TX_BYTES);
              tma_load_2d_hint(sA, &tmA, &bars->full[stage], a_col, a_row, kEvictNormal);
              tma_load_2d_hint(sB, &tmW, &bars->full[stage], w_col, n0, kEvictLast);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (lane == 0) {
          TRACE(7, w_empty);
          TRACE(8, clock64());
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------ MMA issuer (PAIR: leader CTA only)
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      int tidx = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++tidx) {
        if (lane == 0) {
          TRACE(0, clock64());
          TRACE(9, static_cast<long long>(globaltimer_ns()));
        }
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        if (lane == 0) TRACE(1, clock64());
        long long w_full = 0;
        const uint32_t d_tmem = tmem_base + as * BN;
        // instruction shape: only the columns this tile really has (N = 96 / 192 / 384 convolution widths)
        const uint32_t idesc = umma_idesc_f16(TILE_M, mma_cols(g, (tile % n_blks) * BN, BN, PAIR), 0, 0);
        for (int kb = 0; kb < k_blks; ++kb) {
          // k-steps of this block that hold real columns (K = 96 per tap: the second block is half empty)
          const int k_left = g.K - (kb % kb_per_tap) * BK;
          const int ksteps = k_left >= BK ? BK / 16 : (k_left + 15) >> 4;
          const long long tw = (g.debug & 8) ? clock64() : 0;
          mbar_wait(&bars->full[stage], phase);
          if (g.debug & 8) w_full += clock64() - tw;
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Smem::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Smem::A_BYTES;
          if (elect_one()) {
            if (!(g.debug & 4)) {
              const uint64_t da = umma_desc_kmajor_sw128(a_addr);
              const uint64_t db = umma_desc_kmajor_sw128(b_addr);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                if (k < ksteps) {
                  // advancing 16 fp16 along K inside the 128-byte swizzle row = +32 B = +2 in the address field
                  if (PAIR) umma_f16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                  else umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                }
              }
            }
            // frees this smem stage (in both CTAs) once the MMAs above retire
            if (PAIR) umma_commit_pair(&bars->empty[stage], 0x3);
            else umma_commit(&bars->empty[stage]);
            // accumulator complete -> epilogue warps (of both CTAs)
            if (kb == k_blks - 1) {
              if (PAIR) umma_commit_pair(&bars->tmem_full[as], 0x3);
              else umma_commit(&bars->tmem_full[as]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (lane == 0) {
          TRACE(2, w_full);
          TRACE(3, clock64());
          TRACE(10, static_cast<long long>(globaltimer_ns()));
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 0..7), this CTA's 128 rows
    const int quarter = warp & 3;  // TMEM lanes 32 * quarter .. + 31 (hardware: warp id % 4)
    const int chalf = warp >> 2;   // column half of the tile
    const int ep_tid = threadIdx.x;
    uint8_t* stg0 = smem + Smem::OFF_STAGING + warp * NSTG * STG_BYTES;
    uint64_t* res_bar = bars->res_full[warp];
    const int row_off = rank * BM + quarter * 32;  // first row of this warp inside the tile
    const int col_off = chalf * (BN / 2);          // first column of this warp inside the tile
    uint32_t sg = 0;  // running staging-block counter: buffer = sg & 1, residual-barrier parity = (sg >> 1) & 1
    int as = 0;
    uint32_t aphase = 0;

    constexpr bool LN = (EPI == EPI_LN_BIAS || EPI == EPI_LN_QGELU);
    const float inv_k = 1.0f / static_cast<float>(g.K);
    // one bias (LN: c_n and s_n) column per epilogue thread (BN <= 256 = epilogue threads), fetched a tile ahead
    float bias_cur = 0.0f, scale_cur = 0.0f;
    auto fetch_cols = [&](int t) {
      const int c = (t % n_blks) * BN + ep_tid;
      bias_cur = 0.0f;
      scale_cur = 0.0f;
      if (t < num_tiles && ep_tid < BN && c < g.N) {
        if (LN) {
          bias_cur = g.ln_c[c];
          scale_cur = g.ln_s[c];
        } else if (g.bias_f32 != nullptr) {
          bias_cur = g.bias_f32[c];
        } else if (g.bias != nullptr) {
          bias_cur = __half2float(g.bias[c]);
        }
      }
    };
    if (worker < num_tiles) {
      fetch_cols(worker);
      if (EPI == EPI_BIAS_RES && elect_one()) {  // residual block of the very first staging block
        mbar_arrive_expect_tx(&res_bar[0], STG_BYTES);
        tma_load_2d(stg0, &tmR, &res_bar[0], (worker % n_blks) * BN + col_off,
                    (worker / n_blks) * TILE_M + row_off);
      }
    }
    int tidx = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++tidx) {
      const int m0 = (tile / n_blks) * TILE_M;
      const int n0 = (tile % n_blks) * BN;
      const int tile_next = tile + num_workers;
      named_bar_sync(1, EPI_WARPS * 32);  // previous tile's bias fully consumed
      if (ep_tid < BN) {
        sbias[ep_tid] = bias_cur;
        if (LN) sscale[ep_tid] = scale_cur;
      }
      named_bar_sync(1, EPI_WARPS * 32);
      fetch_cols(tile_next);  // next tile's columns: in flight while this tile is processed
      // LayerNorm folding: this thread's row statistics -> out = acc * ln_a + (s_n * ln_b + c_n)
      float ln_a = 0.0f, ln_b = 0.0f;
      const int my_row_idx = m0 + row_off + lane;
      if (LN && my_row_idx < g.M) {
        const float2* st = reinterpret_cast<const float2*>(g.ln_stats) + static_cast<size_t>(my_row_idx) * g.ln_parts;
        float sx = 0.0f, sq = 0.0f;
        for (int q = 0; q < g.ln_parts; ++q) {  // fixed order: bit-reproducible statistics
          const float2 t = st[q];
          sx += t.x;
          sq += t.y;
        }
        const float mean = sx * inv_k;
        const float var = fmaxf(sq * inv_k - mean * mean, 0.0f);
        ln_a = rsqrtf(var + 1e-5f);
        ln_b = -mean * ln_a;
      }
      float st_sum = 0.0f, st_sq = 0.0f;  // EPI_BIAS_RES + stats_out: statistics of the row segment written here
      if (threadIdx.x == 0) TRACE(4, clock64());
      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      if (threadIdx.x == 0) TRACE(5, clock64());
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + col_off;
#pragma unroll 1
      for (int grp = 0; grp < NG; ++grp, ++sg) {
        uint8_t* stg = stg0 + (NSTG == 2 ? (sg & 1) : 0) * STG_BYTES;
        uint8_t* my_row = stg + lane * 128;
        const int gcol0 = n0 + col_off + grp * GRP_COLS;
        // ---- accumulator block -> registers
        uint32_t v[2][32];
        if (!(g.debug & 1)) {
          tmem_ld_32x32(t_row + grp * GRP_COLS, v[0]);
          if (EPI != EPI_F32) tmem_ld_32x32(t_row + grp * GRP_COLS + 32, v[1]);
          tmem_wait_ld();
        }
        if (grp == NG - 1) {  // last TMEM read of this accumulator stage by this warp: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(&bars->tmem_empty[as], 0);
            else mbar_arrive(&bars->tmem_empty[as]);
          }
        }
        // ---- staging buffer management. The elected lane owns this warp's bulk groups (elect.sync picks the
        // same lane for the same member mask every time); an elect-guarded region lets ptxas feed UTMALDG /
        // UTMASTG from uniform registers without a per-value waterfall loop.
        if (elect_one()) {
          if (EPI == EPI_BIAS_RES) {
            tma_store_wait_read<0>();  // every earlier store has drained its buffer: the other buffer is free
            // residual block of the NEXT staging block (possibly the first block of the next tile)
            int nt = tile, ngrp = grp + 1;
            if (ngrp == NG) {
              nt = tile_next;
              ngrp = 0;
            }
            if (nt < num_tiles) {
              uint64_t* nb = &res_bar[(sg + 1) & 1];
              mbar_arrive_expect_tx(nb, STG_BYTES);
              tma_load_2d(stg0 + ((sg + 1) & 1) * STG_BYTES, &tmR, nb,
                          (nt % n_blks) * BN + col_off + ngrp * GRP_COLS, (nt / n_blks) * TILE_M + row_off);
            }
          } else {
            tma_store_wait_read<0>();  // single staging block: the previous store has drained it
          }
        }
        __syncwarp();
        if (EPI == EPI_BIAS_RES) mbar_wait(&res_bar[sg & 1], (sg >> 1) & 1);
        // ---- bias / activation / residual, rounded like the eager reference, into the swizzled row `lane`
        if (!(g.debug & 1)) {
          const float* bcol = sbias + col_off + grp * GRP_COLS;
          if (EPI == EPI_F32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(bcol + 4 * j);
              float4 o;
              o.x = __uint_as_float(v[0][4 * j + 0]) + b.x;
              o.y = __uint_as_float(v[0][4 * j + 1]) + b.y;
              o.z = __uint_as_float(v[0][4 * j + 2]) + b.z;
              o.w = __uint_as_float(v[0][4 * j + 3]) + b.w;
              *reinterpret_cast<float4*>(my_row + ((j ^ (lane & 7)) << 4)) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // 16-byte chunk j = columns 8j .. 8j+7 of the block
              const uint32_t* vv = &v[j >> 2][8 * (j & 3)];
              const float4 b0 = *reinterpret_cast<const float4*>(bcol + 8 * j);
              const float4 b1 = *reinterpret_cast<const float4*>(bcol + 8 * j + 4);
              uint4 pk;
              if (LN) {
                const float* scol = sscale + col_off + grp * GRP_COLS;
                const float4 s0 = *reinterpret_cast<const float4*>(scol + 8 * j);
                const float4 s1 = *reinterpret_cast<const float4*>(scol + 8 * j + 4);
                pk.x = pack_half2(fmaf(__uint_as_float(vv[0]), ln_a, fmaf(s0.x, ln_b, b0.x)),
                                  fmaf(__uint_as_float(vv[1]), ln_a, fmaf(s0.y, ln_b, b0.y)));
                pk.y = pack_half2(fmaf(__uint_as_float(vv[2]), ln_a, fmaf(s0.z, ln_b, b0.z)),
                                  fmaf(__uint_as_float(vv[3]), ln_a, fmaf(s0.w, ln_b, b0.w)));
                pk.z = pack_half2(fmaf(__uint_as_float(vv[4]), ln_a, fmaf(s1.x, ln_b, b1.x)),
                                  fmaf(__uint_as_float(vv[5]), ln_a, fmaf(s1.y, ln_b, b1.y)));
                pk.w = pack_half2(fmaf(__uint_as_float(vv[6]), ln_a, fmaf(s1.z, ln_b, b1.z)),
                                  fmaf(__uint_as_float(vv[7]), ln_a, fmaf(s1.w, ln_b, b1.w)));
              } else {
                pk.x = pack_half2(__uint_as_float(vv[0]) + b0.x, __uint_as_float(vv[1]) + b0.y);
                pk.y = pack_half2(__uint_as_float(vv[2]) + b0.z, __uint_as_float(vv[3]) + b0.w);
                pk.z = pack_half2(__uint_as_float(vv[4]) + b1.x, __uint_as_float(vv[5]) + b1.y);
                pk.w = pack_half2(__uint_as_float(vv[6]) + b1.z, __uint_as_float(vv[7]) + b1.w);
              }
              if (EPI == EPI_BIAS_QGELU || EPI == EPI_LN_QGELU) {
                pk.x = quick_gelu_f16x2(pk.x);
                pk.y = quick_gelu_f16x2(pk.y);
                pk.z = quick_gelu_f16x2(pk.z);
                pk.w = quick_gelu_f16x2(pk.w);
              }
              uint4* slot = reinterpret_cast<uint4*>(my_row + ((j ^ (lane & 7)) << 4));
              if (EPI == EPI_BIAS_RES) {  // x + f16(acc + bias): the residual chunk is already in the slot
                const uint4 r = *slot;
                pk.x = hadd2_u32(r.x, pk.x);
                pk.y = hadd2_u32(r.y, pk.y);
                pk.z = hadd2_u32(r.z, pk.z);
                pk.w = hadd2_u32(r.w, pk.w);
                if (g.stats_out != nullptr) {  // statistics of the fp16 values the next LayerNorm will see
                  const uint32_t w4[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
                    st_sum += f.x + f.y;
                    st_sq = fmaf(f.x, f.x, fmaf(f.y, f.y, st_sq));
                  }
                }
              }
              if ((EPI == EPI_BIAS || EPI == EPI_BIAS_RES) && g.relu) {
                pk.x = relu_f16x2(pk.x);
                pk.y = relu_f16x2(pk.y);
                pk.z = relu_f16x2(pk.z);
                pk.w = relu_f16x2(pk.w);
              }
              *slot = pk;
            }
          }
        }
        fence_async_smem();  // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (elect_one()) {
          if (!(g.debug & 1) && gcol0 < g.N && m0 + row_off < g.M) {
            if (g.conv_w > 0) {  // patch mode: this warp's 32 rows are a sub-box of the CTA's pixel patch
              const int patch = (m0 + rank * BM) >> 7;
              const int t2 = patch / g.px;
              const int tn = t2 / g.py;
              const int r0 = quarter * 32 / g.pw;  // patch rows (of pw pixels) before this warp's first row
              tma_store_4d(&tmC, stg, gcol0, (patch - t2 * g.px) * g.pw, (t2 - tn * g.py) * g.ph + r0 % g.ph,
                           tn * g.pn + r0 / g.ph);
            } else if (g.c_planar) tma_store_3d(&tmC, stg, 0, m0 + row_off, gcol0 / GRP_COLS);  // plane = 64-column block
            else tma_store_2d(&tmC, stg, gcol0, m0 + row_off);
          }
          tma_store_commit();
        }
      }
      if (EPI == EPI_BIAS_RES && g.stats_out != nullptr && my_row_idx < g.M) {
        // this thread's column segment of the row: partial pair (tile column block, column half)
        float2* dst = reinterpret_cast<float2*>(g.stats_out) + static_cast<size_t>(my_row_idx) * (2 * n_blks) +
                      2 * (tile % n_blks) + chalf;
        *dst = make_float2(st_sum, st_sq);
      }
      if (threadIdx.x == 0) TRACE(6, clock64());
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (elect_one()) tma_store_wait_all<0>();  // results written before the CTA (and its shared memory) goes away
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  else __syncthreads();
  if ((g.debug & 16) && threadIdx.x == 0) g.trace[3 * blockIdx.x + 2] = static_cast<long long>(globaltimer_ns());
  if (warp == MMA_WARP) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 2 * BN);
    else tmem_dealloc(tmem_base, 2 * BN);
  }
}

int debug_flags() {
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("PC_GEMM_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  return dbg;
}

// PC_GEMM_SINGLE_CTA=1 keeps every problem on the single-CTA kernel (A/B comparisons in tools/gpu_probe.py).
bool force_single_cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PC_GEMM_SINGLE_CTA");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// 128-pixel patch (bw x bh pixels of bn images) for the patch-mode convolution: bw divides 32 (a warp's 32 staging rows
// are whole patch rows), the patches tile the image exactly.
bool pick_patch(int h, int w, int* bw, int* bh, int* bn) {
  if (h <= 0 || w <= 0) return false;
  if (w % 32 == 0 && h % 4 == 0) { *bw = 32; *bh = 4; *bn = 1; return true; }
  if (w % 16 == 0 && h % 8 == 0) { *bw = 16; *bh = 8; *bn = 1; return true; }
  if (w % 8 == 0 && h % 8 == 0) { *bw = 8; *bh = 8; *bn = 2; return true; }
  if (w % 4 == 0 && h % 4 == 0) { *bw = 4; *bh = 4; *bn = 8; return true; }
  return false;
}

template <bool PAIR, int EPI>
int launch_impl(const GemmArgs& a0, cudaStream_t stream) {
  using Smem = SmemT<EPI>;
  constexpr int BN = PAIR ? 256 : 128;
  constexpr int TILE_M = PAIR ? 256 : 128;
  static int configured[kMaxDevices];
  auto kern = gemm_tn_kernel<PAIR, EPI>;
  PC_CHECK_CUDA(ensure_dynamic_smem(kern, Smem::TOTAL, configured));
  GemmArgs a = a0;
  a.debug = debug_flags();
  CUtensorMap tmA, tmW, tmC, tmR;
  if (a.conv_w > 0 && (EPI != EPI_BIAS || a.c_planar)) {
    set_error("gemm: the patch-mode convolution has the bias (+ ReLU) epilogue only");
    return PC_ERR_ARG;
  } else if (a.conv_w > 0) {
    PC_REQUIRE(pick_patch(a.conv_h, a.conv_w, &a.pw, &a.ph, &a.pn), PC_ERR_ARG,
               "gemm: a %d x %d activation cannot be cut into 128-pixel patches", a.conv_h, a.conv_w);
    a.px = a.conv_w / a.pw;
    a.py = a.conv_h / a.ph;
    a.M = a.px * a.py * ((a.conv_n + a.pn - 1) / a.pn) * BM;
    PC_TRY(make_tmap_f16_nhwc(&tmA, a.A, a.K, a.conv_w, a.conv_h, a.conv_n, static_cast<uint64_t>(a.lda) * 2, a.pw, a.ph, a.pn));
  } else {
    PC_TRY(make_tmap_2d(&tmA, a.A, 2, a.K, a.M, static_cast<uint64_t>(a.lda) * 2, BK, BM));
  }
  PC_TRY(make_tmap_2d(&tmW, a.W, 2, a.conv_taps > 0 ? a.conv_taps * a.K : a.K, a.N, static_cast<uint64_t>(a.ldw) * 2, BK,
                      128));
  if (EPI == EPI_F32) {
    PC_TRY(make_tmap_2d(&tmC, a.C, 4, a.N, a.M, static_cast<uint64_t>(a.ldc) * 4, 32, 32));
  } else if (a.conv_w > 0) {  // a warp's 32 staging rows = pw x (32 / pw) pixels of (32 / pw) / ph ... images
    const int sw = a.pw, sh = (32 / sw) < a.ph ? (32 / sw) : a.ph, sn = 32 / (sw * sh);
    PC_TRY(make_tmap_f16_nhwc(&tmC, a.C, a.N, a.conv_w, a.conv_h, a.conv_n, static_cast<uint64_t>(a.ldc) * 2, sw, sh, sn));
  } else if (a.c_planar) {
    PC_TRY(make_tmap_f16_3d(&tmC, a.C, 64, a.M, a.N / 64, 128, static_cast<uint64_t>(a.M) * 128, 64, 32));
  } else {
    PC_TRY(make_tmap_2d(&tmC, a.C, 2, a.N, a.M, static_cast<uint64_t>(a.ldc) * 2, 64, 32));
  }
  if (EPI == EPI_BIAS_RES) {
    PC_TRY(make_tmap_2d(&tmR, a.residual, 2, a.N, a.M, static_cast<uint64_t>(a.ldr) * 2, 64, 32));
  } else {
    tmR = tmC;
  }
  const int m_blks = (a.M + TILE_M - 1) / TILE_M;
  const int n_blks = (a.N + BN - 1) / BN;
  const int tiles = m_blks * n_blks;
  const int max_workers = PAIR ? device_sm_count() / 2 : device_sm_count();
  const int workers = tiles < max_workers ? tiles : max_workers;
  static long long* trace = nullptr;
  if (a.debug & 8) {
    if (!trace) PC_CHECK_CUDA(cudaMalloc(&trace, 64 * 16 * sizeof(long long)));
    PC_CHECK_CUDA(cudaMemsetAsync(trace, 0, 64 * 16 * sizeof(long long), stream));
    a.trace = trace;
  }
  // bring-up (PC_GEMM_DEBUG=16): globaltimer at kernel entry / after griddep_wait / at exit of every CTA, 16 launches deep
  static long long* trace_all = nullptr;
  static int launch_no = 0;
  constexpr int kSlots = 16, kSlotLen = 3 * 304;
  if (a.debug & 16) {
    if (!trace_all) {
      PC_CHECK_CUDA(cudaMalloc(&trace_all, kSlots * kSlotLen * sizeof(long long)));
      PC_CHECK_CUDA(cudaMemset(trace_all, 0, kSlots * kSlotLen * sizeof(long long)));
    }
    a.trace = trace_all + (launch_no % kSlots) * kSlotLen;
  }
  PC_CHECK_CUDA(launch_pdl(kern, dim3(PAIR ? 2 * workers : workers), dim3(GEMM_THREADS), Smem::TOTAL, stream, PAIR ? 2 : 1,
                           tmA, tmW, tmC, tmR, a));
  if (a.debug & 16) {
    static int report_at = -1;
    if (report_at < 0) {
      const char* e = getenv("PC_GEMM_TRACE_AT");
      report_at = e ? atoi(e) : 40;
    }
    if (++launch_no == report_at) {
      static long long h[kSlots * kSlotLen];
      PC_CHECK_CUDA(cudaStreamSynchronize(stream));
      PC_CHECK_CUDA(cudaMemcpy(h, trace_all, sizeof(h), cudaMemcpyDeviceToHost));
      const int ctas = PAIR ? 2 * workers : workers;
      fprintf(stderr, "[gemm grid trace] M=%d N=%d K=%d pair=%d epi=%d, %d CTAs (ns; launches in stream order)\n", a.M, a.N, a.K,
              (int)PAIR, EPI, ctas);
      fprintf(stderr, "launch | first entry  first go   last go | first exit  median exit  last exit | span(go..exit) gap to prev\n");
      long long prev_end = 0;
      for (int l = report_at - kSlots + 1; l < report_at; ++l) {  // the slot of launch l (0-based launch index l)
        if (l < 0) continue;
        const long long* r = h + (l % kSlots) * kSlotLen;
        long long e0 = r[0], g0 = r[1], g1 = r[1], x0 = r[2], x1 = r[2];
        static long long ex[304];
        for (int c = 0; c < ctas; ++c) {
          e0 = r[3 * c] < e0 ? r[3 * c] : e0;
          g0 = r[3 * c + 1] < g0 ? r[3 * c + 1] : g0;
          g1 = r[3 * c + 1] > g1 ? r[3 * c + 1] : g1;
          x0 = r[3 * c + 2] < x0 ? r[3 * c + 2] : x0;
          x1 = r[3 * c + 2] > x1 ? r[3 * c + 2] : x1;
          ex[c] = r[3 * c + 2];
        }
        for (int i = 1; i < ctas; ++i) {  // insertion sort: median exit
          long long v = ex[i];
          int j = i - 1;
          for (; j >= 0 && ex[j] > v; --j) ex[j + 1] = ex[j];
          ex[j + 1] = v;
        }
        fprintf(stderr, "%6d | %11lld %9lld %9lld | %10lld %12lld %10lld | %14lld %11lld\n", l, e0 - g0, 0LL, g1 - g0, x0 - g0,
                ex[ctas / 2] - g0, x1 - g0, x1 - g0, prev_end ? g0 - prev_end : 0LL);
        prev_end = x1;
      }
    }
  }
  if (a.debug & 8) {
    static int printed = 0;
    long long h[64 * 16];
    PC_CHECK_CUDA(cudaStreamSynchronize(stream));
    PC_CHECK_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    static int trace_at = -1;
    if (trace_at < 0) {
      const char* e = getenv("PC_GEMM_TRACE_AT");
      trace_at = e ? atoi(e) : 3;
    }
    if (printed++ == trace_at) {  // a warmed-up launch
      const long long t0 = h[0];
      fprintf(stderr, "[gemm trace] M=%d N=%d K=%d pair=%d epi=%d  (cycles, CTA 0)\n", a.M, a.N, a.K, (int)PAIR, EPI);
      fprintf(stderr, "tile  mma_start tmem_wait  full_wait   mma_end | epi_wait_from epi_start  epi_end | prod_empty_wait prod_end\n");
      int last = 0;
      while (last + 1 < 64 && h[(last + 1) * 16 + 3]) ++last;
      const double ns = static_cast<double>(h[last * 16 + 10] - h[9]);
      fprintf(stderr, "mma span %lld cycles in %.0f ns -> SM clock %.0f MHz\n", h[last * 16 + 3] - t0, ns,
              ns > 0 ? (h[last * 16 + 3] - t0) / ns * 1e3 : 0.0);
      for (int t = 0; t < 64 && h[t * 16 + 3]; ++t) {
        const long long* r = h + t * 16;
        fprintf(stderr, "%4d %10lld %9lld %10lld %9lld | %13lld %9lld %8lld | %15lld %8lld\n", t, r[0] - t0, r[1] - r[0], r[2],
                r[3] - t0, r[4] - t0, r[5] - t0, r[6] - t0, r[7], r[8] - t0);
      }
    }
  }
  return PC_OK;
}

template <bool PAIR>
int dispatch_epi(const GemmArgs& a, int epi, cudaStream_t stream) {
  switch (epi) {
    case EPI_BIAS: return launch_impl<PAIR, EPI_BIAS>(a, stream);
    case EPI_BIAS_QGELU: return launch_impl<PAIR, EPI_BIAS_QGELU>(a, stream);
    case EPI_BIAS_RES: return launch_impl<PAIR, EPI_BIAS_RES>(a, stream);
    case EPI_F32: return launch_impl<PAIR, EPI_F32>(a, stream);
    case EPI_LN_BIAS: return launch_impl<PAIR, EPI_LN_BIAS>(a, stream);
    case EPI_LN_QGELU: return launch_impl<PAIR, EPI_LN_QGELU>(a, stream);
    default: set_error("unknown GEMM epilogue %d", epi); return PC_ERR_ARG;
  }
}

}  // namespace

// Smallest N that goes to the CTA-pair (256 x 256 tile) kernel; PC_GEMM_PAIR_MIN_N overrides it (A/B timing).
static int pair_min_n() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PC_GEMM_PAIR_MIN_N");
    v = (e && atoi(e) > 0) ? atoi(e) : 192;  // 192: +6 % on the RN50x16 tower (planes = 192 layers), no ViT shape in [192, 256)
  }
  return v;
}
static bool use_pair(int M, int N) { return M >= 256 && N >= pair_min_n() && !force_single_cta(); }

bool conv_patch_supported(int h, int w) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("PC_NO_CONV_PATCH");  // 1: every 3x3 convolution through the bordered copy (A/B timing)
    off = (e && e[0] == '1') ? 1 : 0;
  }
  int bw, bh, bn;
  return !off && pick_patch(h, w, &bw, &bh, &bn);
}

int gemm_stats_parts(int M, int N) {
  const int bn = use_pair(M, N) ? 256 : 128;
  return 2 * ((N + bn - 1) / bn);
}

int launch_gemm(const GemmArgs& a, int epilogue, cudaStream_t stream) {
  static int log_shapes = -1;  // PC_GEMM_LOG=1: one stderr line per launch (joined with an ncu launch list by tools/rn_breakdown.py)
  if (log_shapes < 0) {
    const char* e = getenv("PC_GEMM_LOG");
    log_shapes = (e && e[0] == '1') ? 1 : 0;
  }
  if (log_shapes)
    fprintf(stderr, "[gemm] M=%d N=%d K=%d epi=%d taps=%d patch=%d res=%d\n",
            a.conv_w > 0 ? a.conv_n * a.conv_h * a.conv_w : a.M, a.N, a.K, epilogue, a.conv_taps, a.conv_w > 0 ? 1 : 0,
            a.residual != nullptr ? 1 : 0);
  PC_REQUIRE((a.M > 0 || a.conv_w > 0) && a.N > 0 && a.K > 0, PC_ERR_ARG, "gemm: empty problem %dx%dx%d", a.M, a.N, a.K);
  PC_REQUIRE(a.A && a.W && a.C, PC_ERR_ARG, "gemm: null operand");
  PC_REQUIRE(a.K % 8 == 0 && a.lda % 8 == 0 && a.ldw % 8 == 0, PC_ERR_ALIGN,
             "gemm: K/lda/ldw (%d/%d/%d) must be multiples of 8 fp16 (16-byte TMA rows)", a.K, a.lda, a.ldw);
  // The result leaves through a TMA store: rows of C must start on 16-byte boundaries (columns are clipped at N).
  const int chunk = (epilogue == EPI_F32) ? 4 : 8;
  PC_REQUIRE(a.ldc % chunk == 0 && a.ldc >= a.N && (reinterpret_cast<uintptr_t>(a.C) & 15) == 0, PC_ERR_ALIGN,
             "gemm: output needs a 16-byte aligned C and ldc (%d) >= N (%d), a multiple of %d", a.ldc, a.N, chunk);
  PC_REQUIRE(!a.c_planar || (epilogue != EPI_F32 && epilogue != EPI_BIAS_RES && a.N % 64 == 0), PC_ERR_ARG,
             "gemm: planar output needs an fp16 non-residual epilogue and N %% 64 == 0 (N = %d)", a.N);
  if (epilogue == EPI_BIAS_RES) {
    PC_REQUIRE(a.residual != nullptr && a.ldr % 8 == 0 && a.ldr >= a.N &&
                   (reinterpret_cast<uintptr_t>(a.residual) & 15) == 0,
               PC_ERR_ALIGN, "gemm: residual must be non-null, 16-byte aligned, ldr %% 8 == 0");
  }
  if (epilogue == EPI_LN_BIAS || epilogue == EPI_LN_QGELU) {
    PC_REQUIRE(a.ln_stats != nullptr && a.ln_s != nullptr && a.ln_c != nullptr && a.ln_parts >= 1, PC_ERR_ARG,
               "gemm: the LayerNorm-folded epilogues need ln_stats (ln_parts >= 1), ln_s and ln_c");
  }
  if (a.conv_w > 0) {
    PC_REQUIRE(a.conv_taps == 9 && a.conv_h > 0 && a.conv_n > 0 && a.ldw >= 9 * a.K, PC_ERR_ARG,
               "gemm: patch-mode convolution needs 9 taps, h / w / n and W [N, 9*K]");
    return use_pair(a.conv_n * a.conv_h * a.conv_w, a.N) ? dispatch_epi<true>(a, epilogue, stream)
                                                         : dispatch_epi<false>(a, epilogue, stream);
  }
  PC_REQUIRE(a.conv_taps == 0 || (a.conv_taps == 9 && a.conv_pitch >= 3 && a.ldw >= 9 * a.K), PC_ERR_ARG,
             "gemm: implicit convolution needs 9 taps, a row pitch and W [N, 9*K] (taps %d, pitch %d, ldw %d)",
             a.conv_taps, a.conv_pitch, a.ldw);
  if (use_pair(a.M, a.N)) return dispatch_epi<true>(a, epilogue, stream);
  return dispatch_epi<false>(a, epilogue, stream);
}

}  // namespace pc
