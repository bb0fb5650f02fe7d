// C[M,N] = A[M,K] * W[N,K]^T on the 5th-gen tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators
// in TMEM), operands staged by TMA into 128B-swizzled shared memory through a 4-deep mbarrier ring.
//
// This is the kernel behind every Linear on the hot path: MHA in_proj / out_proj and the MLP
// c_fc / c_proj of ResidualAttentionBlock (reference clip/model.py:169-190), the patch-embedding conv
// restated as a GEMM (clip/model.py:209,222), the visual / text projections (:235-236, :352),
// Adapter_FC's two bias-free Linears (model.py:84-87) and the query x prototype contraction inside
// P() (utils.py:230-233).
//
// Shape of the kernel (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer  : one lane issues cp.async.bulk.tensor for the A (128x64) and W (BNx64) tiles
//   warp 1      MMA issuer    : one lane issues 4 x tcgen05.mma (M=128, N=BN, K=16) per k-block; owns TMEM
//   warps 2..9  epilogue      : tcgen05.ld the 128xBN fp32 accumulator (two warps per TMEM lane quarter),
//                               add bias / QuickGELU / residual in the reference's rounding order, transpose
//                               through an XOR-swizzled smem staging tile and write 128-byte coalesced rows.
// Two accumulator stages (2 x BN TMEM columns) let the epilogue of tile i overlap the main loop of tile i+1.
// Tiles are walked n-fastest so the CTAs of one wave share the same A row-blocks in L2.
#include "kernels.cuh"
#include "ptx.cuh"

namespace pc {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = 128 B = one swizzle row
constexpr int STAGES = 4;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int STG_WARP_BYTES = 32 * 128;  // 32 rows x 128 B, 16-byte chunks XOR-swizzled by (row & 7): conflict-free

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OFF_STAGING = STAGES * STAGE_BYTES;
  static constexpr int STAGING_BYTES = EPI_WARPS * STG_WARP_BYTES;
  static constexpr int OFF_BIAS = OFF_STAGING + STAGING_BYTES;
  static constexpr int OFF_BARS = OFF_BIAS + BN * 4;
  static constexpr int TOTAL = OFF_BARS + 128 + 1024;  // + barrier block + 1024 B alignment slack
};

struct Bars {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

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

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const GemmArgs g) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars* bars = reinterpret_cast<Bars*>(smem + L::OFF_BARS);
  float* sbias = reinterpret_cast<float*>(smem + L::OFF_BIAS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blks = (g.M + BM - 1) / BM;
  const int n_blks = (g.N + BN - 1) / BN;
  const int k_blks = (g.K + BK - 1) / BK;
  const int num_tiles = m_blks * n_blks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&bars->full[i], 1);
        mbar_init(&bars->empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars->tmem_full[i], 1);
        mbar_init(&bars->tmem_empty[i], EPI_WARPS);  // one arrival per epilogue warp
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_base, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_blks) * BM;
        const int n0 = (tile % n_blks) * BN;
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * L::STAGE_BYTES;
          uint8_t* sB = sA + L::A_BYTES;
          mbar_arrive_expect_tx(&bars->full[stage], L::STAGE_BYTES);
          tma_load_2d(sA, &tmA, &bars->full[stage], kb * BK, m0);
          tma_load_2d(sB, &tmW, &bars->full[stage], kb * BK, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blks; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_desc_kmajor_sw128(a_addr + k * 32);
            const uint64_t db = umma_desc_kmajor_sw128(b_addr + k * 32);
            umma_f16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bars->empty[stage]);  // frees this smem stage once the MMAs above retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&bars->tmem_full[as]);  // accumulator complete -> epilogue
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM lane quarter (a warp may only read lanes 32*(warp%4)..+31); the pair splits the
    // tile's columns in halves, so every SM sub-partition always has two epilogue warps to interleave.
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int ep_tid = threadIdx.x - 64;
    uint8_t* stg = smem + L::OFF_STAGING + (warp - 2) * STG_WARP_BYTES;
    constexpr int GRP_COLS = (EPI == EPI_F32) ? 32 : 64;  // columns per 128-byte staging row
    constexpr int GRPS_PER_HALF = BN / GRP_COLS / 2;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_blks) * BM;
      const int n0 = (tile % n_blks) * BN;
      named_bar_sync(1, EPI_WARPS * 32);  // previous tile's bias fully consumed
      for (int i = ep_tid; i < BN; i += EPI_WARPS * 32) {
        const int c = n0 + i;
        sbias[i] = (g.bias != nullptr && c < g.N) ? __half2float(g.bias[c]) : 0.0f;
      }
      named_bar_sync(1, EPI_WARPS * 32);
      const int n_grps = (min(g.N - n0, BN) + GRP_COLS - 1) / GRP_COLS;  // groups with at least one valid column
      const int g_begin = chalf * GRPS_PER_HALF;
      const int g_end = min(g_begin + GRPS_PER_HALF, n_grps);
      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      if (g_begin >= g_end) {  // ragged N: nothing to read for this warp, still release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[as]);
      }
      for (int grp = g_begin; grp < g_end; ++grp) {
        // residual rows for pass 2 are requested first so their latency hides behind pass 1
        // (C may alias the residual: every 16-byte chunk is read and written by the same thread).
        uint4 rs[8];
        if (EPI == EPI_BIAS_RES) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int idx = it * 32 + lane;
            const int grow = m0 + q * 32 + (idx >> 3);
            const int gcol = n0 + grp * 64 + (idx & 7) * 8;
            rs[it] = make_uint4(0, 0, 0, 0);
            if (grow < g.M && gcol < g.N)
              rs[it] = *reinterpret_cast<const uint4*>(g.residual + static_cast<size_t>(grow) * g.ldr + gcol);
          }
        }
        // pass 1: TMEM -> registers -> (bias, activation, round) -> swizzled staging row `lane`
        if (EPI == EPI_F32) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + grp * 32, v);
          tmem_wait_ld();
          const float4* b4 = reinterpret_cast<const float4*>(sbias + grp * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j];
            float4 o;
            o.x = __uint_as_float(v[4 * j + 0]) + b.x;
            o.y = __uint_as_float(v[4 * j + 1]) + b.y;
            o.z = __uint_as_float(v[4 * j + 2]) + b.z;
            o.w = __uint_as_float(v[4 * j + 3]) + b.w;
            *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
          }
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            tmem_ld_32x32(t_row + grp * 64 + h * 32, v);
            tmem_wait_ld();
            const float4* b4 = reinterpret_cast<const float4*>(sbias + grp * 64 + h * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float4 b = b4[2 * j + e];
                pk[2 * e + 0] = pack_half2(__uint_as_float(v[8 * j + 4 * e + 0]) + b.x,
                                           __uint_as_float(v[8 * j + 4 * e + 1]) + b.y);
                pk[2 * e + 1] = pack_half2(__uint_as_float(v[8 * j + 4 * e + 2]) + b.z,
                                           __uint_as_float(v[8 * j + 4 * e + 3]) + b.w);
                if (EPI == EPI_BIAS_QGELU) {
                  pk[2 * e + 0] = quick_gelu_f16x2(pk[2 * e + 0]);
                  pk[2 * e + 1] = quick_gelu_f16x2(pk[2 * e + 1]);
                }
              }
              *reinterpret_cast<uint4*>(stg + lane * 128 + (((h * 4 + j) ^ (lane & 7)) << 4)) =
                  make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
        if (grp == g_end - 1) {
          // last TMEM read of this accumulator stage by this warp: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tmem_empty[as]);
        }
        __syncwarp();
        // pass 2: staging -> global, 8 lanes per 128-byte row (full-line coalesced stores)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int idx = it * 32 + lane;
          const int r = idx >> 3;
          const int ch = idx & 7;
          const int grow = m0 + q * 32 + r;
          uint4 val = *reinterpret_cast<const uint4*>(stg + r * 128 + ((ch ^ (r & 7)) << 4));
          if (EPI == EPI_F32) {
            const int gcol = n0 + grp * 32 + ch * 4;
            if (grow < g.M && gcol < g.N) {
              float* dst = reinterpret_cast<float*>(g.C) + static_cast<size_t>(grow) * g.ldc + gcol;
              *reinterpret_cast<uint4*>(dst) = val;
            }
          } else {
            const int gcol = n0 + grp * 64 + ch * 8;
            if (grow < g.M && gcol < g.N) {
              if (EPI == EPI_BIAS_RES) {
                const __half2* a2 = reinterpret_cast<const __half2*>(&val);
                const __half2* r2 = reinterpret_cast<const __half2*>(&rs[it]);
                uint4 o;
                __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) o2[e] = __hadd2(r2[e], a2[e]);
                val = o;
              }
              __half* dst = reinterpret_cast<__half*>(g.C) + static_cast<size_t>(grow) * g.ldc + gcol;
              *reinterpret_cast<uint4*>(dst) = val;
            }
          }
        }
        __syncwarp();
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, int EPI>
int launch_impl(const GemmArgs& a, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  static bool configured = false;
  auto kern = gemm_tn_kernel<BN, EPI>;
  if (!configured) {
    PC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  CUtensorMap tmA, tmW;
  PC_TRY(make_tmap_f16_2d(&tmA, a.A, a.K, a.M, static_cast<uint64_t>(a.lda) * 2, BK, BM));
  PC_TRY(make_tmap_f16_2d(&tmW, a.W, a.K, a.N, static_cast<uint64_t>(a.ldw) * 2, BK, BN));
  const int m_blks = (a.M + BM - 1) / BM;
  const int n_blks = (a.N + BN - 1) / BN;
  const int tiles = m_blks * n_blks;
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  kern<<<grid, GEMM_THREADS, L::TOTAL, stream>>>(tmA, tmW, a);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

template <int BN>
int dispatch_epi(const GemmArgs& a, int epi, cudaStream_t stream) {
  switch (epi) {
    case EPI_BIAS: return launch_impl<BN, EPI_BIAS>(a, stream);
    case EPI_BIAS_QGELU: return launch_impl<BN, EPI_BIAS_QGELU>(a, stream);
    case EPI_BIAS_RES: return launch_impl<BN, EPI_BIAS_RES>(a, stream);
    case EPI_F32: return launch_impl<BN, EPI_F32>(a, stream);
    default: set_error("unknown GEMM epilogue %d", epi); return PC_ERR_ARG;
  }
}

}  // namespace

int launch_gemm(const GemmArgs& a, int epilogue, cudaStream_t stream) {
  PC_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, PC_ERR_ARG, "gemm: empty problem %dx%dx%d", a.M, a.N, a.K);
  PC_REQUIRE(a.A && a.W && a.C, PC_ERR_ARG, "gemm: null operand");
  PC_REQUIRE(a.K % 8 == 0 && a.lda % 8 == 0 && a.ldw % 8 == 0, PC_ERR_ALIGN,
             "gemm: K/lda/ldw (%d/%d/%d) must be multiples of 8 fp16 (16-byte TMA rows)", a.K, a.lda, a.ldw);
  // Columns are written in 16-byte chunks; a ragged last chunk spills into the row padding, so the leading
  // dimension must cover N rounded up to the chunk.
  const int chunk = (epilogue == EPI_F32) ? 4 : 8;
  PC_REQUIRE(a.ldc % chunk == 0 && a.ldc >= (a.N + chunk - 1) / chunk * chunk &&
                 (reinterpret_cast<uintptr_t>(a.C) & 15) == 0,
             PC_ERR_ALIGN, "gemm: output needs a 16-byte aligned C and ldc (%d) a multiple of %d covering N (%d)",
             a.ldc, chunk, a.N);
  if (epilogue == EPI_BIAS_RES) {
    PC_REQUIRE(a.residual != nullptr && a.ldr % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(a.residual) & 15) == 0,
               PC_ERR_ALIGN, "gemm: residual must be non-null, 16-byte aligned, ldr %% 8 == 0");
  }
  if (a.N > 128) return dispatch_epi<256>(a, epilogue, stream);
  return dispatch_epi<128>(a, epilogue, stream);
}

}  // namespace pc
