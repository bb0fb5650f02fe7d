// HBM-bound row kernels around the tensor-core GEMMs: one warp per row, 16-byte vector loads, fp32
// statistics with warp-shuffle reductions, fp16 storage. They restate
//   LayerNorm (fp32 compute on fp16 storage)           clip/model.py:155-161
//   patch extraction for conv1 (k = s = p, no bias)    clip/model.py:209,222-224
//   CLS concat + positional embedding + ln_pre         clip/model.py:225-227
//   token embedding + positional embedding             clip/model.py:342-344
//   EOT gather (text.argmax(-1))                       clip/model.py:352
//   feature L2 normalisation                           utils.py:352
#include <stdlib.h>

#include "kernels.cuh"
#include "ptx.cuh"

namespace pc {

namespace {

constexpr int ROW_WARPS = 8;   // warps (rows) per CTA
constexpr int MAX_VEC = 8;     // 16-byte vectors per lane: d <= 32 * 8 * 8 = 2048

template <int NV>
struct RowF {  // one row slice held by a lane: NV 16-byte vectors = NV * 8 floats (NV = ceil(d / 256))
  float v[NV][8];
};

template <int NV>
__device__ __forceinline__ void load_row_f16(const __half* row, int d, int lane, RowF<NV>& r, int& nvec) {
  const int vecs = d >> 3;
  nvec = 0;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = lane + k * 32;
    if (vi < vecs) {
      const uint4 u = *reinterpret_cast<const uint4*>(row + vi * 8);
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        r.v[k][2 * e] = f.x;
        r.v[k][2 * e + 1] = f.y;
      }
      nvec = k + 1;
    }
  }
}

// Normalise the row held in `r` (fp32 two-pass mean / biased variance, eps 1e-5) and store fp16.
// stats (nullable): (sum, sum of squares) of the fp16 values STORED, i.e. what launch_row_stats would compute on `out`.
template <int NV>
__device__ __forceinline__ void ln_store(RowF<NV>& r, int d, int lane, const float* gamma, const float* beta,
                                         __half* out, float* stats = nullptr) {
  const int vecs = d >> 3;
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (lane + k * 32 < vecs)
#pragma unroll
      for (int e = 0; e < 8; ++e) s += r.v[k][e];
  const float mean = warp_sum(s) / static_cast<float>(d);
  float q = 0.0f;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (lane + k * 32 < vecs)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float c = r.v[k][e] - mean;
        q += c * c;
      }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(d) + 1e-5f);
  float st_s = 0.0f, st_q = 0.0f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = lane + k * 32;
    if (vi < vecs) {
      uint32_t pk[4];
      float gm[8], bt[8];
      *reinterpret_cast<float4*>(gm) = *reinterpret_cast<const float4*>(gamma + vi * 8);
      *reinterpret_cast<float4*>(gm + 4) = *reinterpret_cast<const float4*>(gamma + vi * 8 + 4);
      *reinterpret_cast<float4*>(bt) = *reinterpret_cast<const float4*>(beta + vi * 8);
      *reinterpret_cast<float4*>(bt + 4) = *reinterpret_cast<const float4*>(beta + vi * 8 + 4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        pk[e] = pack_half2((r.v[k][2 * e] - mean) * rstd * gm[2 * e] + bt[2 * e],
                           (r.v[k][2 * e + 1] - mean) * rstd * gm[2 * e + 1] + bt[2 * e + 1]);
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pk[e]));
        st_s += f.x + f.y;
        st_q = fmaf(f.x, f.x, fmaf(f.y, f.y, st_q));
      }
      *reinterpret_cast<uint4*>(out + vi * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
  if (stats != nullptr) {
    st_s = warp_sum(st_s);
    st_q = warp_sum(st_q);
    if (lane == 0) *reinterpret_cast<float2*>(stats) = make_float2(st_s, st_q);
  }
}

// Persistent LayerNorm: each warp walks rows with a grid stride, keeps its gamma / beta slice in registers and
// has the next row's 16-byte vectors in flight while it reduces and stores the current one (the activations are
// L2-resident between two GEMMs, so the kernel is bound by load latency, not bandwidth).
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const int* __restrict__ rows_idx, int rows, int d,
                 int row_stride_rows) {
  const int lane = threadIdx.x & 31;
  const int vecs = d >> 3;
  const int nwarps = gridDim.x * ROW_WARPS;
  int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  griddep_launch_dependents();
  // gamma | beta (weights, not produced by a kernel of this stream) staged once per CTA in shared memory
  extern __shared__ float4 ln_params4[];
  float* sg = reinterpret_cast<float*>(ln_params4);
  float* sb = sg + d;
  for (int i = threadIdx.x; i < (d >> 2); i += blockDim.x) {
    reinterpret_cast<float4*>(sg)[i] = reinterpret_cast<const float4*>(gamma)[i];
    reinterpret_cast<float4*>(sb)[i] = reinterpret_cast<const float4*>(beta)[i];
  }
  __syncthreads();
  griddep_wait();  // x is the previous kernel's output
  if (row >= rows) return;
  uint4 cur[NV], nxt[NV];
  auto src_of = [&](int r) -> const __half* {
    const size_t src = rows_idx ? static_cast<size_t>(rows_idx[r]) : static_cast<size_t>(r) * row_stride_rows;
    return x + src * d;
  };
  {
    const __half* xr = src_of(row);
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (lane + k * 32 < vecs) cur[k] = *reinterpret_cast<const uint4*>(xr + (lane + k * 32) * 8);
  }
  const float inv_d = 1.0f / static_cast<float>(d);
  while (row < rows) {
    const int next = row + nwarps;
    if (next < rows) {
      const __half* xr = src_of(next);
#pragma unroll
      for (int k = 0; k < NV; ++k)
        if (lane + k * 32 < vecs) nxt[k] = *reinterpret_cast<const uint4*>(xr + (lane + k * 32) * 8);
    }
    float v[NV][8];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (lane + k * 32 < vecs) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&cur[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          v[k][2 * e] = f.x;
          v[k][2 * e + 1] = f.y;
          s += f.x + f.y;
        }
      }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (lane + k * 32 < vecs)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v[k][e] -= mean;
          q += v[k][e] * v[k][e];
        }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + 1e-5f);
    __half* out = y + static_cast<size_t>(row) * d;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int vi = lane + k * 32;
      if (vi < vecs) {
        uint32_t pk[4];
        float gm[8], bt[8];
        *reinterpret_cast<float4*>(gm) = *reinterpret_cast<const float4*>(sg + vi * 8);
        *reinterpret_cast<float4*>(gm + 4) = *reinterpret_cast<const float4*>(sg + vi * 8 + 4);
        *reinterpret_cast<float4*>(bt) = *reinterpret_cast<const float4*>(sb + vi * 8);
        *reinterpret_cast<float4*>(bt + 4) = *reinterpret_cast<const float4*>(sb + vi * 8 + 4);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          pk[e] = pack_half2(v[k][2 * e] * rstd * gm[2 * e] + bt[2 * e], v[k][2 * e + 1] * rstd * gm[2 * e + 1] + bt[2 * e + 1]);
        *reinterpret_cast<uint4*>(out + vi * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) cur[k] = nxt[k];
    row = next;
  }
}

// stats[r] = (sum_k x[r,k], sum_k x[r,k]^2) in fp32: the LayerNorm statistics the GEMM's LN-folded epilogues read.
// Inside a tower the residual GEMMs maintain them (GemmArgs::stats_out); this kernel seeds them for a block input.
__global__ void __launch_bounds__(ROW_WARPS * 32)
row_stats_kernel(const __half* __restrict__ x, float* __restrict__ stats, int rows, int d) {
  griddep_launch_dependents();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int vecs = d >> 3;
  for (int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5); row < rows; row += gridDim.x * ROW_WARPS) {
    const __half* xr = x + static_cast<size_t>(row) * d;
    float s = 0.0f, q = 0.0f;
    for (int vi = lane; vi < vecs; vi += 32) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + vi * 8);
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        s += f.x + f.y;
        q = fmaf(f.x, f.x, fmaf(f.y, f.y, q));
      }
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) *reinterpret_cast<float2*>(stats + 2 * static_cast<size_t>(row)) = make_float2(s, q);
  }
}

// LayerNorm folding of one Linear (bind time): Wf[n,k] = f16(W[n,k] * gamma[k]); s[n] = sum_k f32(Wf[n,k]);
// c[n] = sum_k beta[k] * f32(W[n,k]) + bias[n]. One warp per output row.
__global__ void __launch_bounds__(256)
fold_ln_kernel(const __half* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
               const __half* __restrict__ bias, __half* __restrict__ Wf, float* __restrict__ s_out,
               float* __restrict__ c_out, int N, int K) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  float s = 0.0f, c = 0.0f;
  for (int k = lane; k < K; k += 32) {
    const float w = __half2float(W[static_cast<size_t>(n) * K + k]);
    const __half wf = __float2half_rn(w * gamma[k]);
    Wf[static_cast<size_t>(n) * K + k] = wf;
    s += __half2float(wf);
    c = fmaf(beta[k], w, c);
  }
  s = warp_sum(s);
  c = warp_sum(c);
  if (lane == 0) {
    s_out[n] = s;
    c_out[n] = c + (bias ? __half2float(bias[n]) : 0.0f);
  }
}

// One warp per output row (image b, patch gy, gx): its 3 p segments (channel c, patch row i) of p contiguous input pixels
// become p contiguous fp16 values at column (c p + i) p; lanes take consecutive segments, so the warp writes the whole
// Kp-wide row contiguously and reads 3 p full 4 p-byte (p = 16: two sectors) pieces of image rows. (The first version,
// one thread per segment with consecutive lanes on consecutive PATCHES, wrote 32-byte pieces 1.5 KB apart: 35 % of the
// HBM copy rate in ncu; see profiles/.) Zero-padding columns [K, Kp) are written by the first lanes.
// PT: the patch size when it is one of CLIP's (16, 32, 14), so that the per-segment loops unroll and a lane has all of
// its loads in flight at once; 0 = any even p.
template <int PT>
__global__ void __launch_bounds__(256)
patchify_kernel(const void* __restrict__ images, int img_is_f16, __half* __restrict__ out, int B, int R, int p_rt,
                int g, int K, int Kp) {
  const int p = PT > 0 ? PT : p_rt;
  const int lane = threadIdx.x & 31;
  const int rows = B * g * g, segs = 3 * p;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8) {
    const int gx = row % g, gy = (row / g) % g, b = row / (g * g);
    __half* dst_row = out + static_cast<size_t>(row) * Kp;
    for (int sg = lane; sg < segs; sg += 32) {
      const int c = sg / p, i = sg - c * p;
      const size_t src = ((static_cast<size_t>(b) * 3 + c) * R + (gy * p + i)) * R + gx * p;
      __half2* dst = reinterpret_cast<__half2*>(dst_row + sg * p);
      if (img_is_f16) {
        const __half2* in2 = reinterpret_cast<const __half2*>(static_cast<const __half*>(images) + src);
#pragma unroll
        for (int j = 0; j < (p >> 1); ++j) dst[j] = in2[j];
      } else if ((p & 3) == 0) {  // 16-byte loads (src is a multiple of p floats)
        const float4* in4 = reinterpret_cast<const float4*>(static_cast<const float*>(images) + src);
        if (PT > 0) {
          float4 f[(PT > 0 ? PT : 4) / 4];
#pragma unroll
          for (int j = 0; j < PT / 4; ++j) f[j] = __ldg(in4 + j);
#pragma unroll
          for (int j = 0; j < PT / 8; ++j) {  // 8 pixels -> one 16-byte store
            uint4 o;
            o.x = pack_half2(f[2 * j].x, f[2 * j].y);
            o.y = pack_half2(f[2 * j].z, f[2 * j].w);
            o.z = pack_half2(f[2 * j + 1].x, f[2 * j + 1].y);
            o.w = pack_half2(f[2 * j + 1].z, f[2 * j + 1].w);
            reinterpret_cast<uint4*>(dst)[j] = o;
          }
        } else {
          for (int j = 0; j < (p >> 2); ++j) {
            const float4 f = in4[j];
            dst[2 * j] = __floats2half2_rn(f.x, f.y);
            dst[2 * j + 1] = __floats2half2_rn(f.z, f.w);
          }
        }
      } else {
        const float2* in2 = reinterpret_cast<const float2*>(static_cast<const float*>(images) + src);
#pragma unroll
        for (int j = 0; j < (p >> 1); ++j) {
          const float2 f = __ldg(in2 + j);
          dst[j] = __floats2half2_rn(f.x, f.y);
        }
      }
    }
    for (int k = K + lane; k < Kp; k += 32) dst_row[k] = __float2half_rn(0.0f);
  }
}

// x[b, 0] = cls + pos[0]; x[b, 1 + t] = patch[b, t] + pos[1 + t] (fp16 adds of fp16-cast terms, clip/model.py:225-226),
// then ln_pre (:227). 16-byte vector loads; also emits the (sum, sum of squares) of the stored fp16 row: the
// LayerNorm statistics the first block's folded QKV GEMM reads (saves the row_stats pass over x).
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
embed_ln_pre_kernel(const __half* __restrict__ patch, const float* __restrict__ cls,
                    const float* __restrict__ pos, const float* __restrict__ gamma,
                    const float* __restrict__ beta, __half* __restrict__ x, float* __restrict__ stats, int B, int L,
                    int d) {
  const int lane = threadIdx.x & 31;
  const int vecs = d >> 3;
  for (int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5); row < B * L; row += gridDim.x * ROW_WARPS) {
    const int b = row / L, t = row % L;
    const __half* prow = patch + (static_cast<size_t>(b) * (L - 1) + (t - 1)) * d;  // t >= 1
    RowF<NV> r;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int vi = lane + k * 32;
      if (vi < vecs) {
        float pe[8], tk[8];
        *reinterpret_cast<float4*>(pe) = *reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * d + vi * 8);
        *reinterpret_cast<float4*>(pe + 4) = *reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * d + vi * 8 + 4);
        __half2 tok[4];
        if (t == 0) {
          *reinterpret_cast<float4*>(tk) = *reinterpret_cast<const float4*>(cls + vi * 8);
          *reinterpret_cast<float4*>(tk + 4) = *reinterpret_cast<const float4*>(cls + vi * 8 + 4);
#pragma unroll
          for (int e = 0; e < 4; ++e) tok[e] = __floats2half2_rn(tk[2 * e], tk[2 * e + 1]);
        } else {
          *reinterpret_cast<uint4*>(tok) = *reinterpret_cast<const uint4*>(prow + vi * 8);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(__hadd2(tok[e], __floats2half2_rn(pe[2 * e], pe[2 * e + 1])));
          r.v[k][2 * e] = f.x;
          r.v[k][2 * e + 1] = f.y;
        }
      }
    }
    ln_store(r, d, lane, gamma, beta, x + static_cast<size_t>(row) * d,
             stats != nullptr ? stats + 2 * static_cast<size_t>(row) : nullptr);
  }
}

__global__ void __launch_bounds__(256)
text_embed_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ tok_emb,
                  const float* __restrict__ pos, __half* __restrict__ x, int rows, int L, int d, int vocab) {
  const int pairs = d >> 1;
  const size_t total = static_cast<size_t>(rows) * pairs;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(t % pairs) * 2;
    const size_t row = t / pairs;
    long long tok = tokens[row];
    tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
    const float2 e = *reinterpret_cast<const float2*>(tok_emb + static_cast<size_t>(tok) * d + c);
    const float2 pe = *reinterpret_cast<const float2*>(pos + static_cast<size_t>(row % L) * d + c);
    const __half2 s = __hadd2(__floats2half2_rn(e.x, e.y), __floats2half2_rn(pe.x, pe.y));
    *reinterpret_cast<__half2*>(x + row * d + c) = s;
  }
}

__global__ void eot_index_kernel(const int64_t* __restrict__ tokens, int* __restrict__ rows, int P, int L) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  long long best = tokens[static_cast<size_t>(p) * L];
  int arg = 0;
  for (int t = 1; t < L; ++t) {
    const long long v = tokens[static_cast<size_t>(p) * L + t];
    if (v > best) {
      best = v;
      arg = t;
    }
  }
  rows[p] = p * L + arg;
}

template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
l2norm_kernel(const __half* __restrict__ x, __half* __restrict__ y, int rows, int d) {
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  RowF<NV> r;
  int nvec;
  load_row_f16(x + static_cast<size_t>(row) * d, d, lane, r, nvec);
  const int vecs = d >> 3;
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if (lane + k * 32 < vecs)
#pragma unroll
      for (int e = 0; e < 8; ++e) s += r.v[k][e] * r.v[k][e];
  // x.norm(dim=-1) on an fp16 tensor: fp32 accumulation, fp16 result; then an fp16 division.
  const float n = __half2float(__float2half_rn(sqrtf(warp_sum(s))));
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = lane + k * 32;
    if (vi < vecs) {
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) pk[e] = pack_half2(r.v[k][2 * e] / n, r.v[k][2 * e + 1] / n);
      *reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * d + vi * 8) =
          make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// NV = ceil(d / 256): instantiate 1, 2, 3, 4 and 8 so the common widths (<= 1024) keep registers low.
#define PC_DISPATCH_NV(d, CALL)              \
  do {                                       \
    const int _nv = ((d) + 255) / 256;       \
    if (_nv <= 1) { CALL(1); }               \
    else if (_nv == 2) { CALL(2); }          \
    else if (_nv == 3) { CALL(3); }          \
    else if (_nv == 4) { CALL(4); }          \
    else { CALL(8); }                        \
  } while (0)

int check_row_dims(const char* what, int rows, int d) {
  PC_REQUIRE(rows > 0 && d > 0, PC_ERR_ARG, "%s: empty input (%d x %d)", what, rows, d);
  PC_REQUIRE(d % 8 == 0 && d <= 32 * 8 * MAX_VEC, PC_ERR_ARG, "%s: width %d must be a multiple of 8 and <= %d",
             what, d, 32 * 8 * MAX_VEC);
  return PC_OK;
}

int grid_1d(size_t work, int block) {
  size_t g = (work + block - 1) / block;
  const size_t cap = static_cast<size_t>(device_sm_count()) * 16;
  return static_cast<int>(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

// persistent grid: up to 4 CTAs (32 warps) per SM, fewer when there are not enough rows
static int ln_grid(int rows) {
  const int want = (rows + ROW_WARPS - 1) / ROW_WARPS;
  const int cap = device_sm_count() * 4;
  return want < cap ? want : cap;
}

int launch_layernorm(const __half* x, __half* y, const float* gamma, const float* beta, int rows, int d,
                     int row_stride_rows, cudaStream_t stream) {
  PC_TRY(check_row_dims("layernorm", rows, d));
  const int grid = ln_grid(rows);
#define CALL(NV) PC_CHECK_CUDA(launch_pdl(layernorm_kernel<NV>, dim3(grid), dim3(ROW_WARPS * 32), 2 * d * sizeof(float), stream, 1, \
                                          x, y, gamma, beta, static_cast<const int*>(nullptr), rows, d, row_stride_rows))
  PC_DISPATCH_NV(d, CALL);
#undef CALL
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_row_stats(const __half* x, float* stats, int rows, int d, cudaStream_t stream) {
  PC_TRY(check_row_dims("row_stats", rows, d));
  PC_REQUIRE(x && stats, PC_ERR_ARG, "row_stats: null buffer");
  PC_CHECK_CUDA(launch_pdl(row_stats_kernel, dim3(ln_grid(rows)), dim3(ROW_WARPS * 32), 0, stream, 1, x, stats, rows, d));
  return PC_OK;
}

int launch_fold_ln(const __half* W, const float* gamma, const float* beta, const __half* bias, __half* Wf, float* s,
                   float* c, int N, int K, cudaStream_t stream) {
  PC_REQUIRE(W && gamma && beta && Wf && s && c && N > 0 && K > 0, PC_ERR_ARG, "fold_ln: bad arguments");
  fold_ln_kernel<<<(N + 7) / 8, 256, 0, stream>>>(W, gamma, beta, bias, Wf, s, c, N, K);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_layernorm_gather(const __half* x, const int* rows_idx, __half* y, const float* gamma,
                            const float* beta, int n, int d, cudaStream_t stream) {
  PC_TRY(check_row_dims("layernorm_gather", n, d));
  const int grid = ln_grid(n);
#define CALL(NV) PC_CHECK_CUDA(launch_pdl(layernorm_kernel<NV>, dim3(grid), dim3(ROW_WARPS * 32), 2 * d * sizeof(float), stream, 1, \
                                          x, y, gamma, beta, rows_idx, n, d, 1))
  PC_DISPATCH_NV(d, CALL);
#undef CALL
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_patchify(const void* images, int img_is_f16, __half* out, int B, int R, int p, int Kp,
                    cudaStream_t stream) {
  PC_REQUIRE(images && out && B > 0 && p > 0 && R % p == 0 && p % 2 == 0 && R % 2 == 0, PC_ERR_ARG,
             "patchify: bad geometry B=%d R=%d p=%d", B, R, p);
  const int g = R / p;
  const int K = 3 * p * p;
  PC_REQUIRE(Kp >= K && Kp % 8 == 0, PC_ERR_ARG, "patchify: padded K %d < %d or not a multiple of 8", Kp, K);
  // dst rows start on 16-byte boundaries (Kp % 8 == 0); a segment's 16-byte stores also need p % 8 == 0
  const int grid = ln_grid(B * g * g);
  if (p == 16) patchify_kernel<16><<<grid, 256, 0, stream>>>(images, img_is_f16, out, B, R, p, g, K, Kp);
  else if (p == 32) patchify_kernel<32><<<grid, 256, 0, stream>>>(images, img_is_f16, out, B, R, p, g, K, Kp);
  else if (p == 14) patchify_kernel<14><<<grid, 256, 0, stream>>>(images, img_is_f16, out, B, R, p, g, K, Kp);
  else patchify_kernel<0><<<grid, 256, 0, stream>>>(images, img_is_f16, out, B, R, p, g, K, Kp);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_embed_ln_pre(const __half* patch, const float* cls, const float* pos, const float* gamma,
                        const float* beta, __half* x, float* stats, int B, int L, int d, cudaStream_t stream) {
  PC_TRY(check_row_dims("embed_ln_pre", B * L, d));
#define CALL(NV) embed_ln_pre_kernel<NV><<<ln_grid(B * L), ROW_WARPS * 32, 0, stream>>>( \
      patch, cls, pos, gamma, beta, x, stats, B, L, d)
  PC_DISPATCH_NV(d, CALL);
#undef CALL
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_text_embed(const int64_t* tokens, const float* tok_emb, const float* pos, __half* x, int P, int L,
                      int d, int vocab, cudaStream_t stream) {
  PC_REQUIRE(tokens && tok_emb && pos && x && P > 0 && L > 0 && d % 2 == 0, PC_ERR_ARG, "text_embed: bad args");
  const size_t total = static_cast<size_t>(P) * L * (d / 2);
  text_embed_kernel<<<grid_1d(total, 256), 256, 0, stream>>>(tokens, tok_emb, pos, x, P * L, L, d, vocab);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_eot_index(const int64_t* tokens, int* rows, int P, int L, cudaStream_t stream) {
  PC_REQUIRE(tokens && rows && P > 0 && L > 0, PC_ERR_ARG, "eot_index: bad args");
  eot_index_kernel<<<(P + 127) / 128, 128, 0, stream>>>(tokens, rows, P, L);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_l2norm(const __half* x, __half* y, int rows, int d, cudaStream_t stream) {
  PC_TRY(check_row_dims("l2norm", rows, d));
#define CALL(NV) l2norm_kernel<NV><<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, stream>>>(x, y, rows, d)
  PC_DISPATCH_NV(d, CALL);
#undef CALL
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

}  // namespace pc
