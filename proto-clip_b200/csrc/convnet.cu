// ModifiedResNet support kernels (reference clip/model.py:10-152). Activations are NHWC fp16, i.e. pixel-major
// [B*H*W, C] matrices, so that every 1x1 convolution IS a TN GEMM on gemm.cu's tcgen05 kernel and every 3x3
// convolution is an implicit GEMM over a zero-bordered copy (GemmArgs::conv_taps: nine row-shifted TMA views of the
// same tensor, no im2col matrix); eval-mode BatchNorm is folded into the fp16 weights and an
// fp32 per-channel shift at bind time, ReLU and the identity add run in the GEMM epilogue. What is left for this
// file is HBM-bound data movement: 16-byte vector loads / stores, one 8-channel vector per thread, grid-stride.
#include "kernels.cuh"
#include "ptx.cuh"

namespace pc {
namespace {

constexpr int CT = 256;  // threads per block

inline int grid_for(size_t work) {
  const size_t blocks = (work + CT - 1) / CT;
  const size_t cap = static_cast<size_t>(device_sm_count()) * 16;
  return static_cast<int>(blocks < cap ? (blocks ? blocks : 1) : cap);
}

// w [Cout, Cin, k, k] fp16 (nn.Conv2d layout) + BatchNorm2d(weight, bias, running_mean, running_var; eps 1e-5) ->
// wf [Cout, Kp] fp16 with column = (ky * k + kx) * Cin + ci (the im2col order below), zero-padded to Kp, scaled by
// gamma / sqrt(var + eps); shift[co] = beta - mean * scale (fp32).
__global__ void fold_conv_bn_kernel(const __half* __restrict__ w, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, __half* __restrict__ wf, float* __restrict__ shift,
                                    int Cout, int Cin, int k, int Kp) {
  const size_t total = static_cast<size_t>(Cout) * Kp;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(t % Kp);
    const int co = static_cast<int>(t / Kp);
    const float scale = gamma[co] * rsqrtf(var[co] + 1e-5f);
    float v = 0.0f;
    if (col < Cin * k * k) {
      const int ci = col % Cin, tap = col / Cin;
      v = __half2float(w[(static_cast<size_t>(co) * Cin + ci) * k * k + tap]) * scale;
    }
    wf[t] = __float2half_rn(v);
    if (col == 0) shift[co] = beta[co] - mean[co] * scale;
  }
}

// Stem conv1 (3 -> w/2, 3x3, stride 2, padding 1; clip/model.py:109) as a GEMM operand: images [B,3,R,R] (fp32 or
// fp16, NCHW) -> rows [B*Ho*Wo, 32] fp16, column = (ky*3 + kx)*3 + c, columns 27..31 zero. One thread per output
// pixel; neighbouring threads read neighbouring input columns.
__global__ void __launch_bounds__(CT)
stem_im2col_kernel(const void* __restrict__ images, int img_is_f16, __half* __restrict__ out, int B, int R, int Ho) {
  const size_t total = static_cast<size_t>(B) * Ho * Ho;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(t % Ho);
    const int oy = static_cast<int>((t / Ho) % Ho);
    const int b = static_cast<int>(t / (static_cast<size_t>(Ho) * Ho));
    __align__(16) __half v[32];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
        const bool in = iy >= 0 && iy < R && ix >= 0 && ix < R;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float f = 0.0f;
          if (in) {
            const size_t src = ((static_cast<size_t>(b) * 3 + c) * R + iy) * R + ix;
            f = img_is_f16 ? __half2float(static_cast<const __half*>(images)[src]) : static_cast<const float*>(images)[src];
          }
          v[(ky * 3 + kx) * 3 + c] = __float2half_rn(f);
        }
      }
    }
#pragma unroll
    for (int c = 27; c < 32; ++c) v[c] = __float2half_rn(0.0f);
    uint4* dst = reinterpret_cast<uint4*>(out + t * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = reinterpret_cast<const uint4*>(v)[q];
  }
}

// Zero-bordered copy for the implicit 3x3 convolution (GemmArgs::conv_taps): x [B,H,W,C] -> xp [B,H+2,W+2,C] with
// a one-pixel zero frame (the padding = 1 of nn.Conv2d, clip/model.py:20,111-113). One thread per 8-channel vector.
__global__ void __launch_bounds__(CT)
pad_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ xp, int B, int H, int W, int C) {
  const int cv = C >> 3, Hp = H + 2, Wp = W + 2;
  const size_t total = static_cast<size_t>(B) * Hp * Wp * cv;
  const uint4* xs = reinterpret_cast<const uint4*>(x);
  uint4* xd = reinterpret_cast<uint4*>(xp);
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(t % cv);
    size_t r = t / cv;
    const int px = static_cast<int>(r % Wp) - 1;
    r /= Wp;
    const int py = static_cast<int>(r % Hp) - 1;
    const size_t b = r / Hp;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (py >= 0 && py < H && px >= 0 && px < W) val = xs[((b * H + py) * W + px) * cv + v];
    xd[t] = val;
  }
}

// Inverse: the interior of a bordered tensor xp [B,H+2,W+2,C] -> x [B,H,W,C] (drops the garbage frame rows the
// implicit convolution produces).
__global__ void __launch_bounds__(CT)
unpad_nhwc_kernel(const __half* __restrict__ xp, __half* __restrict__ x, int B, int H, int W, int C) {
  const int cv = C >> 3, Hp = H + 2, Wp = W + 2;
  const size_t total = static_cast<size_t>(B) * H * W * cv;
  const uint4* xs = reinterpret_cast<const uint4*>(xp);
  uint4* xd = reinterpret_cast<uint4*>(x);
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(t % cv);
    size_t r = t / cv;
    const int px = static_cast<int>(r % W);
    r /= W;
    const int py = static_cast<int>(r % H);
    const size_t b = r / H;
    xd[t] = xs[((b * Hp + py + 1) * Wp + px + 1) * cv + v];
  }
}

// Re-zero the frame of a bordered tensor in place (between two chained implicit convolutions of the stem).
__global__ void __launch_bounds__(CT)
zero_border_kernel(__half* __restrict__ xp, int B, int H, int W, int C) {
  const int cv = C >> 3, Hp = H + 2, Wp = W + 2;
  const int frame = 2 * Wp + 2 * H;  // frame pixels per image: top + bottom rows, left + right columns of the rest
  const size_t total = static_cast<size_t>(B) * frame * cv;
  uint4* xd = reinterpret_cast<uint4*>(xp);
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(t % cv);
    size_t r = t / cv;
    const int f = static_cast<int>(r % frame);
    const size_t b = r / frame;
    int py, px;
    if (f < Wp) { py = 0; px = f; }
    else if (f < 2 * Wp) { py = Hp - 1; px = f - Wp; }
    else { const int e = f - 2 * Wp; py = 1 + (e >> 1); px = (e & 1) ? Wp - 1 : 0; }
    xd[((b * Hp + py) * Wp + px) * cv + v] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// nn.AvgPool2d(s) on NHWC fp16 (clip/model.py:23,35,115): fp32 accumulation, one rounding. pad = 1: the input is a
// bordered tensor [B,H+2,W+2,C] (output of an implicit convolution) whose interior is pooled.
__global__ void __launch_bounds__(CT)
avgpool_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H, int W, int C, int s, int pad) {
  const int cv = C >> 3, Ho = H / s, Wo = W / s;
  const int Hi = H + 2 * pad, Wi = W + 2 * pad;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * cv;
  const uint4* xs = reinterpret_cast<const uint4*>(x);
  const float inv = 1.0f / static_cast<float>(s * s);
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(t % cv);
    size_t r = t / cv;
    const int ox = static_cast<int>(r % Wo);
    r /= Wo;
    const int oy = static_cast<int>(r % Ho);
    const size_t b = r / Ho;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int dy = 0; dy < s; ++dy)
      for (int dx = 0; dx < s; ++dx) {
        const uint4 q = xs[((b * Hi + oy * s + dy + pad) * Wi + ox * s + dx + pad) * cv + v];
        const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          acc[2 * e] += f.x;
          acc[2 * e + 1] += f.y;
        }
      }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(acc[2 * e] * inv, acc[2 * e + 1] * inv);
    reinterpret_cast<uint4*>(y)[t] = o;
  }
}

// AttentionPool2d token assembly (clip/model.py:68-70): tok[b,0,:] = f16(mean_p x[b,p,:]) + f16(pos[0]),
// tok[b,1+p,:] = x[b,p,:] + f16(pos[1+p]) (fp16 adds). One thread per (image, 8-channel vector): the column walk
// over the HW pixels is coalesced across the threads of a warp.
__global__ void __launch_bounds__(CT)
attnpool_tokens_kernel(const __half* __restrict__ x, const float* __restrict__ pos, __half* __restrict__ tok, int B,
                       int HW, int C) {
  const int cv = C >> 3;
  const size_t total = static_cast<size_t>(B) * cv;
  const uint4* xs = reinterpret_cast<const uint4*>(x);
  uint4* ts = reinterpret_cast<uint4*>(tok);
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(t % cv);
    const size_t b = t / cv;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p <= HW; ++p) {
      // p == HW handles the mean token (row 0) once every pixel has been accumulated
      const float* pr = pos + static_cast<size_t>(p == HW ? 0 : p + 1) * C + v * 8;
      const float4 p0 = *reinterpret_cast<const float4*>(pr), p1 = *reinterpret_cast<const float4*>(pr + 4);
      const float pf[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      uint4 q;
      __half2* h = reinterpret_cast<__half2*>(&q);
      if (p < HW) {
        q = xs[(b * HW + p) * cv + v];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          acc[2 * e] += f.x;
          acc[2 * e + 1] += f.y;
        }
      } else {
        const float inv = 1.0f / static_cast<float>(HW);
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(acc[2 * e] * inv, acc[2 * e + 1] * inv);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __hadd2(h[e], __floats2half2_rn(pf[2 * e], pf[2 * e + 1]));
      ts[(b * (HW + 1) + (p == HW ? 0 : p + 1)) * cv + v] = q;
    }
  }
}

}  // namespace

int launch_fold_conv_bn(const __half* w, const float* gamma, const float* beta, const float* mean, const float* var,
                        __half* wf, float* shift, int Cout, int Cin, int k, int Kp, cudaStream_t stream) {
  PC_REQUIRE(w && gamma && beta && mean && var && wf && shift, PC_ERR_ARG, "fold_conv_bn: null tensor");
  PC_REQUIRE(Cout > 0 && Cin > 0 && (k == 1 || k == 3) && Kp >= Cin * k * k && Kp % 8 == 0, PC_ERR_ARG,
             "fold_conv_bn: Cout=%d Cin=%d k=%d Kp=%d", Cout, Cin, k, Kp);
  fold_conv_bn_kernel<<<grid_for(static_cast<size_t>(Cout) * Kp), CT, 0, stream>>>(w, gamma, beta, mean, var, wf, shift,
                                                                                   Cout, Cin, k, Kp);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_stem_im2col(const void* images, int img_is_f16, __half* out, int B, int R, cudaStream_t stream) {
  PC_REQUIRE(images && out && B > 0 && R > 0 && R % 2 == 0, PC_ERR_ARG, "stem_im2col: B=%d R=%d", B, R);
  stem_im2col_kernel<<<grid_for(static_cast<size_t>(B) * (R / 2) * (R / 2)), CT, 0, stream>>>(images, img_is_f16, out, B,
                                                                                              R, R / 2);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_pad_nhwc(const __half* x, __half* xp, int B, int H, int W, int C, cudaStream_t stream) {
  PC_REQUIRE(x && xp && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, PC_ERR_ARG,
             "pad_nhwc: B=%d H=%d W=%d C=%d (C must be a multiple of 8)", B, H, W, C);
  pad_nhwc_kernel<<<grid_for(static_cast<size_t>(B) * (H + 2) * (W + 2) * (C / 8)), CT, 0, stream>>>(x, xp, B, H, W, C);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_unpad_nhwc(const __half* xp, __half* x, int B, int H, int W, int C, cudaStream_t stream) {
  PC_REQUIRE(x && xp && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, PC_ERR_ARG,
             "unpad_nhwc: B=%d H=%d W=%d C=%d (C must be a multiple of 8)", B, H, W, C);
  unpad_nhwc_kernel<<<grid_for(static_cast<size_t>(B) * H * W * (C / 8)), CT, 0, stream>>>(xp, x, B, H, W, C);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_zero_border(__half* xp, int B, int H, int W, int C, cudaStream_t stream) {
  PC_REQUIRE(xp && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, PC_ERR_ARG, "zero_border: B=%d H=%d W=%d C=%d", B, H, W, C);
  zero_border_kernel<<<grid_for(static_cast<size_t>(B) * (2 * (W + 2) + 2 * H) * (C / 8)), CT, 0, stream>>>(xp, B, H, W, C);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_avgpool_nhwc(const __half* x, __half* y, int B, int H, int W, int C, int s, int in_bordered,
                        cudaStream_t stream) {
  PC_REQUIRE(x && y && B > 0 && s > 0 && H % s == 0 && W % s == 0 && C % 8 == 0, PC_ERR_ARG,
             "avgpool: B=%d H=%d W=%d C=%d s=%d", B, H, W, C, s);
  avgpool_nhwc_kernel<<<grid_for(static_cast<size_t>(B) * (H / s) * (W / s) * (C / 8)), CT, 0, stream>>>(
      x, y, B, H, W, C, s, in_bordered ? 1 : 0);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

// AttentionPool2d returns token 0 only (clip/model.py:72-92 computes all HW + 1 query rows and drops the rest): ONE
// query row per (image, head). One warp per (image, head): lanes take keys for the scores (64-wide dot products of
// 128-byte rows), the probabilities go through shared memory, then lanes take two head-dim columns each for P V.
// q: [B, E] (token 0's projected query, bias included), kv: [B*L, 2E] (k | v, biases included), out: [B, E] fp16.
constexpr int POOL_WARPS = 4;
__global__ void __launch_bounds__(POOL_WARPS * 32)
attnpool_query0_kernel(const __half* __restrict__ q, const __half* __restrict__ kv, __half* __restrict__ out, int B,
                       int L, int heads) {
  extern __shared__ float pool_smem[];  // [POOL_WARPS][64 + Lp]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int Lp = (L + 31) & ~31;
  float* qs = pool_smem + wib * (64 + Lp);
  float* ps = qs + 64;
  const int E = heads * 64;
  for (int item = blockIdx.x * POOL_WARPS + wib; item < B * heads; item += gridDim.x * POOL_WARPS) {
    const int b = item / heads, h = item % heads;
    __syncwarp();
    {  // q * head_dim^-0.5 (clip/model.py:78-79 -> F.multi_head_attention_forward scales the query)
      const __half2 qq = *reinterpret_cast<const __half2*>(q + static_cast<size_t>(b) * E + h * 64 + 2 * lane);
      const float2 f = __half22float2(qq);
      qs[2 * lane] = f.x * 0.125f;
      qs[2 * lane + 1] = f.y * 0.125f;
    }
    __syncwarp();
    const __half* kbase = kv + static_cast<size_t>(b) * L * 2 * E + h * 64;
    float mx = -INFINITY;
    for (int j = lane; j < Lp; j += 32) {
      float sc = -INFINITY;
      if (j < L) {
        const uint4* kr = reinterpret_cast<const uint4*>(kbase + static_cast<size_t>(j) * 2 * E);
        float acc = 0.0f;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const uint4 u = kr[v];
          const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h2[e]);
            acc = fmaf(f.x, qs[8 * v + 2 * e], fmaf(f.y, qs[8 * v + 2 * e + 1], acc));
          }
        }
        sc = acc;
      }
      ps[j] = sc;
      mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int j = lane; j < Lp; j += 32) {
      const float pj = (j < L) ? __expf(ps[j] - mx) : 0.0f;
      ps[j] = pj;
      sum += pj;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const __half* vbase = kbase + E;
    float o0 = 0.0f, o1 = 0.0f;
    for (int j0 = 0; j0 < L; j0 += 32) {  // 32 independent row loads in flight, then the FMAs
      __half2 vv[32];
#pragma unroll
      for (int u = 0; u < 32; ++u)
        vv[u] = *reinterpret_cast<const __half2*>(vbase + static_cast<size_t>(min(j0 + u, L - 1)) * 2 * E + 2 * lane);
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const float2 f = __half22float2(vv[u]);
        const float pj = (j0 + u < L) ? ps[j0 + u] : 0.0f;
        o0 = fmaf(pj, f.x, o0);
        o1 = fmaf(pj, f.y, o1);
      }
    }
    const float inv = 1.0f / sum;
    *reinterpret_cast<__half2*>(out + static_cast<size_t>(b) * E + h * 64 + 2 * lane) = __floats2half2_rn(o0 * inv, o1 * inv);
  }
}

int launch_attnpool_query0(const __half* q, const __half* kv, __half* out, int B, int L, int heads, cudaStream_t stream) {
  PC_REQUIRE(q && kv && out && B > 0 && L > 0 && heads > 0, PC_ERR_ARG, "attnpool_query0: bad arguments");
  const int Lp = (L + 31) & ~31;
  const size_t smem = static_cast<size_t>(POOL_WARPS) * (64 + Lp) * sizeof(float);
  PC_REQUIRE(smem <= 48 * 1024, PC_ERR_ARG, "attnpool_query0: %d tokens need %zu B of shared memory", L, smem);
  const int items = B * heads;
  const int want = (items + POOL_WARPS - 1) / POOL_WARPS, cap = device_sm_count() * 8;
  attnpool_query0_kernel<<<want < cap ? want : cap, POOL_WARPS * 32, smem, stream>>>(q, kv, out, B, L, heads);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_attnpool_tokens(const __half* x, const float* pos, __half* tok, int B, int HW, int C, cudaStream_t stream) {
  PC_REQUIRE(x && pos && tok && B > 0 && HW > 0 && C % 8 == 0, PC_ERR_ARG, "attnpool_tokens: B=%d HW=%d C=%d", B, HW, C);
  attnpool_tokens_kernel<<<grid_for(static_cast<size_t>(B) * (C / 8)), CT, 0, stream>>>(x, pos, tok, B, HW, C);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

}  // namespace pc
