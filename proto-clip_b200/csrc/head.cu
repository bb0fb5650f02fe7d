// Proto-CLIP head: prototype construction, the learnable query adapters, and the tail of P().
//   prototypes       main.py:399-405 (per-shot L2 norm -> mean over K -> L2 norm; text: L2 norm)
//   Adapter_FC       model.py:81-95  (the two bias-free Linears run on the tcgen05 GEMM; LN + blend here)
//   Adapter conv     model.py:12-78  (one CTA per query, all 16xSxS activations resident in smem)
//   P()              utils.py:225-244 (dots come from the tcgen05 GEMM; distance, dual softmax, blend, argmax here)
// All of it is HBM/latency-bound elementwise + reduction work: warp-shuffle reductions, fp32 statistics,
// fp16 storage with the reference's rounding points.
#include "kernels.cuh"
#include "ptx.cuh"

namespace pc {

namespace {

__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.0f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : -INFINITY;
  t = warp_max(t);
  return t;
}

// One CTA per class: z[n] = norm(mean_k(norm(V[n,k]))), every stage rounded to fp16 like the eager fp16
// tensor ops of the reference; zn2[n] = sum(float(z)^2) feeds the distance in P().
__global__ void __launch_bounds__(128)
prototypes_kernel(const __half* __restrict__ V, int K, int D, int per_shot_norm, __half* __restrict__ z,
                  float* __restrict__ zn2) {
  extern __shared__ float sm[];  // [D] running sum over shots
  __shared__ float red[32];
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < D; c += blockDim.x) sm[c] = 0.0f;
  for (int k = 0; k < K; ++k) {
    const __half* row = V + (static_cast<size_t>(n) * K + k) * D;
    float nrm = 1.0f;
    if (per_shot_norm) {
      float s = 0.0f;
      for (int c = threadIdx.x; c < D; c += blockDim.x) {
        const float x = __half2float(row[c]);
        s += x * x;
      }
      nrm = r16(sqrtf(block_sum(s, red)));
    }
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      const float x = __half2float(row[c]);
      sm[c] += per_shot_norm ? r16(x / nrm) : x;
    }
  }
  __syncthreads();
  float s = 0.0f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float m = r16(sm[c] / static_cast<float>(K));  // fp16 mean (fp32 accumulate)
    sm[c] = m;
    s += m * m;
  }
  const float nrm = r16(sqrtf(block_sum(s, red)));
  float s2 = 0.0f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = r16(sm[c] / nrm);
    z[static_cast<size_t>(n) * D + c] = __float2half_rn(v);
    s2 += v * v;
  }
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0 && zn2) zn2[n] = s2;
}

// LayerNorm with fp16 affine parameters (nn.LayerNorm(dtype=half): fp32 statistics, fp16 result),
// optionally followed by Adapter_FC's blend  out = ratio * LN(h) + (1 - ratio) * x_in  in fp16 ops.
__global__ void __launch_bounds__(256)
ln_f16_kernel(const __half* __restrict__ h, const __half* __restrict__ x_in, __half* __restrict__ y,
              const __half* __restrict__ gamma, const __half* __restrict__ beta, float ratio, int rows, int d) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const __half* src = h + static_cast<size_t>(row) * d;
  float s = 0.0f;
  for (int c = lane; c < d; c += 32) s += __half2float(src[c]);
  const float mean = warp_sum(s) / static_cast<float>(d);
  float q = 0.0f;
  for (int c = lane; c < d; c += 32) {
    const float t = __half2float(src[c]) - mean;
    q += t * t;
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(d) + 1e-5f);
  for (int c = lane; c < d; c += 32) {
    float v = r16((__half2float(src[c]) - mean) * rstd * __half2float(gamma[c]) + __half2float(beta[c]));
    if (x_in) {
      const float xi = __half2float(x_in[static_cast<size_t>(row) * d + c]);
      v = r16(r16(ratio * v) + r16((1.0f - ratio) * xi));
    }
    y[static_cast<size_t>(row) * d + c] = __float2half_rn(v);
  }
}

// Adapter (conv-2x / conv-3x), one CTA per query. a/b are ping-pong [16][S][S] fp16 planes in smem.
__device__ void plane_layernorm(__half* buf, int n, const __half* __restrict__ w, const __half* __restrict__ b,
                                float* red) {
  float s = 0.0f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += __half2float(buf[i]);
  const float mean = block_sum(s, red) / static_cast<float>(n);
  float q = 0.0f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float t = __half2float(buf[i]) - mean;
    q += t * t;
  }
  const float rstd = rsqrtf(block_sum(q, red) / static_cast<float>(n) + 1e-5f);
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    buf[i] = __float2half_rn((__half2float(buf[i]) - mean) * rstd * __half2float(w[i]) + __half2float(b[i]));
  __syncthreads();
}

__global__ void __launch_bounds__(256)
adapter_conv_kernel(const AdapterConvW w, int three_x, const __half* __restrict__ q, __half* __restrict__ out,
                    int D, int S) {
  extern __shared__ __half sh[];
  __shared__ float red[32];
  __shared__ float w1[16], w3[16];
  __shared__ float w2[16 * 16 * 9];
  const int SS = S * S;
  __half* x = sh;            // [SS]   zero-padded input plane (identity branch)
  __half* a = sh + SS;       // [16][SS]
  __half* b = a + 16 * SS;   // [16][SS]
  const size_t qi = blockIdx.x;
  if (threadIdx.x < 16) {
    w1[threadIdx.x] = __half2float(w.conv1[threadIdx.x]);
    w3[threadIdx.x] = __half2float(w.conv3[threadIdx.x]);
  }
  if (three_x)
    for (int i = threadIdx.x; i < 16 * 16 * 9; i += blockDim.x) w2[i] = __half2float(w.conv2[i]);
  for (int i = threadIdx.x; i < SS; i += blockDim.x) x[i] = (i < D) ? q[qi * D + i] : __float2half_rn(0.0f);
  __syncthreads();
  // conv1: 1x1, 1 -> 16
  for (int i = threadIdx.x; i < 16 * SS; i += blockDim.x)
    a[i] = __float2half_rn(w1[i / SS] * __half2float(x[i % SS]));
  __syncthreads();
  plane_layernorm(a, 16 * SS, w.bn1_w, w.bn1_b, red);
  __half* cur = a;
  if (three_x) {
    // conv2: 3x3, 16 -> 16, zero padding 1, fp32 accumulation, fp16 result
    for (int i = threadIdx.x; i < 16 * SS; i += blockDim.x) {
      const int co = i / SS, yx = i % SS, yy = yx / S, xx = yx % S;
      float acc = 0.0f;
      for (int ci = 0; ci < 16; ++ci) {
        const __half* plane = a + ci * SS;
        const float* wk = w2 + (co * 16 + ci) * 9;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int y2 = yy + dy;
          if (y2 < 0 || y2 >= S) continue;
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int x2 = xx + dx;
            if (x2 < 0 || x2 >= S) continue;
            acc = fmaf(wk[(dy + 1) * 3 + (dx + 1)], __half2float(plane[y2 * S + x2]), acc);
          }
        }
      }
      b[i] = __float2half_rn(acc);
    }
    __syncthreads();
    plane_layernorm(b, 16 * SS, w.bn2_w, w.bn2_b, red);
    cur = b;
  }
  // conv3: 1x1, 16 -> 1 (result into the other buffer's first plane)
  __half* o = (cur == a) ? b : a;
  for (int i = threadIdx.x; i < SS; i += blockDim.x) {
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < 16; ++c) acc = fmaf(w3[c], __half2float(cur[c * SS + i]), acc);
    o[i] = __float2half_rn(acc);
  }
  __syncthreads();
  plane_layernorm(o, SS, w.bn3_w, w.bn3_b, red);
  for (int i = threadIdx.x; i < D; i += blockDim.x) out[qi * D + i] = __hadd(o[i], x[i]);
}

// Tail of P(): one CTA per query row. dots[row, 0:N] = q.z_img, dots[row, N:2N] = q.z_txt (fp32).
// dist^2 = max(|q|^2 + |z|^2 - 2 q.z, 0) (torch.cdist's matmul form); p = a*softmax(-b*d_i) + (1-a)*softmax(-b*d_t).
__global__ void __launch_bounds__(256)
proto_softmax_kernel(const float* __restrict__ dots, int ld, const __half* __restrict__ q, int D,
                     const float* __restrict__ zi_n2, const float* __restrict__ zt_n2, int N, float alpha,
                     float beta, float* __restrict__ p_out, int64_t* __restrict__ argmax,
                     float* __restrict__ pmax) {
  __shared__ float red[32];
  __shared__ int red_i[32];
  const size_t row = blockIdx.x;
  const float* di = dots + row * ld;
  const float* dt = di + (ld >> 1);  // second bank starts at the padded half
  float s = 0.0f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float x = __half2float(q[row * D + c]);
    s += x * x;
  }
  const float qn2 = block_sum(s, red);
  float mi = -INFINITY, mt = -INFINITY;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    mi = fmaxf(mi, -beta * fmaxf(qn2 + zi_n2[n] - 2.0f * di[n], 0.0f));
    mt = fmaxf(mt, -beta * fmaxf(qn2 + zt_n2[n] - 2.0f * dt[n], 0.0f));
  }
  mi = block_max(mi, red);
  mt = block_max(mt, red);
  float si = 0.0f, st = 0.0f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    si += __expf(-beta * fmaxf(qn2 + zi_n2[n] - 2.0f * di[n], 0.0f) - mi);
    st += __expf(-beta * fmaxf(qn2 + zt_n2[n] - 2.0f * dt[n], 0.0f) - mt);
  }
  si = block_sum(si, red);
  st = block_sum(st, red);
  const float ci = alpha / si, ct = (1.0f - alpha) / st;
  float best = -1.0f;
  int arg = 0x7fffffff;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float p = ci * __expf(-beta * fmaxf(qn2 + zi_n2[n] - 2.0f * di[n], 0.0f) - mi) +
                    ct * __expf(-beta * fmaxf(qn2 + zt_n2[n] - 2.0f * dt[n], 0.0f) - mt);
    if (p_out) p_out[row * N + n] = p;
    if (p > best) {  // strided ascending n: keeps the first maximum per thread
      best = p;
      arg = n;
    }
  }
  // arg-max with lowest-index tie-break
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) {
      best = ob;
      arg = oa;
    }
  }
  const int wi = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) {
    red[wi] = best;
    red_i[wi] = arg;
  }
  __syncthreads();
  if (wi == 0) {
    best = (l < nw) ? red[l] : -1.0f;
    arg = (l < nw) ? red_i[l] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) {
        best = ob;
        arg = oa;
      }
    }
    if (l == 0) {
      if (argmax) argmax[row] = arg;
      if (pmax) pmax[row] = best;
    }
  }
}

// (alpha, beta) grid of P() + argmax + accuracy in one pass over the similarities (main.py:187-199, 419-430: the
// reference recomputes the two cdist matmuls and both softmaxes 319 x 3 times). One CTA per query row: the squared
// distances to both prototype banks stay in shared memory; per beta the two softmaxes are evaluated once, per alpha
// only the blend + argmax. counts[a * n_beta + b] += (argmax == label). Same arithmetic as proto_softmax_kernel.
__global__ void __launch_bounds__(256)
proto_grid_kernel(const float* __restrict__ dots, int ld, const __half* __restrict__ q, int D,
                  const float* __restrict__ zi_n2, const float* __restrict__ zt_n2, int N,
                  const int64_t* __restrict__ labels, const float* __restrict__ alphas, int n_alpha,
                  const float* __restrict__ betas, int n_beta, int* __restrict__ counts) {
  extern __shared__ float gsm[];  // d_i [N] | d_t [N] | e_i [N] | e_t [N]
  __shared__ float red[32];
  __shared__ int red_i[32];
  float* d_i = gsm;
  float* d_t = gsm + N;
  float* e_i = gsm + 2 * N;
  float* e_t = gsm + 3 * N;
  const size_t row = blockIdx.x;
  const float* di = dots + row * ld;
  const float* dt = di + (ld >> 1);
  float s = 0.0f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float x = __half2float(q[row * D + c]);
    s += x * x;
  }
  const float qn2 = block_sum(s, red);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    d_i[n] = fmaxf(qn2 + zi_n2[n] - 2.0f * di[n], 0.0f);
    d_t[n] = fmaxf(qn2 + zt_n2[n] - 2.0f * dt[n], 0.0f);
  }
  __syncthreads();
  const int label = static_cast<int>(labels[row]);
  const int wi = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int b = 0; b < n_beta; ++b) {
    const float beta = betas[b];
    float mi = -INFINITY, mt = -INFINITY;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      mi = fmaxf(mi, -beta * d_i[n]);
      mt = fmaxf(mt, -beta * d_t[n]);
    }
    mi = block_max(mi, red);
    mt = block_max(mt, red);
    float si = 0.0f, st = 0.0f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const float a = __expf(-beta * d_i[n] - mi), c = __expf(-beta * d_t[n] - mt);
      e_i[n] = a;
      e_t[n] = c;
      si += a;
      st += c;
    }
    si = block_sum(si, red);
    st = block_sum(st, red);
    for (int a = 0; a < n_alpha; ++a) {
      const float alpha = alphas[a];
      const float ci = alpha / si, ct = (1.0f - alpha) / st;
      float best = -1.0f;
      int arg = 0x7fffffff;
      for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float pv = ci * e_i[n] + ct * e_t[n];
        if (pv > best) {
          best = pv;
          arg = n;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) {
          best = ob;
          arg = oa;
        }
      }
      __syncthreads();
      if (l == 0) {
        red[wi] = best;
        red_i[wi] = arg;
      }
      __syncthreads();
      if (wi == 0) {
        best = (l < nw) ? red[l] : -1.0f;
        arg = (l < nw) ? red_i[l] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
          if (ob > best || (ob == best && oa < arg)) {
            best = ob;
            arg = oa;
          }
        }
        if (l == 0 && arg == label) atomicAdd(&counts[a * n_beta + b], 1);
      }
    }
    __syncthreads();  // e_i / e_t are rewritten by the next beta
  }
}

}  // namespace

int launch_proto_grid(const float* dots, int ld, const __half* q, int D, const float* zi_n2, const float* zt_n2, int Q,
                      int N, const int64_t* labels, const float* alphas, int n_alpha, const float* betas, int n_beta,
                      int* counts, cudaStream_t stream) {
  PC_REQUIRE(dots && q && zi_n2 && zt_n2 && labels && alphas && betas && counts && Q > 0 && N > 0 && n_alpha > 0 &&
                 n_beta > 0,
             PC_ERR_ARG, "proto_grid: bad args");
  const int smem = 4 * N * static_cast<int>(sizeof(float));
  PC_REQUIRE(smem <= 200 * 1024, PC_ERR_ARG, "proto_grid: N=%d needs %d B smem", N, smem);
  static int configured[kMaxDevices];
  if (smem > 48 * 1024) PC_CHECK_CUDA(ensure_dynamic_smem(proto_grid_kernel, smem, configured));
  proto_grid_kernel<<<Q, 256, smem, stream>>>(dots, ld, q, D, zi_n2, zt_n2, N, labels, alphas, n_alpha, betas, n_beta,
                                              counts);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_build_prototypes(const __half* V, int N, int K, int D, int per_shot_norm, __half* z, float* zn2,
                            cudaStream_t stream) {
  PC_REQUIRE(V && z && N > 0 && K > 0 && D > 0, PC_ERR_ARG, "build_prototypes: bad args N=%d K=%d D=%d", N, K, D);
  PC_REQUIRE(D * 4 <= 48 * 1024, PC_ERR_ARG, "build_prototypes: D=%d too large", D);
  prototypes_kernel<<<N, 128, D * sizeof(float), stream>>>(V, K, D, per_shot_norm, z, zn2);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_text_prototypes(const __half* T, int N, int D, __half* z, float* zn2, cudaStream_t stream) {
  // z_txt = T / |T| is the K = 1 case without the (idempotent up to rounding) per-shot stage:
  // mean over one shot is the identity, then one L2 normalisation (main.py:404-405).
  return launch_build_prototypes(T, N, 1, D, 0, z, zn2, stream);
}

int launch_ln_f16(const __half* x, __half* y, const __half* gamma, const __half* beta, int rows, int d,
                  cudaStream_t stream) {
  PC_REQUIRE(x && y && gamma && beta && rows > 0 && d > 0, PC_ERR_ARG, "ln_f16: bad args");
  ln_f16_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(x, nullptr, y, gamma, beta, 0.0f, rows, d);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_ln_blend_f16(const __half* h, const __half* x_in, __half* y, const __half* gamma,
                        const __half* beta, float ratio, int rows, int d, cudaStream_t stream) {
  PC_REQUIRE(h && x_in && y && gamma && beta && rows > 0 && d > 0, PC_ERR_ARG, "ln_blend_f16: bad args");
  ln_f16_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(h, x_in, y, gamma, beta, ratio, rows, d);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_adapter_conv(const AdapterConvW& w, int three_x, const __half* q, __half* out, int Q, int D,
                        cudaStream_t stream) {
  PC_REQUIRE(q && out && Q > 0 && D > 0, PC_ERR_ARG, "adapter_conv: bad args");
  int S = 1;
  while (S * S < D) ++S;  // ceil(sqrt(D)) (model.py:27)
  const int smem = (1 + 32) * S * S * static_cast<int>(sizeof(__half));
  PC_REQUIRE(smem <= 200 * 1024, PC_ERR_ARG, "adapter_conv: D=%d needs %d B smem", D, smem);
  static int configured[kMaxDevices];
  PC_CHECK_CUDA(ensure_dynamic_smem(adapter_conv_kernel, smem, configured));
  adapter_conv_kernel<<<Q, 256, smem, stream>>>(w, three_x, q, out, D, S);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

int launch_proto_softmax(const float* dots, int ld, const __half* q, int D, const float* zi_n2,
                         const float* zt_n2, int Q, int N, float alpha, float beta, float* p_out,
                         int64_t* argmax, float* pmax, cudaStream_t stream) {
  PC_REQUIRE(dots && q && zi_n2 && zt_n2 && Q > 0 && N > 0, PC_ERR_ARG, "proto_softmax: bad args");
  proto_softmax_kernel<<<Q, 256, 0, stream>>>(dots, ld, q, D, zi_n2, zt_n2, N, alpha, beta, p_out, argmax, pmax);
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}
}  // namespace pc
