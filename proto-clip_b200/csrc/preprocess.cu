// CLIP image preprocessing (reference clip/clip.py:77-84):
//   Resize(n_px, BICUBIC) -> CenterCrop(n_px) -> convert("RGB") -> ToTensor() -> Normalize(mean, std)
// for a batch of same-size RGB uint8 images [B, H, W, 3] already in device memory, byte-exact with what the reference runs on the host:
// Pillow's 8-bit antialiased resampler (src/libImaging/Resample.c: separable, fixed-point weights with 22 fractional
// bits, a uint8 intermediate image between the horizontal and the vertical pass) and torchvision's size / crop
// arithmetic. Only the n_px x n_px window that survives the centre crop is computed.
//
// Integer / byte work, HBM-bound and tiny: the weight tables are built on the host in double precision with the exact
// expression order of precompute_coeffs / normalize_coeffs_8bpc and travel to the device with one small H2D copy; two
// kernels (horizontal taps -> uint8 rows; vertical taps -> byte -> /255, -mean, /std with correctly rounded fp32 ops,
// tabulated per byte value) write the [3, n_px, n_px] tensor the encoders consume.
//
// The training-time augmentation of the support images (reference datasets/imagenet.py:8-23 `get_random_train_tfm`:
// RandomResizedCrop(224, scale (0.5, 1), BICUBIC) -> RandomHorizontalFlip -> ToTensor -> Normalize) is the same two
// passes over a different plan: the box drawn on the host is resampled as an image of its own (torchvision crops the
// PIL image first), both axes scaled independently, and the flip is the horizontal tables in reverse column order.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "kernels.cuh"

namespace pc {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;  // Resample.c

double bicubic_filter(double x) {  // Resample.c bicubic_filter, a = -0.5
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// precompute_coeffs + normalize_coeffs_8bpc for output indices [o0, o0 + n) of a resize in_size -> out_size.
// bounds[i] = (first source index, tap count), kk[i * ksize + t] = fixed-point weight. Returns ksize.
int coeffs(int in_size, int out_size, int o0, int n, std::vector<int>* bounds, std::vector<int>* kk) {
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
  bounds->assign(static_cast<size_t>(n) * 2, 0);
  kk->assign(static_cast<size_t>(n) * ksize, 0);
  if (in_size == out_size) {  // Pillow skips the pass: identity taps reproduce the bytes exactly
    for (int i = 0; i < n; ++i) {
      (*bounds)[2 * i] = o0 + i;
      (*bounds)[2 * i + 1] = 1;
      (*kk)[static_cast<size_t>(i) * ksize] = 1 << PRECISION_BITS;
    }
    return ksize;
  }
  const double ss = 1.0 / filterscale;
  std::vector<double> w(ksize);
  for (int i = 0; i < n; ++i) {
    const int xx = o0 + i;
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      (*kk)[static_cast<size_t>(i) * ksize + x] =
          v < 0 ? static_cast<int>(-0.5 + v * (1 << PRECISION_BITS)) : static_cast<int>(0.5 + v * (1 << PRECISION_BITS));
    }
    (*bounds)[2 * i] = xmin;
    (*bounds)[2 * i + 1] = xmax;
  }
  return ksize;
}

__device__ __forceinline__ int clip8(int acc) {  // Resample.c clip8
  const int v = acc >> PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Both passes are issue-bound byte work (ncu, round 2: 77 % issue slots at 0.14 - 0.16 of the HBM copy rate with one
// thread per element of a flat index space: two 64-bit div / mod pairs per element, a rolled tap loop and six IEEE
// divisions per output pixel). Here a block is a 64 x 4 patch of (column, row) so the index math is a handful of 32-bit
// operations, the tap loops are unrolled by four, the horizontal weights are stored tap-major (one coalesced load per tap
// for the 64 columns of a warp pair), and ToTensor + Normalize is a 3 x 256 entry table built on the host with the same
// correctly rounded fp32 operations (a byte has 256 values).
constexpr int PP_TX = 64, PP_TY = 4;

// horizontal pass: tmp[b][r][xx][c] for source rows y0 + r, r in [0, rows), and the n_px surviving output columns of
// every image b of the batch (images [B, H, W, 3], same size). The weights come in two layouts: kk_t tap-major
// (kk_t[x * n_px + xx], the byte loop) and kk4 in groups of four taps, zero-padded (kk4[g * n_px + xx] = taps 4 g .. 4 g + 3
// of column xx: one 16-byte load per group).
//
// A thread's taps are one contiguous run of 3 n source bytes. Byte loads at the 6.4-byte lane stride of a 2.1x downscale
// kept the L1 at 88 % of its sector rate (ncu); so the run is read as ALIGNED 32-bit words (a quarter of the requests),
// realigned with funnel shifts by the run's byte offset, and the twelve bytes of four taps are picked out of three
// registers. The words of the last tap group may reach up to 16 bytes past the run (zero weights there): they stay inside
// the buffer on every row but the last one of the last image (`last_row`), which takes the byte loop -- as does row 0 of
// image 0 when the buffer itself is not word-aligned (`head_unsafe`: the first word would start before it).
__device__ __forceinline__ int byte_at(uint32_t w, int i) { return static_cast<int>(__byte_perm(w, 0u, 0x4440u | i)); }  // one PRMT

// A block walks `rpt` steps of PP_TY adjacent rows of ONE image (blockIdx.z), so that the per-thread setup (bounds, table
// pointers, 64-bit address arithmetic: more instructions than the taps of one row) is paid once per rpt rows.
__global__ void __launch_bounds__(PP_TX * PP_TY)
resample_h_kernel(const uint8_t* __restrict__ src, int B, int H, int W, int y0, int rows, int n_px, int rpt,
                  const int2* __restrict__ bounds, const int* __restrict__ kk_t, const int4* __restrict__ kk4, int last_row,
                  int head_unsafe, size_t src_step, size_t out_step, uint8_t* __restrict__ tmp) {
  const int xx = blockIdx.x * PP_TX + threadIdx.x;
  if (xx >= n_px) return;
  const int2 bd = bounds[xx];  // (first source column, taps)
  const int* __restrict__ k = kk_t + xx;
  // taps in groups of four; a remainder of one or two taps takes a half group (two words, six bytes), of three a whole one
  const int groups = (bd.y + 1) >> 2;
  const bool half_group = static_cast<unsigned>((bd.y & 3) - 1) < 2u;
  const int4* __restrict__ k4_base = kk4 + xx;
  const int r_first = blockIdx.y * (PP_TY * rpt) + threadIdx.y;  // src_step / out_step: PP_TY rows of the source / of tmp
  for (int b = blockIdx.z; b < B; b += gridDim.z) {
    const uint8_t* __restrict__ p = src + ((static_cast<size_t>(b) * H + y0 + r_first) * W + bd.x) * 3;
    uint8_t* __restrict__ o = tmp + ((static_cast<size_t>(b) * rows + r_first) * n_px + xx) * 3;
    const int unsafe_from = b == B - 1 ? last_row - y0 : rows;     // rows r >= this take the byte loop
    const int unsafe_head = (head_unsafe && b == 0) ? -y0 : -1;    // ... and row r == this (source row 0)
    for (int i = 0, r = r_first; i < rpt && r < rows; ++i, r += PP_TY, p += src_step, o += out_step) {
      int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
      if (r >= unsafe_from || r == unsafe_head) {  // warp-uniform (a warp is one row)
#pragma unroll 4
        for (int x = 0; x < bd.y; ++x) {
          const int kv = k[x * n_px];
          a0 += p[3 * x] * kv;
          a1 += p[3 * x + 1] * kv;
          a2 += p[3 * x + 2] * kv;
        }
      } else {
        const int off = static_cast<int>(reinterpret_cast<uintptr_t>(p) & 3);
        const uint32_t* __restrict__ wp = reinterpret_cast<const uint32_t*>(p - off);  // the aligned word the run starts in
        const int4* __restrict__ k4 = k4_base;
        const int sh = off * 8;
        uint32_t w0 = __ldg(wp);
#pragma unroll 1
        for (int g = 0; g < groups; ++g) {
          const uint32_t w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
          const int4 kq = __ldg(k4);  // the weights of taps 4 g .. 4 g + 3 (zero past the run)
          // bytes 12 g .. 12 g + 11 of the run = taps 4 g .. 4 g + 3, three channels each
          const uint32_t r0 = __funnelshift_r(w0, w1, sh), r1 = __funnelshift_r(w1, w2, sh), r2 = __funnelshift_r(w2, w3, sh);
          a0 += byte_at(r0, 0) * kq.x + byte_at(r0, 3) * kq.y + byte_at(r1, 2) * kq.z + byte_at(r2, 1) * kq.w;
          a1 += byte_at(r0, 1) * kq.x + byte_at(r1, 0) * kq.y + byte_at(r1, 3) * kq.z + byte_at(r2, 2) * kq.w;
          a2 += byte_at(r0, 2) * kq.x + byte_at(r1, 1) * kq.y + byte_at(r2, 0) * kq.z + byte_at(r2, 3) * kq.w;
          w0 = w3;
          wp += 3;
          k4 += n_px;
        }
        if (half_group) {  // the last one or two taps: bytes 0 .. 5 of what is left of the run
          const uint32_t w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
          const int4 kq = __ldg(k4);  // .y is zero when a single tap is left
          const uint32_t r0 = __funnelshift_r(w0, w1, sh), r1 = __funnelshift_r(w1, w2, sh);
          a0 += byte_at(r0, 0) * kq.x + byte_at(r0, 3) * kq.y;
          a1 += byte_at(r0, 1) * kq.x + byte_at(r1, 0) * kq.y;
          a2 += byte_at(r0, 2) * kq.x + byte_at(r1, 1) * kq.y;
        }
      }
      o[0] = static_cast<uint8_t>(clip8(a0));
      o[1] = static_cast<uint8_t>(clip8(a1));
      o[2] = static_cast<uint8_t>(clip8(a2));
    }
  }
}

// vertical pass + ToTensor + Normalize: out[b][c][yy][xx] = lut[c][byte] with lut[c][v] = ((v / 255) - mean_c) / std_c,
// every operation rounded to fp32 (norm_lut() below)
template <typename OutT>
__global__ void __launch_bounds__(PP_TX * PP_TY)
resample_v_norm_kernel(const uint8_t* __restrict__ tmp, int B, int rows, int y0, int n_px, const int2* __restrict__ bounds,
                       const int* __restrict__ kk, int ksize, const float* __restrict__ lut, OutT* __restrict__ out) {
  const int xx = blockIdx.x * PP_TX + threadIdx.x;
  const int yy = blockIdx.y * PP_TY + threadIdx.y;
  if (xx >= n_px || yy >= n_px) return;
  const int2 bd = bounds[yy];  // (first source row, taps): the same for the whole warp
  const int* __restrict__ k = kk + yy * ksize;
  const int pitch = n_px * 3;
  const size_t plane = static_cast<size_t>(n_px) * n_px;
  for (int b = blockIdx.z; b < B; b += gridDim.z) {
    const uint8_t* __restrict__ p = tmp + ((static_cast<size_t>(b) * rows + (bd.x - y0)) * n_px + xx) * 3;
    int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
#pragma unroll 4
    for (int y = 0; y < bd.y; ++y) {
      const int kv = k[y];
      a0 += p[0] * kv;
      a1 += p[1] * kv;
      a2 += p[2] * kv;
      p += pitch;
    }
    OutT* o = out + static_cast<size_t>(b) * 3 * plane + static_cast<size_t>(yy) * n_px + xx;
    o[0] = static_cast<OutT>(__ldg(lut + clip8(a0)));
    o[plane] = static_cast<OutT>(__ldg(lut + 256 + clip8(a1)));
    o[2 * plane] = static_cast<OutT>(__ldg(lut + 512 + clip8(a2)));
  }
}

// The same for n_px % 4 == 0 (every CLIP resolution): a thread owns FOUR adjacent pixels = 12 consecutive bytes of the
// interleaved intermediate row = three aligned 32-bit words per tap instead of twelve byte loads (the byte version is
// bound by load-instruction issue), and writes one 16-byte (fp16: 8-byte) vector per channel plane.
template <typename OutT>
__global__ void __launch_bounds__(PP_TX * PP_TY)
resample_v_norm_wide_kernel(const uint8_t* __restrict__ tmp, int B, int rows, int y0, int n_px,
                            const int2* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                            const float* __restrict__ lut, OutT* __restrict__ out) {
  const int x4 = blockIdx.x * PP_TX + threadIdx.x;  // pixels 4 x4 .. 4 x4 + 3
  const int yy = blockIdx.y * PP_TY + threadIdx.y;
  if (4 * x4 >= n_px || yy >= n_px) return;
  const int2 bd = bounds[yy];
  const int* __restrict__ k = kk + yy * ksize;
  const int pitch_w = (n_px * 3) >> 2;  // words per intermediate row
  const size_t plane = static_cast<size_t>(n_px) * n_px;
  for (int b = blockIdx.z; b < B; b += gridDim.z) {
    const uint32_t* __restrict__ p =
        reinterpret_cast<const uint32_t*>(tmp + (static_cast<size_t>(b) * rows + (bd.x - y0)) * n_px * 3) + 3 * x4;
    int a[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = 1 << (PRECISION_BITS - 1);
#pragma unroll 2
    for (int y = 0; y < bd.y; ++y) {
      const int kv = k[y];
      const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] += byte_at(w0, i) * kv;
        a[4 + i] += byte_at(w1, i) * kv;
        a[8 + i] += byte_at(w2, i) * kv;
      }
      p += pitch_w;
    }
    OutT* o = out + static_cast<size_t>(b) * 3 * plane + static_cast<size_t>(yy) * n_px + 4 * x4;
#pragma unroll
    for (int c = 0; c < 3; ++c) {  // byte 3 px + c of the group: channel c of pixel px
      const float v0 = __ldg(lut + 256 * c + clip8(a[c])), v1 = __ldg(lut + 256 * c + clip8(a[3 + c]));
      const float v2 = __ldg(lut + 256 * c + clip8(a[6 + c])), v3 = __ldg(lut + 256 * c + clip8(a[9 + c]));
      if (sizeof(OutT) == 4) {
        *reinterpret_cast<float4*>(o + c * plane) = make_float4(v0, v1, v2, v3);
      } else {
        const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&h01);
        u.y = *reinterpret_cast<const uint32_t*>(&h23);
        *reinterpret_cast<uint2*>(o + c * plane) = u;
      }
    }
  }
}

// ToTensor (byte / 255) and Normalize ((x - mean) / std) of clip/clip.py:82-83 for every byte value: IEEE single-precision
// division and subtraction, round to nearest even -- what torch's CPU kernels and the device's __fdiv_rn / __fsub_rn compute
// (the host compiler keeps float expressions in float: SSE, no fast-math)
struct NormLut {
  float v[3 * 256];
  NormLut() {
    const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};  // clip/clip.py:83
    const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
    for (int c = 0; c < 3; ++c)
      for (int b = 0; b < 256; ++b) {
        volatile float x = static_cast<float>(b) / 255.0f;  // volatile: every intermediate is a rounded fp32 value
        volatile float d = x - mean[c];
        v[c * 256 + b] = d / stdv[c];
      }
  }
};
const float* norm_lut() {
  static const NormLut lut;  // thread-safe one-time construction
  return lut.v;
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct Plan {
  int new_h, new_w, top, left, y0, y1;
  int ksize_h, ksize_v;
  std::vector<int> bh, kh, bv, kv;
};

// Python's round(): half to even (torchvision F.center_crop: int(round((size - crop) / 2.0)))
int round_half_even(double v) {
  const double f = floor(v);
  const double d = v - f;
  if (d > 0.5) return static_cast<int>(f) + 1;
  if (d < 0.5) return static_cast<int>(f);
  return (static_cast<long long>(f) % 2 == 0) ? static_cast<int>(f) : static_cast<int>(f) + 1;
}

void make_plan(int H, int W, int n_px, Plan* p) {
  // torchvision _compute_resized_output_size: shorter side -> n_px, longer side int(n_px * long / short)
  const int shrt = W <= H ? W : H, lng = W <= H ? H : W;
  const int new_long = static_cast<int>(static_cast<double>(n_px) * lng / shrt);
  p->new_w = W <= H ? n_px : new_long;
  p->new_h = W <= H ? new_long : n_px;
  p->top = round_half_even((p->new_h - n_px) / 2.0);
  p->left = round_half_even((p->new_w - n_px) / 2.0);
  p->ksize_h = coeffs(W, p->new_w, p->left, n_px, &p->bh, &p->kh);
  p->ksize_v = coeffs(H, p->new_h, p->top, n_px, &p->bv, &p->kv);
  p->y0 = p->bv[0];
  p->y1 = p->bv[2 * (n_px - 1)] + p->bv[2 * (n_px - 1) + 1];
}

constexpr int LUT_WORDS = 3 * 256;
inline int pad4(int v) { return (v + 3) & ~3; }
size_t table_bytes(int n_px, int ksize_h, int ksize_v) {  // horizontal taps zero-padded to a multiple of four, two layouts
  return align256((static_cast<size_t>(n_px) * (4 + 2 * pad4(ksize_h) + ksize_v) + LUT_WORDS) * sizeof(int));
}

int taps(int in_size, int out_size) {  // ksize of precompute_coeffs
  const double scale = static_cast<double>(in_size) / out_size;
  return static_cast<int>(ceil(2.0 * (scale < 1.0 ? 1.0 : scale))) * 2 + 1;
}

// host tables -> device, then the two passes; `src` points at the first pixel of the resampled window (row pitch W,
// image pitch H * W), plan.y0 / y1 are rows of that window
// `last_row`: first row (relative to `src`, in the last image) whose tap words could leave the buffer -- the last row, or
// the last few when a row is shorter than the 16 bytes a run may be over-read by; `head_unsafe`: the buffer does not
// start on a word boundary
int run_plan(const uint8_t* src, int B, int H, int W, int n_px, const Plan& p, int last_row, int head_unsafe, void* out,
             int out_f16, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int rows = p.y1 - p.y0;
  const size_t tb = table_bytes(n_px, p.ksize_h, p.ksize_v);
  PC_REQUIRE(workspace_bytes >= tb + align256(static_cast<size_t>(B) * rows * n_px * 3), PC_ERR_WORKSPACE,
             "preprocess: workspace %zu < %zu", workspace_bytes, tb + align256(static_cast<size_t>(B) * rows * n_px * 3));
  // one H2D copy of all tables: [bh | bv | kh grouped by four taps | kh tap-major | kv | lut]
  std::vector<int> host;
  const int khp = pad4(p.ksize_h);
  host.reserve(static_cast<size_t>(n_px) * (4 + 2 * khp + p.ksize_v) + LUT_WORDS);
  host.insert(host.end(), p.bh.begin(), p.bh.end());
  host.insert(host.end(), p.bv.begin(), p.bv.end());
  const size_t k40 = host.size(), kt0 = k40 + static_cast<size_t>(n_px) * khp;
  host.resize(kt0 + static_cast<size_t>(n_px) * khp, 0);
  for (int xx = 0; xx < n_px; ++xx)
    for (int x = 0; x < p.ksize_h; ++x) {
      const int w = p.kh[static_cast<size_t>(xx) * p.ksize_h + x];
      host[k40 + (static_cast<size_t>(x >> 2) * n_px + xx) * 4 + (x & 3)] = w;
      host[kt0 + static_cast<size_t>(x) * n_px + xx] = w;
    }
  host.insert(host.end(), p.kv.begin(), p.kv.end());
  const size_t lut0 = host.size();
  host.resize(lut0 + LUT_WORDS);
  memcpy(host.data() + lut0, norm_lut(), LUT_WORDS * sizeof(float));
  int* d_bh = static_cast<int*>(workspace);
  int* d_bv = d_bh + 2 * n_px;
  int* d_k4 = d_bv + 2 * n_px;  // 16 n_px bytes into a 256-byte aligned buffer: int4-aligned
  int* d_kh = d_k4 + static_cast<size_t>(n_px) * khp;
  int* d_kv = d_kh + static_cast<size_t>(n_px) * khp;
  const float* d_lut = reinterpret_cast<const float*>(d_kv + static_cast<size_t>(n_px) * p.ksize_v);
  uint8_t* tmp = static_cast<uint8_t*>(workspace) + tb;
  PC_CHECK_CUDA(cudaMemcpyAsync(d_bh, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  const dim3 block(PP_TX, PP_TY);
  const unsigned gx = static_cast<unsigned>((n_px + PP_TX - 1) / PP_TX);
  // rows per thread of the horizontal pass: up to 8 once the grid still fills the device several times over
  const long long row_steps = (static_cast<long long>(rows) + PP_TY - 1) / PP_TY;
  long long rpt = row_steps * gx * B / (static_cast<long long>(device_sm_count()) * 32);
  rpt = rpt < 1 ? 1 : (rpt > 8 ? 8 : rpt);
  const dim3 g1(gx, static_cast<unsigned>((row_steps + rpt - 1) / rpt), static_cast<unsigned>(B < 65535 ? B : 65535));
  const dim3 g2(gx, static_cast<unsigned>((n_px + PP_TY - 1) / PP_TY), static_cast<unsigned>(B < 65535 ? B : 65535));
  resample_h_kernel<<<g1, block, 0, stream>>>(src, B, H, W, p.y0, rows, n_px, static_cast<int>(rpt),
                                              reinterpret_cast<const int2*>(d_bh), d_kh, reinterpret_cast<const int4*>(d_k4),
                                              last_row, head_unsafe, static_cast<size_t>(PP_TY) * W * 3,
                                              static_cast<size_t>(PP_TY) * n_px * 3, tmp);
  PC_CHECK_CUDA(cudaGetLastError());
  // four pixels per thread when the rows split into aligned words and the output into aligned vectors
  static const bool wide_ok = [] { const char* e = getenv("PC_PP_WIDE"); return !(e && e[0] == '0'); }();  // A/B switch
  const bool wide = wide_ok && (n_px & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (wide) {
    const dim3 g2w(static_cast<unsigned>((n_px / 4 + PP_TX - 1) / PP_TX), g2.y, g2.z);
    if (out_f16)
      resample_v_norm_wide_kernel<__half><<<g2w, block, 0, stream>>>(tmp, B, rows, p.y0, n_px, reinterpret_cast<const int2*>(d_bv),
                                                                    d_kv, p.ksize_v, d_lut, static_cast<__half*>(out));
    else
      resample_v_norm_wide_kernel<float><<<g2w, block, 0, stream>>>(tmp, B, rows, p.y0, n_px, reinterpret_cast<const int2*>(d_bv),
                                                                   d_kv, p.ksize_v, d_lut, static_cast<float*>(out));
  } else if (out_f16)
    resample_v_norm_kernel<__half><<<g2, block, 0, stream>>>(tmp, B, rows, p.y0, n_px, reinterpret_cast<const int2*>(d_bv),
                                                             d_kv, p.ksize_v, d_lut, static_cast<__half*>(out));
  else
    resample_v_norm_kernel<float><<<g2, block, 0, stream>>>(tmp, B, rows, p.y0, n_px, reinterpret_cast<const int2*>(d_bv),
                                                            d_kv, p.ksize_v, d_lut, static_cast<float*>(out));
  PC_CHECK_CUDA(cudaGetLastError());
  return PC_OK;
}

// RandomResizedCrop's resize (torchvision F.resized_crop on a PIL image): the ch x cw box is cropped FIRST, so Pillow
// resamples an image of that size -- taps stop at the box edge -- to n_px x n_px (both axes scaled independently).
// RandomHorizontalFlip after it = the horizontal tables in reverse column order.
void make_plan_window(int ch, int cw, int n_px, int flip, Plan* p) {
  p->new_h = p->new_w = n_px;
  p->top = p->left = 0;
  p->ksize_h = coeffs(cw, n_px, 0, n_px, &p->bh, &p->kh);
  p->ksize_v = coeffs(ch, n_px, 0, n_px, &p->bv, &p->kv);
  if (flip) {
    for (int a = 0, b = n_px - 1; a < b; ++a, --b) {
      std::swap(p->bh[2 * a], p->bh[2 * b]);
      std::swap(p->bh[2 * a + 1], p->bh[2 * b + 1]);
      std::swap_ranges(p->kh.begin() + static_cast<size_t>(a) * p->ksize_h,
                       p->kh.begin() + static_cast<size_t>(a + 1) * p->ksize_h,
                       p->kh.begin() + static_cast<size_t>(b) * p->ksize_h);
    }
  }
  p->y0 = p->bv[0];
  p->y1 = p->bv[2 * (n_px - 1)] + p->bv[2 * (n_px - 1) + 1];
}

}  // namespace

size_t preprocess_workspace_bytes(int B, int H, int W, int n_px) {
  if (B <= 0 || H <= 0 || W <= 0 || n_px <= 0) return 0;
  const int shrt = W <= H ? W : H, lng = W <= H ? H : W;
  const int new_long = static_cast<int>(static_cast<double>(n_px) * lng / shrt);
  // the intermediate image holds at most H source rows of the n_px surviving columns
  return table_bytes(n_px, taps(W, W <= H ? n_px : new_long), taps(H, W <= H ? new_long : n_px)) +
         align256(static_cast<size_t>(B) * H * n_px * 3);
}

size_t preprocess_train_workspace_bytes(int ch, int cw, int n_px) {
  if (ch <= 0 || cw <= 0 || n_px <= 0) return 0;
  return table_bytes(n_px, taps(cw, n_px), taps(ch, n_px)) + align256(static_cast<size_t>(ch) * n_px * 3);
}

int launch_preprocess(const uint8_t* rgb, int B, int H, int W, int n_px, void* out, int out_f16, void* workspace,
                      size_t workspace_bytes, cudaStream_t stream) {
  PC_REQUIRE(rgb && out && workspace, PC_ERR_ARG, "preprocess: null buffer");
  PC_REQUIRE(B > 0 && H > 0 && W > 0 && n_px > 0 && n_px <= 4096 && H <= 32768 && W <= 32768, PC_ERR_ARG,
             "preprocess: %d images %dx%d -> %d px", B, H, W, n_px);
  // the filter tables depend on (H, W, n_px) only: the last plan is kept (a loader's images mostly share one size)
  static thread_local Plan p;
  static thread_local int pH = 0, pW = 0, pN = 0;
  if (pH != H || pW != W || pN != n_px) {
    make_plan(H, W, n_px, &p);
    pH = H; pW = W; pN = n_px;
  }
  PC_REQUIRE(p.new_h >= n_px && p.new_w >= n_px, PC_ERR_ARG, "preprocess: resized image %dx%d smaller than the crop %d",
             p.new_h, p.new_w, n_px);
  return run_plan(rgb, B, H, W, n_px, p, H - 1 - 15 / (3 * W), (reinterpret_cast<uintptr_t>(rgb) & 3) != 0, out, out_f16,
                  workspace, workspace_bytes, stream);
}

int launch_preprocess_train(const uint8_t* rgb, int H, int W, int top, int left, int ch, int cw, int flip, int n_px,
                            void* out, int out_f16, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PC_REQUIRE(rgb && out && workspace, PC_ERR_ARG, "preprocess_train: null buffer");
  PC_REQUIRE(H > 0 && W > 0 && n_px > 0 && n_px <= 4096 && H <= 32768 && W <= 32768, PC_ERR_ARG,
             "preprocess_train: image %dx%d -> %d px", H, W, n_px);
  PC_REQUIRE(top >= 0 && left >= 0 && ch > 0 && cw > 0 && top <= H - ch && left <= W - cw, PC_ERR_ARG,
             "preprocess_train: box (top %d, left %d, %d x %d) outside the %d x %d image", top, left, ch, cw, H, W);
  // every call draws a new box: no plan cache (the tables are n_px * (ksize_h + ksize_v) doubles of host work)
  Plan p;
  make_plan_window(ch, cw, n_px, flip != 0, &p);
  const uint8_t* src = rgb + (static_cast<size_t>(top) * W + left) * 3;
  return run_plan(src, 1, H, W, n_px, p, H - 1 - top - 15 / (3 * W), (reinterpret_cast<uintptr_t>(rgb) & 3) != 0 && top == 0,
                  out, out_f16, workspace, workspace_bytes, stream);
}

}  // namespace pc
