// C ABI of libprotoclip_b200 (include/protoclip_b200.h): context, weight binding and the orchestration of
// the kernels in gemm.cu / attention.cu / rowops.cu / head.cu into the reference's hot-path functions
// (encode_image, encode_text, ResidualAttentionBlock, Adapter*, prototypes, P).
#include <stdlib.h>

#include <new>
#include <vector>

#include "kernels.cuh"

using namespace pc;

static_assert(int(PC_EPI_BIAS) == int(EPI_BIAS) && int(PC_EPI_BIAS_QUICKGELU) == int(EPI_BIAS_QGELU) &&
                  int(PC_EPI_BIAS_RESIDUAL) == int(EPI_BIAS_RES) && int(PC_EPI_F32) == int(EPI_F32),
              "public epilogue enum must match the kernel enum");

namespace {

// LayerNorm folded into the Linear that consumes it (kernels.cuh: EPI_LN_*): per layer, for in_proj (ln_1) and
// c_fc (ln_2), the gamma-scaled weight and the two per-column vectors. Owned by the tower, built at bind time.
struct FoldedLn {
  __half* w = nullptr;  // [N, K] f16(W * gamma)
  float* s = nullptr;   // [N]
  float* c = nullptr;   // [N]
};

struct Tower {
  bool bound = false;
  int width = 0, layers = 0, heads = 0, L = 0, embed = 0;
  std::vector<pc_resblock_weights> blocks;
  std::vector<FoldedLn> qkv_ln, fc_ln;  // per layer
  void* fold_pool = nullptr;            // owned: one allocation behind qkv_ln / fc_ln
  __half* proj_t = nullptr;             // owned: [embed, width] (transposed projection -> TN GEMM)
};

// ModifiedResNet (clip/model.py:95-152): every conv + eval BatchNorm pair folded at bind time into a TN GEMM operand.
struct ConvBn {
  __half* w = nullptr;     // owned: [cout, Kp], column = tap * cin + ci, BatchNorm scale folded in
  float* shift = nullptr;  // owned: [cout] fp32
  int cout = 0, cin = 0, k = 0, Kp = 0;
};
struct RnBlock {
  ConvBn c1, c2, c3, down;
  bool has_down = false;
  int inplanes = 0, planes = 0, stride = 1;
};
struct RnTower {
  bool bound = false;
  int res = 0, width = 0, heads = 0, embed = 0, out_dim = 0, tokens = 0;
  ConvBn stem[3];
  std::vector<RnBlock> blocks;
  __half* qkv_w = nullptr;  // owned: [3E, E] = [q_proj; k_proj; v_proj] (packed in-proj order of the attention kernel)
  __half* qkv_b = nullptr;  // owned: [3E]
  const float* pos = nullptr;
  const __half *c_w = nullptr, *c_b = nullptr;
  std::vector<void*> owned;
  size_t max_act = 0, max_col = 0;  // per image, in fp16 elements: largest activation / im2col operand
};
// ModifiedResNet towers: at most kRnMicroBatch images per pass, the batch cut into equal passes. RN50x16 @384 px on one
// box (tools/rn_pass_sweep.py, 512 images): 4 496 - 4 510 img/s at 64 per pass, 4 538 at 128, 4 583 - 4 604 at 192,
// 4 635 - 4 679 at 256, 4 531 - 4 550 at 512 (bit-identical features at every pass size).
constexpr int kRnMicroBatch = 256;
inline int pick_rn_micro_batch(int B, int micro_batch) {
  if (micro_batch > 0) return micro_batch;
  const int passes = (B + kRnMicroBatch - 1) / kRnMicroBatch;
  return passes > 0 ? (B + passes - 1) / passes : kRnMicroBatch;
}

// Micro-batching. One row-block WAVE of the CTA-pair GEMM is sm_count / 2 pairs x 256 rows: ViT-B/16 (L = 197) -> 96
// images = 18 912 rows = 73.9 of 74 row blocks; ViT-L/14 -> 73; @336px -> 32. Round 1 ran one wave per pass (activations
// L2-resident); measured on B200, longer passes win: every launch pays ~2 us of launch gap plus the fill of its first
// tile and the un-overlapped epilogue of its last one, and the Linears stay compute-bound from HBM (ViT-B/16, batch
// 1024: 26.1 k img/s at 96 per pass, 26.7 k at 288, 27.3 k at 512; later in the round 29.4 k at 512 against 29.9 k for
// the whole 1024 in one pass). Default: at most kMaxWaves waves per pass (1 152 ViT-B/16 images: ~3 GB of workspace), and
// the batch cut into EQUAL passes (2048 -> 2 x 1024, not 1152 + 896): tiles, not row blocks, are what the workers balance.
constexpr int kMaxWaves = 12;
inline int wave_images(int L) {
  const int rows = (device_sm_count() / 2) * 256;
  const int mb = rows / (L > 0 ? L : 1);
  return mb > 0 ? mb : 1;
}
inline int default_micro_batch(int L) { return kMaxWaves * wave_images(L); }  // workspace sizing: the largest pass
// images per pass for a batch of B (micro_batch > 0: the caller's choice)
inline int pick_micro_batch(int L, int B, int micro_batch) {
  if (micro_batch > 0) return micro_batch;
  const int cap = default_micro_batch(L);
  const int passes = (B + cap - 1) / cap;
  return passes > 0 ? (B + passes - 1) / passes : cap;
}
constexpr int kClassifyChunk = 8192;    // queries per P() pass: [8192, 2N] fp32 dots stay L2-resident

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct pc_ctx {
  int device = 0;
  int full_last_block = -1;  // -1: follow PC_FULL_LAST_BLOCK; 0 / 1: pc_ctx_set_full_last_block
  Tower vis, txt;
  RnTower rn;  // bound instead of `vis` for ModifiedResNet checkpoints
  // visual stem
  int res = 0, patch = 0, grid = 0, Kpatch = 0, Kp = 0;
  __half* conv1_p = nullptr;  // owned: [width, Kp] zero-padded flattening of conv1.weight
  const float *cls = nullptr, *vpos = nullptr, *ln_pre_w = nullptr, *ln_pre_b = nullptr, *ln_post_w = nullptr,
              *ln_post_b = nullptr;
  // text stem
  int vocab = 0;
  const float *tok_emb = nullptr, *tpos = nullptr, *ln_final_w = nullptr, *ln_final_b = nullptr;
};

namespace {

__global__ void transpose_f16_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int rows, int cols) {
  __shared__ __half tile[32][33];
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    if (x < cols && y0 + j < rows) tile[j][threadIdx.x] = src[static_cast<size_t>(y0 + j) * cols + x];
  __syncthreads();
  const int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    if (xo < rows && yo0 + j < cols) dst[static_cast<size_t>(yo0 + j) * rows + xo] = tile[threadIdx.x][j];
}

int make_transposed(const void* src, int rows, int cols, __half** out) {
  if (*out) cudaFree(*out);
  PC_CHECK_CUDA(cudaMalloc(out, static_cast<size_t>(rows) * cols * sizeof(__half)));
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_f16_kernel<<<grid, block>>>(static_cast<const __half*>(src), *out, rows, cols);
  PC_CHECK_CUDA(cudaGetLastError());
  PC_CHECK_CUDA(cudaDeviceSynchronize());
  return PC_OK;
}

int check_blocks(const pc_resblock_weights* b, int layers) {
  PC_REQUIRE(b != nullptr, PC_ERR_ARG, "bind: blocks array is null");
  for (int i = 0; i < layers; ++i) {
    const void* const* p = reinterpret_cast<const void* const*>(&b[i]);
    for (size_t k = 0; k < sizeof(pc_resblock_weights) / sizeof(void*); ++k)
      PC_REQUIRE(p[k] != nullptr, PC_ERR_ARG, "bind: resblock %d tensor %zu is null", i, k);
  }
  return PC_OK;
}

bool planar_qkv_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PC_NO_PLANAR_QKV");  // A/B switch: 1 keeps the packed [rows, 3d] qkv layout
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

int use_device(const pc_ctx* ctx) {
  PC_REQUIRE(ctx != nullptr, PC_ERR_ARG, "null context");
  PC_CHECK_CUDA(cudaSetDevice(ctx->device));
  return PC_OK;
}

// x += MHA(LN1(x)); x += MLP(LN2(x))  on token-major x [rows = B*L, d]. h: [rows, d], big: [rows, 4d].
int resblock(const Tower& t, int layer, __half* x, __half* h, __half* big, int B, int L, int causal,
             cudaStream_t s) {
  const pc_resblock_weights& w = t.blocks[layer];
  const int d = t.width, rows = B * L;
  PC_TRY(launch_layernorm(x, h, static_cast<const float*>(w.ln_1_weight), static_cast<const float*>(w.ln_1_bias),
                          rows, d, 1, s));
  GemmArgs g{};
  g.M = rows; g.N = 3 * d; g.K = d;
  g.A = h; g.lda = d;
  g.W = static_cast<const __half*>(w.in_proj_weight); g.ldw = d;
  g.C = big; g.ldc = 3 * d;
  g.bias = static_cast<const __half*>(w.in_proj_bias);
  PC_TRY(launch_gemm(g, EPI_BIAS, s));
  PC_TRY(launch_attention(big, h, B, L, t.heads, causal, s));
  g = GemmArgs{};
  g.M = rows; g.N = d; g.K = d;
  g.A = h; g.lda = d;
  g.W = static_cast<const __half*>(w.out_proj_weight); g.ldw = d;
  g.C = x; g.ldc = d;
  g.bias = static_cast<const __half*>(w.out_proj_bias);
  g.residual = x; g.ldr = d;
  PC_TRY(launch_gemm(g, EPI_BIAS_RES, s));
  PC_TRY(launch_layernorm(x, h, static_cast<const float*>(w.ln_2_weight), static_cast<const float*>(w.ln_2_bias),
                          rows, d, 1, s));
  g = GemmArgs{};
  g.M = rows; g.N = 4 * d; g.K = d;
  g.A = h; g.lda = d;
  g.W = static_cast<const __half*>(w.c_fc_weight); g.ldw = d;
  g.C = big; g.ldc = 4 * d;
  g.bias = static_cast<const __half*>(w.c_fc_bias);
  PC_TRY(launch_gemm(g, EPI_BIAS_QGELU, s));
  g = GemmArgs{};
  g.M = rows; g.N = d; g.K = 4 * d;
  g.A = big; g.lda = 4 * d;
  g.W = static_cast<const __half*>(w.c_proj_weight); g.ldw = 4 * d;
  g.C = x; g.ldc = d;
  g.bias = static_cast<const __half*>(w.c_proj_bias);
  g.residual = x; g.ldr = d;
  PC_TRY(launch_gemm(g, EPI_BIAS_RES, s));
  return PC_OK;
}

// PC_NO_FUSED_LN=1 keeps LayerNorm as its own kernel (A/B timing, bring-up).
bool fused_ln_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PC_NO_FUSED_LN");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

void free_folds(Tower& t) {
  if (t.fold_pool) cudaFree(t.fold_pool);
  t.fold_pool = nullptr;
  t.qkv_ln.clear();
  t.fc_ln.clear();
}

// Build the LN-folded copies of in_proj (with ln_1) and c_fc (with ln_2) for every layer.
int build_folds(Tower& t) {
  free_folds(t);
  const size_t d = t.width;
  const size_t w_bytes = align_up((3 * d + 4 * d) * d * sizeof(__half), 256);
  const size_t v_bytes = align_up((3 * d + 4 * d) * sizeof(float), 256);
  const size_t per_layer = w_bytes + 2 * v_bytes;
  PC_CHECK_CUDA(cudaMalloc(&t.fold_pool, per_layer * t.layers));
  t.qkv_ln.resize(t.layers);
  t.fc_ln.resize(t.layers);
  for (int l = 0; l < t.layers; ++l) {
    uint8_t* base = static_cast<uint8_t*>(t.fold_pool) + per_layer * l;
    __half* w = reinterpret_cast<__half*>(base);
    float* sv = reinterpret_cast<float*>(base + w_bytes);
    float* cv = reinterpret_cast<float*>(base + w_bytes + v_bytes);
    t.qkv_ln[l] = FoldedLn{w, sv, cv};
    t.fc_ln[l] = FoldedLn{w + 3 * d * d, sv + 3 * d, cv + 3 * d};
    const pc_resblock_weights& b = t.blocks[l];
    PC_TRY(launch_fold_ln(static_cast<const __half*>(b.in_proj_weight), static_cast<const float*>(b.ln_1_weight),
                          static_cast<const float*>(b.ln_1_bias), static_cast<const __half*>(b.in_proj_bias),
                          t.qkv_ln[l].w, t.qkv_ln[l].s, t.qkv_ln[l].c, static_cast<int>(3 * d), static_cast<int>(d), 0));
    PC_TRY(launch_fold_ln(static_cast<const __half*>(b.c_fc_weight), static_cast<const float*>(b.ln_2_weight),
                          static_cast<const float*>(b.ln_2_bias), static_cast<const __half*>(b.c_fc_bias),
                          t.fc_ln[l].w, t.fc_ln[l].s, t.fc_ln[l].c, static_cast<int>(4 * d), static_cast<int>(d), 0));
  }
  PC_CHECK_CUDA(cudaDeviceSynchronize());
  return PC_OK;
}

// The same block with both LayerNorms folded into the GEMMs that consume them (no LayerNorm launches):
//   in : s1 = `parts_in` partial (sum, sum^2) pairs for every row of x
//   out: s1 = gemm_stats_parts(rows, d) partial pairs of the new x (ready for the next block's ln_1)
// QKV reads x directly with the gamma-scaled in_proj; out_proj writes the statistics of the row segments it
// produces into s2; c_fc reads them (ln_2); c_proj writes the new s1. No atomics: results are bit-reproducible.
// `parts` (measurement only, pc_resblock_forward_parts): which launches run -- PC_PART_QKV | ATTN | OUT | FC | PROJ.
int resblock_fused(const Tower& t, int layer, __half* x, __half* h, __half* big, float* s1, int parts_in, float* s2,
                   int B, int L, int causal, cudaStream_t s, int parts = 31) {
  const pc_resblock_weights& w = t.blocks[layer];
  const int d = t.width, rows = B * L;
  GemmArgs g{};
  g.M = rows; g.N = 3 * d; g.K = d;
  g.A = x; g.lda = d;
  g.W = t.qkv_ln[layer].w; g.ldw = d;
  g.C = big; g.ldc = 3 * d;
  g.ln_stats = s1; g.ln_parts = parts_in; g.ln_s = t.qkv_ln[layer].s; g.ln_c = t.qkv_ln[layer].c;
  // L <= 208: the projection writes q / k / v as 3 * heads contiguous [rows, 64] planes -- an item's Q, K and V are three
  // contiguous 25 KB blocks for the attention kernel's TMA loads instead of 128-byte pieces at a 6 d-byte stride
  const int planar = attention6_supports(L) && planar_qkv_enabled() ? 1 : 0;
  g.c_planar = planar;
  if (parts & 1) PC_TRY(launch_gemm(g, EPI_LN_BIAS, s));
  if (parts & 2) PC_TRY(launch_attention_layout(big, planar, h, B, L, t.heads, causal, s));
  g = GemmArgs{};
  g.M = rows; g.N = d; g.K = d;
  g.A = h; g.lda = d;
  g.W = static_cast<const __half*>(w.out_proj_weight); g.ldw = d;
  g.C = x; g.ldc = d;
  g.bias = static_cast<const __half*>(w.out_proj_bias);
  g.residual = x; g.ldr = d;
  g.stats_out = s2;
  if (parts & 4) PC_TRY(launch_gemm(g, EPI_BIAS_RES, s));
  g = GemmArgs{};
  g.M = rows; g.N = 4 * d; g.K = d;
  g.A = x; g.lda = d;
  g.W = t.fc_ln[layer].w; g.ldw = d;
  g.C = big; g.ldc = 4 * d;
  g.ln_stats = s2; g.ln_parts = gemm_stats_parts(rows, d); g.ln_s = t.fc_ln[layer].s; g.ln_c = t.fc_ln[layer].c;
  if (parts & 8) PC_TRY(launch_gemm(g, EPI_LN_QGELU, s));
  g = GemmArgs{};
  g.M = rows; g.N = d; g.K = 4 * d;
  g.A = big; g.lda = 4 * d;
  g.W = static_cast<const __half*>(w.c_proj_weight); g.ldw = 4 * d;
  g.C = x; g.ldc = d;
  g.bias = static_cast<const __half*>(w.c_proj_bias);
  g.residual = x; g.ldr = d;
  g.stats_out = s1;
  if (parts & 16) PC_TRY(launch_gemm(g, EPI_BIAS_RES, s));
  return PC_OK;
}

// The LAST block of the visual tower, CLS rows only. VisionTransformer.forward keeps x[:, 0, :] after the last block
// (clip/model.py:232-233): of that block only K / V of every token and the CLS row of everything else can reach the
// output. QKV projection for all tokens (the folded GEMM as usual; its q columns of the other tokens are the only
// unused work left), attention of the CLS query alone, then out_proj + residual, ln_2 (folded), c_fc, QuickGELU,
// c_proj + residual on B rows instead of B * L: 18 (L - 1) d^2 + 4 L (L - 1) d of the block's 24 L d^2 + 4 L^2 d flop
// never run (6.3 % of a ViT-B/16 tower). Per-row arithmetic is the towers' own (same GEMM kernel and epilogues; the
// single-row attention keeps P in fp32). xc_out: compact [B, d] rows of the block output (in `h`).
// PC_FULL_LAST_BLOCK=1 runs the whole block like the reference (A/B, and the figure bench.py reports beside).
bool cls_only_last_block_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PC_FULL_LAST_BLOCK");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}
int resblock_cls_only(const Tower& t, int layer, __half* x, __half* h, __half* big, float* s1, int parts_in, float* s2,
                      int B, int L, __half** xc_out, cudaStream_t s) {
  const pc_resblock_weights& w = t.blocks[layer];
  const int d = t.width, rows = B * L;
  GemmArgs g{};
  g.M = rows; g.N = 3 * d; g.K = d;
  g.A = x; g.lda = d;
  g.W = t.qkv_ln[layer].w; g.ldw = d;
  g.C = big; g.ldc = 3 * d;
  g.ln_stats = s1; g.ln_parts = parts_in; g.ln_s = t.qkv_ln[layer].s; g.ln_c = t.qkv_ln[layer].c;
  const int planar = attention6_supports(L) && planar_qkv_enabled() ? 1 : 0;
  g.c_planar = planar;
  PC_TRY(launch_gemm(g, EPI_LN_BIAS, s));
  __half* attn = h;                                // [B, d]  CLS rows of the attention output
  __half* xc = h + static_cast<size_t>(B) * d;     // [B, d]  CLS rows of the residual stream
  __half* hid = big + static_cast<size_t>(rows) * 3 * d;  // [B, 4d] behind qkv (rows * d elements are free there)
  PC_TRY(launch_attention_rows(big, planar, attn, B, L, t.heads, 0, 1, 0, s));
  g = GemmArgs{};
  g.M = B; g.N = d; g.K = d;
  g.A = attn; g.lda = d;
  g.W = static_cast<const __half*>(w.out_proj_weight); g.ldw = d;
  g.C = xc; g.ldc = d;
  g.bias = static_cast<const __half*>(w.out_proj_bias);
  g.residual = x; g.ldr = L * d;  // row b of the residual = the CLS row of sequence b
  g.stats_out = s2;
  PC_TRY(launch_gemm(g, EPI_BIAS_RES, s));
  g = GemmArgs{};
  g.M = B; g.N = 4 * d; g.K = d;
  g.A = xc; g.lda = d;
  g.W = t.fc_ln[layer].w; g.ldw = d;
  g.C = hid; g.ldc = 4 * d;
  g.ln_stats = s2; g.ln_parts = gemm_stats_parts(B, d); g.ln_s = t.fc_ln[layer].s; g.ln_c = t.fc_ln[layer].c;
  PC_TRY(launch_gemm(g, EPI_LN_QGELU, s));
  g = GemmArgs{};
  g.M = B; g.N = d; g.K = 4 * d;
  g.A = hid; g.lda = 4 * d;
  g.W = static_cast<const __half*>(w.c_proj_weight); g.ldw = 4 * d;
  g.C = xc; g.ldc = d;
  g.bias = static_cast<const __half*>(w.c_proj_bias);
  g.residual = xc; g.ldr = d;
  PC_TRY(launch_gemm(g, EPI_BIAS_RES, s));
  *xc_out = xc;
  return PC_OK;
}

// workspace carve-up shared by both towers: x [rows,d] | h [rows,d] | big [rows,4d] | idx [mb] ints
struct TowerWs {
  __half *x, *h, *big;
  int* idx;
  float *s1, *s2;  // LayerNorm statistics [rows][parts][2] (ln_1 / ln_2 inputs)
};
// partial pairs per row: gemm_stats_parts() = 2 per column tile of the residual GEMM; sized for the worst case, the
// 128-column tiles of the single-CTA kernel (2 * ceil(d / 128); d = 64 -> 2 pairs, not d / 64 = 1)
size_t stats_bytes(int rows, int d) {
  return align_up(static_cast<size_t>(rows) * (2 * ((d + 127) / 128)) * 8, 256);
}
size_t tower_ws_bytes(int rows, int d, int mb) {
  return align_up(static_cast<size_t>(rows) * d * 2, 256) * 2 + align_up(static_cast<size_t>(rows) * d * 8, 256) +
         align_up(static_cast<size_t>(mb) * 4, 256) + 2 * stats_bytes(rows, d);
}
TowerWs carve(void* ws, int rows, int d, int mb) {
  TowerWs t;
  uint8_t* p = static_cast<uint8_t*>(ws);
  const size_t a = align_up(static_cast<size_t>(rows) * d * 2, 256);
  t.x = reinterpret_cast<__half*>(p);
  t.h = reinterpret_cast<__half*>(p + a);
  t.big = reinterpret_cast<__half*>(p + 2 * a);
  uint8_t* q = p + 2 * a + align_up(static_cast<size_t>(rows) * d * 8, 256);
  t.idx = reinterpret_cast<int*>(q);
  q += align_up(static_cast<size_t>(mb) * 4, 256);
  t.s1 = reinterpret_cast<float*>(q);
  t.s2 = reinterpret_cast<float*>(q + stats_bytes(rows, d));
  return t;
}


// ------------------------------------------------------------------------------------------ ModifiedResNet
void free_rn(RnTower& t) {
  for (void* p : t.owned) cudaFree(p);
  t = RnTower{};
}

int fold_one(RnTower& t, const pc_conv_bn_weights& src, int cout, int cin, int k, ConvBn* out) {
  PC_REQUIRE(src.conv_weight && src.bn_weight && src.bn_bias && src.bn_running_mean && src.bn_running_var, PC_ERR_ARG,
             "pc_rn_bind_weights: null conv / BatchNorm tensor (Cout %d, Cin %d, k %d)", cout, cin, k);
  PC_REQUIRE(cout % 8 == 0 && (k == 3 || cin % 8 == 0), PC_ERR_ARG,
             "pc_rn_bind_weights: channel counts must be multiples of 8 (Cout %d, Cin %d)", cout, cin);
  out->cout = cout; out->cin = cin; out->k = k;
  out->Kp = static_cast<int>(align_up(static_cast<size_t>(cin) * k * k, 8));
  PC_CHECK_CUDA(cudaMalloc(&out->w, static_cast<size_t>(cout) * out->Kp * sizeof(__half)));
  t.owned.push_back(out->w);
  PC_CHECK_CUDA(cudaMalloc(&out->shift, static_cast<size_t>(cout) * sizeof(float)));
  t.owned.push_back(out->shift);
  return launch_fold_conv_bn(static_cast<const __half*>(src.conv_weight), static_cast<const float*>(src.bn_weight),
                             static_cast<const float*>(src.bn_bias), static_cast<const float*>(src.bn_running_mean),
                             static_cast<const float*>(src.bn_running_var), out->w, out->shift, cout, cin, k, out->Kp, 0);
}

// conv (+ folded BN) [+ identity] [+ ReLU] of `rows` pixels: a [rows, c.Kp] x c.w [cout, Kp]^T -> out [rows, cout]
int conv_gemm(const ConvBn& c, const __half* a, __half* out, int rows, const __half* identity, int relu, cudaStream_t s) {
  GemmArgs g{};
  g.M = rows; g.N = c.cout; g.K = c.Kp;
  g.A = a; g.lda = c.Kp;
  g.W = c.w; g.ldw = c.Kp;
  g.C = out; g.ldc = c.cout;
  g.bias_f32 = c.shift;
  g.relu = relu;
  if (identity) {
    g.residual = identity; g.ldr = c.cout;
    return launch_gemm(g, EPI_BIAS_RES, s);
  }
  return launch_gemm(g, EPI_BIAS, s);
}

struct RnWs {
  __half* buf[5];  // activation buffers of max_act * mb elements each
  __half* col;     // stem im2col operand / bordered 3x3 input, max_col * mb elements
};
size_t rn_ws_bytes(const RnTower& t, int mb) {
  return 5 * align_up(t.max_act * mb * 2, 256) + align_up(t.max_col * mb * 2, 256);
}
RnWs rn_carve(const RnTower& t, void* ws, int mb) {
  RnWs r;
  uint8_t* p = static_cast<uint8_t*>(ws);
  const size_t a = align_up(t.max_act * mb * 2, 256);
  for (int i = 0; i < 5; ++i) r.buf[i] = reinterpret_cast<__half*>(p + i * a);
  r.col = reinterpret_cast<__half*>(p + 5 * a);
  return r;
}

// 3x3 / pad 1 conv (+ folded BN + ReLU) as an implicit GEMM: xp and out are bordered [n, h+2, w+2, C] tensors
int conv3x3_gemm(const ConvBn& c, const __half* xp, __half* out, int n, int h, cudaStream_t s) {
  GemmArgs g{};
  g.M = n * (h + 2) * (h + 2); g.N = c.cout; g.K = c.cin;
  g.A = xp; g.lda = c.cin;
  g.W = c.w; g.ldw = c.Kp;
  g.C = out; g.ldc = c.cout;
  g.bias_f32 = c.shift;
  g.relu = 1;
  g.conv_taps = 9;
  g.conv_pitch = h + 2;
  return launch_gemm(g, EPI_BIAS, s);
}

// the same on plain NHWC tensors x [n, h, h, cin] -> out [n, h, h, cout] (GemmArgs patch mode; conv_patch_supported(h, h))
int conv3x3_patch_gemm(const ConvBn& c, const __half* x, __half* out, int n, int h, cudaStream_t s) {
  GemmArgs g{};
  g.N = c.cout; g.K = c.cin;
  g.A = x; g.lda = c.cin;
  g.W = c.w; g.ldw = c.Kp;
  g.C = out; g.ldc = c.cout;
  g.bias_f32 = c.shift;
  g.relu = 1;
  g.conv_taps = 9;
  g.conv_h = h; g.conv_w = h; g.conv_n = n;
  return launch_gemm(g, EPI_BIAS, s);
}

// largest activation / staging operand per image along the forward below
void rn_plan(RnTower& t) {
  size_t act = 0, col = 0;
  auto A = [&](size_t v) { if (v > act) act = v; };
  auto C = [&](size_t v) { if (v > col) col = v; };
  size_t h = t.res / 2;
  C(h * h * 32); A(h * h * t.stem[0].cout);
  C((h + 2) * (h + 2) * t.stem[1].cin); A((h + 2) * (h + 2) * t.stem[1].cout);
  A((h + 2) * (h + 2) * t.stem[2].cout);
  h /= 2;
  for (const RnBlock& b : t.blocks) {
    A(h * h * b.planes);
    C((h + 2) * (h + 2) * b.planes);
    A((h + 2) * (h + 2) * b.planes);
    const size_t ho = h / b.stride;
    A(ho * ho * b.planes * 4);
    A(ho * ho * b.inplanes);
    h = ho;
  }
  A(static_cast<size_t>(t.tokens) * 3 * t.embed);
  t.max_act = act;
  t.max_col = col;
}

// ModifiedResNet.forward (clip/model.py:137-152) on n images -> feat [n, out_dim]
int rn_forward(const RnTower& t, const void* img, int img_is_f16, int n, __half* feat, const RnWs& ws, cudaStream_t s) {
  __half *x = ws.buf[0], *y = ws.buf[1], *t1 = ws.buf[2], *t2 = ws.buf[3], *t3 = ws.buf[4];
  int h = t.res / 2;
  // stem (:139-141): three conv + BN + ReLU, then avgpool(2). conv1 (3 channels, stride 2) goes through a 32-column
  // im2col; conv2 / conv3 are implicit GEMMs chained in the bordered layout (frame re-zeroed in between)
  PC_TRY(launch_stem_im2col(img, img_is_f16, ws.col, n, t.res, s));
  PC_TRY(conv_gemm(t.stem[0], ws.col, t1, n * h * h, nullptr, 1, s));
  if (conv_patch_supported(h, h)) {  // 3x3 convolutions straight on the NHWC tensors (padding = TMA zero fill)
    PC_TRY(conv3x3_patch_gemm(t.stem[1], t1, t2, n, h, s));
    PC_TRY(conv3x3_patch_gemm(t.stem[2], t2, t1, n, h, s));
    PC_TRY(launch_avgpool_nhwc(t1, x, n, h, h, t.stem[2].cout, 2, 0, s));
  } else {
    PC_TRY(launch_pad_nhwc(t1, ws.col, n, h, h, t.stem[1].cin, s));
    PC_TRY(conv3x3_gemm(t.stem[1], ws.col, t2, n, h, s));
    PC_TRY(launch_zero_border(t2, n, h, h, t.stem[2].cin, s));
    PC_TRY(conv3x3_gemm(t.stem[2], t2, t1, n, h, s));
    PC_TRY(launch_avgpool_nhwc(t1, x, n, h, h, t.stem[2].cout, 2, 1, s));
  }
  h /= 2;
  // layer1..layer4 (:146-149), Bottleneck.forward (:40-53)
  for (const RnBlock& b : t.blocks) {
    const int ho = h / b.stride;
    PC_TRY(conv_gemm(b.c1, x, t1, n * h * h, nullptr, 1, s));
    const __half* c3_in = t1;
    if (conv_patch_supported(h, h)) {
      PC_TRY(conv3x3_patch_gemm(b.c2, t1, t2, n, h, s));
      if (b.stride > 1) PC_TRY(launch_avgpool_nhwc(t2, t1, n, h, h, b.planes, b.stride, 0, s));  // anti-aliasing (:45)
      else c3_in = t2;
    } else {
      PC_TRY(launch_pad_nhwc(t1, ws.col, n, h, h, b.planes, s));
      PC_TRY(conv3x3_gemm(b.c2, ws.col, t2, n, h, s));
      // interior of the bordered conv2 output -> t1, through the anti-aliasing avgpool when the block strides (:45)
      if (b.stride > 1) PC_TRY(launch_avgpool_nhwc(t2, t1, n, h, h, b.planes, b.stride, 1, s));
      else PC_TRY(launch_unpad_nhwc(t2, t1, n, h, h, b.planes, s));
    }
    const __half* identity = x;
    if (b.has_down) {  // :34-38, :48-49
      const __half* src = x;
      if (b.stride > 1) {
        PC_TRY(launch_avgpool_nhwc(x, t3, n, h, h, b.inplanes, b.stride, 0, s));
        src = t3;
      }
      __half* down_out = b.stride > 1 ? t2 : t3;  // t2 may hold conv2's result (c3_in), t3 the pooled x
      PC_TRY(conv_gemm(b.down, src, down_out, n * ho * ho, nullptr, 0, s));
      identity = down_out;
    }
    PC_TRY(conv_gemm(b.c3, c3_in, y, n * ho * ho, identity, 1, s));
    __half* sw = x; x = y; y = sw;
    h = ho;
  }
  // AttentionPool2d (:67-92) returns token 0 only: k / v projections of every token, the q projection and the
  // attention of the mean token alone (the reference computes all HW + 1 query rows and drops the rest), then c_proj
  const int E = t.embed, L = t.tokens;
  PC_TRY(launch_attnpool_tokens(x, t.pos, t1, n, h * h, E, s));
  GemmArgs g{};
  g.M = n * L; g.N = 2 * E; g.K = E;
  g.A = t1; g.lda = E;
  g.W = t.qkv_w + static_cast<size_t>(E) * E; g.ldw = E;  // rows [E, 3E): k_proj, v_proj
  g.C = t2; g.ldc = 2 * E;
  g.bias = t.qkv_b + E;
  PC_TRY(launch_gemm(g, EPI_BIAS, s));
  __half* q0 = t3;                                   // [n, E]
  __half* pooled = t3 + static_cast<size_t>(n) * E;  // [n, E]
  g = GemmArgs{};
  g.M = n; g.N = E; g.K = E;
  g.A = t1; g.lda = L * E;  // token 0 of every image
  g.W = t.qkv_w; g.ldw = E;
  g.C = q0; g.ldc = E;
  g.bias = t.qkv_b;
  PC_TRY(launch_gemm(g, EPI_BIAS, s));
  PC_TRY(launch_attnpool_query0(q0, t2, pooled, n, L, t.heads, s));
  g = GemmArgs{};
  g.M = n; g.N = t.out_dim; g.K = E;
  g.A = pooled; g.lda = E;
  g.W = t.c_w; g.ldw = E;
  g.C = feat; g.ldc = t.out_dim;
  g.bias = t.c_b;
  return launch_gemm(g, EPI_BIAS, s);
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int pc_version(void) { return PC_VERSION; }
const char* pc_last_error(void) { return get_error(); }

int pc_ctx_create(int device, pc_ctx** out) {
  PC_REQUIRE(out != nullptr, PC_ERR_ARG, "pc_ctx_create: out is null");
  *out = nullptr;
  int count = 0;
  PC_CHECK_CUDA(cudaGetDeviceCount(&count));
  PC_REQUIRE(device >= 0 && device < count, PC_ERR_ARG, "pc_ctx_create: device %d out of range (%d visible)",
             device, count);
  cudaDeviceProp prop;
  PC_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  PC_REQUIRE(prop.major == 10, PC_ERR_ARCH,
             "pc_ctx_create: device %d is sm_%d%d; libprotoclip_b200 is sm_100a-only and has no fallback",
             device, prop.major, prop.minor);
  PC_CHECK_CUDA(cudaSetDevice(device));
  pc_ctx* c = new (std::nothrow) pc_ctx();
  PC_REQUIRE(c != nullptr, PC_ERR_CUDA, "pc_ctx_create: out of host memory");
  c->device = device;
  *out = c;
  return PC_OK;
}

int pc_ctx_set_full_last_block(pc_ctx* ctx, int full) {
  PC_REQUIRE(ctx != nullptr, PC_ERR_ARG, "pc_ctx_set_full_last_block: null context");
  ctx->full_last_block = full < 0 ? -1 : (full ? 1 : 0);
  return PC_OK;
}

void pc_ctx_destroy(pc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->vis.proj_t) cudaFree(ctx->vis.proj_t);
  if (ctx->txt.proj_t) cudaFree(ctx->txt.proj_t);
  free_folds(ctx->vis);
  free_folds(ctx->txt);
  if (ctx->conv1_p) cudaFree(ctx->conv1_p);
  free_rn(ctx->rn);
  delete ctx;
}

int pc_vit_bind_weights(pc_ctx* ctx, const pc_vit_weights* w) {
  PC_TRY(use_device(ctx));
  PC_REQUIRE(w != nullptr, PC_ERR_ARG, "pc_vit_bind_weights: weights are null");
  PC_REQUIRE(w->width > 0 && w->width % 64 == 0 && w->heads * 64 == w->width, PC_ERR_ARG,
             "pc_vit_bind_weights: width %d / heads %d (head_dim must be 64)", w->width, w->heads);
  PC_REQUIRE(w->patch_size > 0 && w->patch_size % 2 == 0 && w->image_resolution % w->patch_size == 0,
             PC_ERR_ARG, "pc_vit_bind_weights: resolution %d / patch %d", w->image_resolution, w->patch_size);
  PC_REQUIRE(w->layers > 0 && w->embed_dim > 0 && w->embed_dim % 8 == 0, PC_ERR_ARG,
             "pc_vit_bind_weights: layers %d embed_dim %d", w->layers, w->embed_dim);
  PC_REQUIRE(w->conv1_weight && w->class_embedding && w->positional_embedding && w->ln_pre_weight &&
                 w->ln_pre_bias && w->ln_post_weight && w->ln_post_bias && w->proj,
             PC_ERR_ARG, "pc_vit_bind_weights: null stem tensor");
  PC_TRY(check_blocks(w->blocks, w->layers));
  ctx->vis.bound = false;
  free_rn(ctx->rn);
  ctx->res = w->image_resolution;
  ctx->patch = w->patch_size;
  ctx->grid = ctx->res / ctx->patch;
  ctx->Kpatch = 3 * ctx->patch * ctx->patch;
  ctx->Kp = static_cast<int>(align_up(ctx->Kpatch, 8));
  Tower& t = ctx->vis;
  t.width = w->width; t.layers = w->layers; t.heads = w->heads; t.embed = w->embed_dim;
  t.L = ctx->grid * ctx->grid + 1;
  t.blocks.assign(w->blocks, w->blocks + w->layers);
  ctx->cls = static_cast<const float*>(w->class_embedding);
  ctx->vpos = static_cast<const float*>(w->positional_embedding);
  ctx->ln_pre_w = static_cast<const float*>(w->ln_pre_weight);
  ctx->ln_pre_b = static_cast<const float*>(w->ln_pre_bias);
  ctx->ln_post_w = static_cast<const float*>(w->ln_post_weight);
  ctx->ln_post_b = static_cast<const float*>(w->ln_post_bias);
  // conv1.weight [d, 3, p, p] is already the [N, K] GEMM operand; re-pitch rows to Kp (16-byte TMA rows).
  if (ctx->conv1_p) cudaFree(ctx->conv1_p);
  PC_CHECK_CUDA(cudaMalloc(&ctx->conv1_p, static_cast<size_t>(t.width) * ctx->Kp * 2));
  PC_CHECK_CUDA(cudaMemset(ctx->conv1_p, 0, static_cast<size_t>(t.width) * ctx->Kp * 2));
  PC_CHECK_CUDA(cudaMemcpy2D(ctx->conv1_p, static_cast<size_t>(ctx->Kp) * 2, w->conv1_weight,
                             static_cast<size_t>(ctx->Kpatch) * 2, static_cast<size_t>(ctx->Kpatch) * 2, t.width,
                             cudaMemcpyDeviceToDevice));
  PC_TRY(make_transposed(w->proj, t.width, t.embed, &t.proj_t));
  PC_TRY(build_folds(t));
  t.bound = true;
  return PC_OK;
}

int pc_rn_bind_weights(pc_ctx* ctx, const pc_rn_weights* w) {
  PC_TRY(use_device(ctx));
  PC_REQUIRE(w != nullptr && w->blocks != nullptr, PC_ERR_ARG, "pc_rn_bind_weights: weights are null");
  PC_REQUIRE(w->width >= 16 && w->width % 16 == 0 && w->heads * 64 == w->width * 32, PC_ERR_ARG,
             "pc_rn_bind_weights: width %d / heads %d (width must be a multiple of 16, head_dim 64)", w->width, w->heads);
  PC_REQUIRE(w->image_resolution > 0 && w->image_resolution % 32 == 0 && w->output_dim > 0 && w->output_dim % 8 == 0,
             PC_ERR_ARG, "pc_rn_bind_weights: resolution %d / output_dim %d", w->image_resolution, w->output_dim);
  PC_REQUIRE(w->attnpool_positional_embedding && w->q_proj_weight && w->q_proj_bias && w->k_proj_weight &&
                 w->k_proj_bias && w->v_proj_weight && w->v_proj_bias && w->c_proj_weight && w->c_proj_bias,
             PC_ERR_ARG, "pc_rn_bind_weights: null attention-pool tensor");
  free_rn(ctx->rn);
  ctx->vis.bound = false;
  RnTower& t = ctx->rn;
  t.res = w->image_resolution; t.width = w->width; t.heads = w->heads; t.embed = w->width * 32;
  t.out_dim = w->output_dim;
  const int g = t.res / 32;
  t.tokens = g * g + 1;
  const int half_w = w->width / 2;
  int rc = fold_one(t, w->stem[0], half_w, 3, 3, &t.stem[0]);
  if (rc == PC_OK) rc = fold_one(t, w->stem[1], half_w, half_w, 3, &t.stem[1]);
  if (rc == PC_OK) rc = fold_one(t, w->stem[2], w->width, half_w, 3, &t.stem[2]);
  // the stem im2col writes 32 columns (27 taps x channels + padding)
  if (rc == PC_OK && t.stem[0].Kp != 32) { set_error("pc_rn_bind_weights: stem K %d", t.stem[0].Kp); rc = PC_ERR_ARG; }
  int inpl = w->width, nb = 0;
  for (int li = 0; li < 4 && rc == PC_OK; ++li) {
    if (w->layers[li] <= 0) { set_error("pc_rn_bind_weights: layer%d has %d blocks", li + 1, w->layers[li]); rc = PC_ERR_ARG; }
    for (int b = 0; b < w->layers[li] && rc == PC_OK; ++b, ++nb) {
      const pc_bottleneck_weights& src = w->blocks[nb];
      const int planes = w->width << li, stride = (li > 0 && b == 0) ? 2 : 1;
      if (src.inplanes != inpl || src.planes != planes || src.stride != stride) {
        set_error("pc_rn_bind_weights: block %d is (%d, %d, stride %d), expected (%d, %d, stride %d)", nb, src.inplanes,
                  src.planes, src.stride, inpl, planes, stride);
        rc = PC_ERR_ARG;
        break;
      }
      RnBlock blk;
      blk.inplanes = inpl; blk.planes = planes; blk.stride = stride;
      rc = fold_one(t, src.conv1, planes, inpl, 1, &blk.c1);
      if (rc == PC_OK) rc = fold_one(t, src.conv2, planes, planes, 3, &blk.c2);
      if (rc == PC_OK) rc = fold_one(t, src.conv3, planes * 4, planes, 1, &blk.c3);
      blk.has_down = stride > 1 || inpl != planes * 4;
      if (rc == PC_OK && blk.has_down) rc = fold_one(t, src.downsample, planes * 4, inpl, 1, &blk.down);
      t.blocks.push_back(blk);
      inpl = planes * 4;
    }
  }
  if (rc == PC_OK && inpl != t.embed) { set_error("pc_rn_bind_weights: trunk width %d != 32 * width", inpl); rc = PC_ERR_ARG; }
  if (rc == PC_OK) {
    // packed [q; k; v] in-projection: the attention kernel's qkv column order
    const size_t E = t.embed, wb = E * E * sizeof(__half), bb = E * sizeof(__half);
    cudaError_t e = cudaMalloc(&t.qkv_w, 3 * wb);
    if (e == cudaSuccess) { t.owned.push_back(t.qkv_w); e = cudaMalloc(&t.qkv_b, 3 * bb); }
    if (e == cudaSuccess) {
      t.owned.push_back(t.qkv_b);
      const void* ws3[3] = {w->q_proj_weight, w->k_proj_weight, w->v_proj_weight};
      const void* bs3[3] = {w->q_proj_bias, w->k_proj_bias, w->v_proj_bias};
      for (int i = 0; i < 3 && e == cudaSuccess; ++i) {
        e = cudaMemcpy(t.qkv_w + i * E * E, ws3[i], wb, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(t.qkv_b + i * E, bs3[i], bb, cudaMemcpyDeviceToDevice);
      }
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { set_error("pc_rn_bind_weights: %s", cudaGetErrorString(e)); rc = PC_ERR_CUDA; }
  }
  if (rc != PC_OK) {
    free_rn(ctx->rn);
    return rc;
  }
  t.pos = static_cast<const float*>(w->attnpool_positional_embedding);
  t.c_w = static_cast<const __half*>(w->c_proj_weight);
  t.c_b = static_cast<const __half*>(w->c_proj_bias);
  rn_plan(t);
  t.bound = true;
  return PC_OK;
}

int pc_text_bind_weights(pc_ctx* ctx, const pc_text_weights* w) {
  PC_TRY(use_device(ctx));
  PC_REQUIRE(w != nullptr, PC_ERR_ARG, "pc_text_bind_weights: weights are null");
  PC_REQUIRE(w->width > 0 && w->width % 64 == 0 && w->heads * 64 == w->width, PC_ERR_ARG,
             "pc_text_bind_weights: width %d / heads %d (head_dim must be 64)", w->width, w->heads);
  PC_REQUIRE(w->layers > 0 && w->context_length > 0 && w->vocab_size > 0 && w->embed_dim % 8 == 0, PC_ERR_ARG,
             "pc_text_bind_weights: bad descriptor");
  PC_REQUIRE(w->token_embedding && w->positional_embedding && w->ln_final_weight && w->ln_final_bias &&
                 w->text_projection,
             PC_ERR_ARG, "pc_text_bind_weights: null stem tensor");
  PC_TRY(check_blocks(w->blocks, w->layers));
  Tower& t = ctx->txt;
  t.bound = false;
  t.width = w->width; t.layers = w->layers; t.heads = w->heads; t.embed = w->embed_dim;
  t.L = w->context_length;
  t.blocks.assign(w->blocks, w->blocks + w->layers);
  ctx->vocab = w->vocab_size;
  ctx->tok_emb = static_cast<const float*>(w->token_embedding);
  ctx->tpos = static_cast<const float*>(w->positional_embedding);
  ctx->ln_final_w = static_cast<const float*>(w->ln_final_weight);
  ctx->ln_final_b = static_cast<const float*>(w->ln_final_bias);
  PC_TRY(make_transposed(w->text_projection, t.width, t.embed, &t.proj_t));
  PC_TRY(build_folds(t));
  t.bound = true;
  return PC_OK;
}

size_t pc_encode_image_workspace_bytes(const pc_ctx* ctx, int micro_batch) {
  if (ctx && ctx->rn.bound) return rn_ws_bytes(ctx->rn, micro_batch > 0 ? micro_batch : kRnMicroBatch);
  if (!ctx || !ctx->vis.bound) return 0;
  const int mb = micro_batch > 0 ? micro_batch : default_micro_batch(ctx->vis.L);
  return tower_ws_bytes(mb * ctx->vis.L, ctx->vis.width, mb);
}

int pc_encode_image(pc_ctx* ctx, const void* images, int img_dtype, int B, void* feat_out, int l2norm,
                    int micro_batch, void* workspace, size_t workspace_bytes, void* stream) {
  PC_TRY(use_device(ctx));
  PC_REQUIRE(ctx->vis.bound || ctx->rn.bound, PC_ERR_STATE, "pc_encode_image: visual weights are not bound");
  PC_REQUIRE(images && feat_out && B > 0, PC_ERR_ARG, "pc_encode_image: null buffer or empty batch");
  PC_REQUIRE(img_dtype == PC_IMG_F32 || img_dtype == PC_IMG_F16, PC_ERR_ARG, "pc_encode_image: image dtype %d",
             img_dtype);
  if (ctx->rn.bound) {  // ModifiedResNet tower (clip/model.py:137-152)
    const RnTower& r = ctx->rn;
    const int rmb = pick_rn_micro_batch(B, micro_batch);
    PC_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PC_ERR_ALIGN,
               "pc_encode_image: workspace must be 256-byte aligned");
    PC_REQUIRE(workspace_bytes >= rn_ws_bytes(r, rmb), PC_ERR_WORKSPACE, "pc_encode_image: workspace %zu < %zu",
               workspace_bytes, rn_ws_bytes(r, rmb));
    cudaStream_t rs = static_cast<cudaStream_t>(stream);
    const RnWs rws = rn_carve(r, workspace, rmb);
    const size_t ib = static_cast<size_t>(3) * r.res * r.res * (img_dtype == PC_IMG_F16 ? 2 : 4);
    __half* rf = static_cast<__half*>(feat_out);
    for (int b0 = 0; b0 < B; b0 += rmb) {
      const int n = (B - b0 < rmb) ? (B - b0) : rmb;
      __half* f = rf + static_cast<size_t>(b0) * r.out_dim;
      PC_TRY(rn_forward(r, static_cast<const uint8_t*>(images) + static_cast<size_t>(b0) * ib,
                        img_dtype == PC_IMG_F16, n, f, rws, rs));
      if (l2norm) PC_TRY(launch_l2norm(f, f, n, r.out_dim, rs));
    }
    return PC_OK;
  }
  const Tower& t = ctx->vis;
  const int mb = pick_micro_batch(t.L, B, micro_batch);
  PC_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PC_ERR_ALIGN,
             "pc_encode_image: workspace must be 256-byte aligned");
  PC_REQUIRE(workspace_bytes >= tower_ws_bytes(mb * t.L, t.width, mb), PC_ERR_WORKSPACE,
             "pc_encode_image: workspace %zu < %zu", workspace_bytes, tower_ws_bytes(mb * t.L, t.width, mb));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const TowerWs ws = carve(workspace, mb * t.L, t.width, mb);
  const int d = t.width, g2 = ctx->grid * ctx->grid;
  const size_t img_elems = static_cast<size_t>(3) * ctx->res * ctx->res;
  const size_t img_bytes = img_elems * (img_dtype == PC_IMG_F16 ? 2 : 4);
  __half* feat = static_cast<__half*>(feat_out);
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int n = (B - b0 < mb) ? (B - b0) : mb;
    const void* img = static_cast<const uint8_t*>(images) + static_cast<size_t>(b0) * img_bytes;
    // conv1 as a GEMM: patches [n*g2, Kp] (in `big`) x conv1_p [d, Kp]^T -> patch tokens [n*g2, d] (in `h`)
    PC_TRY(launch_patchify(img, img_dtype == PC_IMG_F16, ws.big, n, ctx->res, ctx->patch, ctx->Kp, s));
    GemmArgs g{};
    g.M = n * g2; g.N = d; g.K = ctx->Kp;
    g.A = ws.big; g.lda = ctx->Kp;
    g.W = ctx->conv1_p; g.ldw = ctx->Kp;
    g.C = ws.h; g.ldc = d;
    PC_TRY(launch_gemm(g, EPI_BIAS, s));
    PC_TRY(launch_embed_ln_pre(ws.h, ctx->cls, ctx->vpos, ctx->ln_pre_w, ctx->ln_pre_b, ws.x,
                               fused_ln_enabled() ? ws.s1 : nullptr, n, t.L, d, s));
    __half* cls_rows = nullptr;  // compact CLS rows when the last block ran on them alone
    if (fused_ln_enabled()) {
      const bool want_cls_only = ctx->full_last_block < 0 ? cls_only_last_block_enabled() : ctx->full_last_block == 0;
      const bool cls_only = want_cls_only && t.L >= 4 && t.L <= 1024;
      const int full_layers = cls_only ? t.layers - 1 : t.layers;
      for (int l = 0; l < full_layers; ++l)
        PC_TRY(resblock_fused(t, l, ws.x, ws.h, ws.big, ws.s1, l == 0 ? 1 : gemm_stats_parts(n * t.L, d), ws.s2, n, t.L, 0, s));
      if (cls_only)
        PC_TRY(resblock_cls_only(t, t.layers - 1, ws.x, ws.h, ws.big, ws.s1,
                                 t.layers == 1 ? 1 : gemm_stats_parts(n * t.L, d), ws.s2, n, t.L, &cls_rows, s));
    } else {
      for (int l = 0; l < t.layers; ++l) PC_TRY(resblock(t, l, ws.x, ws.h, ws.big, n, t.L, 0, s));
    }
    // ln_post on the CLS rows, then @ proj (clip/model.py:233-236)
    __half* post = cls_rows ? cls_rows + static_cast<size_t>(n) * d : ws.h;  // h: [attn | xc | ln_post(xc)] or [ln_post]
    if (cls_rows) PC_TRY(launch_layernorm(cls_rows, post, ctx->ln_post_w, ctx->ln_post_b, n, d, 1, s));
    else PC_TRY(launch_layernorm(ws.x, post, ctx->ln_post_w, ctx->ln_post_b, n, d, t.L, s));
    g = GemmArgs{};
    g.M = n; g.N = t.embed; g.K = d;
    g.A = post; g.lda = d;
    g.W = t.proj_t; g.ldw = d;
    g.C = feat + static_cast<size_t>(b0) * t.embed; g.ldc = t.embed;
    PC_TRY(launch_gemm(g, EPI_BIAS, s));
    if (l2norm) {
      __half* f = feat + static_cast<size_t>(b0) * t.embed;
      PC_TRY(launch_l2norm(f, f, n, t.embed, s));
    }
  }
  return PC_OK;
}

size_t pc_preprocess_workspace_bytes(int H, int W, int n_px) { return preprocess_workspace_bytes(1, H, W, n_px); }
size_t pc_preprocess_batch_workspace_bytes(int B, int H, int W, int n_px) { return preprocess_workspace_bytes(B, H, W, n_px); }

int pc_preprocess_batch(const uint8_t* rgb, int B, int H, int W, int n_px, void* out, int out_dtype, void* workspace,
                        size_t workspace_bytes, void* stream) {
  PC_REQUIRE(out_dtype == PC_IMG_F32 || out_dtype == PC_IMG_F16, PC_ERR_ARG, "pc_preprocess_batch: out dtype %d", out_dtype);
  PC_REQUIRE(workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PC_ERR_ALIGN,
             "pc_preprocess_batch: workspace must be 256-byte aligned");
  return launch_preprocess(rgb, B, H, W, n_px, out, out_dtype == PC_IMG_F16, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

int pc_preprocess_image(const uint8_t* rgb, int H, int W, int n_px, void* out, int out_dtype, void* workspace,
                        size_t workspace_bytes, void* stream) {
  return pc_preprocess_batch(rgb, 1, H, W, n_px, out, out_dtype, workspace, workspace_bytes, stream);
}

size_t pc_preprocess_train_workspace_bytes(int crop_h, int crop_w, int n_px) {
  return preprocess_train_workspace_bytes(crop_h, crop_w, n_px);
}

int pc_preprocess_train_image(const uint8_t* rgb, int H, int W, int top, int left, int crop_h, int crop_w, int flip,
                              int n_px, void* out, int out_dtype, void* workspace, size_t workspace_bytes, void* stream) {
  PC_REQUIRE(out_dtype == PC_IMG_F32 || out_dtype == PC_IMG_F16, PC_ERR_ARG, "pc_preprocess_train_image: out dtype %d",
             out_dtype);
  PC_REQUIRE(workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PC_ERR_ALIGN,
             "pc_preprocess_train_image: workspace must be 256-byte aligned");
  return launch_preprocess_train(rgb, H, W, top, left, crop_h, crop_w, flip, n_px, out, out_dtype == PC_IMG_F16,
                                 workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t pc_encode_text_workspace_bytes(const pc_ctx* ctx, int micro_batch) {
  if (!ctx || !ctx->txt.bound) return 0;
  const int mb = micro_batch > 0 ? micro_batch : default_micro_batch(ctx->txt.L);
  return tower_ws_bytes(mb * ctx->txt.L, ctx->txt.width, mb);
}

int pc_encode_text(pc_ctx* ctx, const int64_t* tokens, int P, void* out, int l2norm, int micro_batch,
                   void* workspace, size_t workspace_bytes, void* stream) {
  PC_TRY(use_device(ctx));
  PC_REQUIRE(ctx->txt.bound, PC_ERR_STATE, "pc_encode_text: text weights are not bound");
  PC_REQUIRE(tokens && out && P > 0, PC_ERR_ARG, "pc_encode_text: null buffer or empty batch");
  const Tower& t = ctx->txt;
  const int mb = pick_micro_batch(t.L, P, micro_batch);
  PC_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PC_ERR_ALIGN,
             "pc_encode_text: workspace must be 256-byte aligned");
  PC_REQUIRE(workspace_bytes >= tower_ws_bytes(mb * t.L, t.width, mb), PC_ERR_WORKSPACE,
             "pc_encode_text: workspace %zu < %zu", workspace_bytes, tower_ws_bytes(mb * t.L, t.width, mb));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const TowerWs ws = carve(workspace, mb * t.L, t.width, mb);
  const int d = t.width;
  __half* o = static_cast<__half*>(out);
  for (int p0 = 0; p0 < P; p0 += mb) {
    const int n = (P - p0 < mb) ? (P - p0) : mb;
    const int64_t* tok = tokens + static_cast<size_t>(p0) * t.L;
    PC_TRY(launch_text_embed(tok, ctx->tok_emb, ctx->tpos, ws.x, n, t.L, d, ctx->vocab, s));
    if (fused_ln_enabled()) {
      PC_TRY(launch_row_stats(ws.x, ws.s1, n * t.L, d, s));
      for (int l = 0; l < t.layers; ++l)
        PC_TRY(resblock_fused(t, l, ws.x, ws.h, ws.big, ws.s1, l == 0 ? 1 : gemm_stats_parts(n * t.L, d), ws.s2, n, t.L, 1, s));
    } else {
      for (int l = 0; l < t.layers; ++l) PC_TRY(resblock(t, l, ws.x, ws.h, ws.big, n, t.L, 1, s));
    }
    // ln_final is row-wise, so LN(gather(x)) == gather(LN(x)) (clip/model.py:348-352)
    PC_TRY(launch_eot_index(tok, ws.idx, n, t.L, s));
    PC_TRY(launch_layernorm_gather(ws.x, ws.idx, ws.h, ctx->ln_final_w, ctx->ln_final_b, n, d, s));
    GemmArgs g{};
    g.M = n; g.N = t.embed; g.K = d;
    g.A = ws.h; g.lda = d;
    g.W = t.proj_t; g.ldw = d;
    g.C = o + static_cast<size_t>(p0) * t.embed; g.ldc = t.embed;
    PC_TRY(launch_gemm(g, EPI_BIAS, s));
    if (l2norm) {
      __half* f = o + static_cast<size_t>(p0) * t.embed;
      PC_TRY(launch_l2norm(f, f, n, t.embed, s));
    }
  }
  return PC_OK;
}

size_t pc_resblock_workspace_bytes(const pc_ctx* ctx, int tower, int B, int L) {
  if (!ctx) return 0;
  const Tower& t = tower == PC_TOWER_TEXT ? ctx->txt : ctx->vis;
  if (!t.bound) return 0;
  return tower_ws_bytes(B * L, t.width, 1);
}

int pc_resblock_forward(pc_ctx* ctx, int tower, int layer, void* x, int B, int L, int causal, void* workspace,
                        size_t workspace_bytes, void* stream) {
  PC_TRY(use_device(ctx));
  PC_REQUIRE(tower == PC_TOWER_VISUAL || tower == PC_TOWER_TEXT, PC_ERR_ARG, "pc_resblock_forward: tower %d", tower);
  const Tower& t = tower == PC_TOWER_TEXT ? ctx->txt : ctx->vis;
  PC_REQUIRE(t.bound, PC_ERR_STATE, "pc_resblock_forward: tower %d is not bound", tower);
  PC_REQUIRE(layer >= 0 && layer < t.layers, PC_ERR_ARG, "pc_resblock_forward: layer %d of %d", layer, t.layers);
  PC_REQUIRE(x && B > 0 && L > 0, PC_ERR_ARG, "pc_resblock_forward: null x or empty batch");
  PC_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PC_ERR_ALIGN,
             "pc_resblock_forward: workspace must be 256-byte aligned");
  PC_REQUIRE(workspace_bytes >= tower_ws_bytes(B * L, t.width, 1), PC_ERR_WORKSPACE,
             "pc_resblock_forward: workspace %zu < %zu", workspace_bytes, tower_ws_bytes(B * L, t.width, 1));
  const TowerWs ws = carve(workspace, B * L, t.width, 1);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (fused_ln_enabled()) {
    PC_TRY(launch_row_stats(static_cast<const __half*>(x), ws.s1, B * L, t.width, s));
    return resblock_fused(t, layer, static_cast<__half*>(x), ws.h, ws.big, ws.s1, 1, ws.s2, B, L, causal, s);
  }
  return resblock(t, layer, static_cast<__half*>(x), ws.h, ws.big, B, L, causal, s);
}

int pc_resblock_forward_parts(pc_ctx* ctx, int tower, int layer, void* x, int B, int L, int causal, int parts,
                              int chained, void* workspace, size_t workspace_bytes, void* stream) {
  PC_TRY(use_device(ctx));
  PC_REQUIRE(tower == PC_TOWER_VISUAL || tower == PC_TOWER_TEXT, PC_ERR_ARG, "pc_resblock_forward_parts: tower %d", tower);
  const Tower& t = tower == PC_TOWER_TEXT ? ctx->txt : ctx->vis;
  PC_REQUIRE(t.bound, PC_ERR_STATE, "pc_resblock_forward_parts: tower %d is not bound", tower);
  PC_REQUIRE(layer >= 0 && layer < t.layers, PC_ERR_ARG, "pc_resblock_forward_parts: layer %d of %d", layer, t.layers);
  PC_REQUIRE(x && B > 0 && L > 0 && parts >= 1 && parts <= 31, PC_ERR_ARG, "pc_resblock_forward_parts: bad arguments");
  PC_REQUIRE(fused_ln_enabled(), PC_ERR_STATE, "pc_resblock_forward_parts: needs the LayerNorm-folded path (PC_NO_FUSED_LN unset)");
  PC_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PC_ERR_ALIGN,
             "pc_resblock_forward_parts: workspace must be 256-byte aligned");
  PC_REQUIRE(workspace_bytes >= tower_ws_bytes(B * L, t.width, 1), PC_ERR_WORKSPACE,
             "pc_resblock_forward_parts: workspace %zu < %zu", workspace_bytes, tower_ws_bytes(B * L, t.width, 1));
  const TowerWs ws = carve(workspace, B * L, t.width, 1);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!chained) PC_TRY(launch_row_stats(static_cast<const __half*>(x), ws.s1, B * L, t.width, s));
  return resblock_fused(t, layer, static_cast<__half*>(x), ws.h, ws.big, ws.s1,
                        chained ? gemm_stats_parts(B * L, t.width) : 1, ws.s2, B, L, causal, s, parts);
}

int pc_linear_forward(const void* x, int ldx, const void* w, int ldw, const void* bias, const void* residual,
                      int ldr, void* out, int ldo, int M, int N, int K, int epilogue, void* stream) {
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = static_cast<const __half*>(x); g.lda = ldx;
  g.W = static_cast<const __half*>(w); g.ldw = ldw;
  g.C = out; g.ldc = ldo;
  g.bias = static_cast<const __half*>(bias);
  g.residual = static_cast<const __half*>(residual); g.ldr = ldr;
  return launch_gemm(g, epilogue, static_cast<cudaStream_t>(stream));
}

int pc_linear_shift_relu_forward(const void* x, int ldx, const void* w, int ldw, const float* shift,
                                 const void* residual, int ldr, void* out, int ldo, int M, int N, int K, int epilogue,
                                 int relu, void* stream) {
  PC_REQUIRE(epilogue == EPI_BIAS || epilogue == EPI_BIAS_RES, PC_ERR_ARG,
             "pc_linear_shift_relu_forward: epilogue %d (PC_EPI_BIAS or PC_EPI_BIAS_RESIDUAL)", epilogue);
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.A = static_cast<const __half*>(x); g.lda = ldx;
  g.W = static_cast<const __half*>(w); g.ldw = ldw;
  g.C = out; g.ldc = ldo;
  g.bias_f32 = shift;
  g.relu = relu ? 1 : 0;
  g.residual = static_cast<const __half*>(residual); g.ldr = ldr;
  return launch_gemm(g, epilogue, static_cast<cudaStream_t>(stream));
}

int pc_conv3x3_shift_relu_forward(const void* x, const void* w, int ldw, const float* shift, void* out, int n, int h,
                                  int width, int cin, int cout, int relu, void* stream) {
  PC_REQUIRE(x && w && out && n > 0 && h > 0 && width > 0 && cin > 0 && cout > 0, PC_ERR_ARG,
             "pc_conv3x3_shift_relu_forward: null buffer or empty problem");
  PC_REQUIRE(cin % 8 == 0 && cout % 8 == 0 && ldw % 8 == 0, PC_ERR_ALIGN,
             "pc_conv3x3_shift_relu_forward: cin / cout / ldw (%d / %d / %d) must be multiples of 8", cin, cout, ldw);
  GemmArgs g{};
  g.N = cout; g.K = cin;
  g.A = static_cast<const __half*>(x); g.lda = cin;
  g.W = static_cast<const __half*>(w); g.ldw = ldw;
  g.C = out; g.ldc = cout;
  g.bias_f32 = shift;
  g.relu = relu ? 1 : 0;
  g.conv_taps = 9;
  g.conv_h = h; g.conv_w = width; g.conv_n = n;
  return launch_gemm(g, EPI_BIAS, static_cast<cudaStream_t>(stream));
}

int pc_layernorm_forward(const void* x, void* y, const void* gamma, const void* beta, int rows, int d,
                         void* stream) {
  PC_REQUIRE(x && y && gamma && beta, PC_ERR_ARG, "pc_layernorm_forward: null buffer");
  return launch_layernorm(static_cast<const __half*>(x), static_cast<__half*>(y), static_cast<const float*>(gamma),
                          static_cast<const float*>(beta), rows, d, 1, static_cast<cudaStream_t>(stream));
}

int pc_attention_forward(const void* qkv, void* out, int B, int L, int heads, int causal, void* stream) {
  return launch_attention(static_cast<const __half*>(qkv), static_cast<__half*>(out), B, L, heads, causal,
                          static_cast<cudaStream_t>(stream));
}

int pc_attention_rows_forward(const void* qkv, void* out, int B, int L, int heads, int row0, int nrows, int causal,
                              void* stream) {
  return launch_attention_rows(static_cast<const __half*>(qkv), 0, static_cast<__half*>(out), B, L, heads, row0, nrows,
                               causal, static_cast<cudaStream_t>(stream));
}

int pc_l2_normalize(const void* x, void* y, int rows, int d, void* stream) {
  PC_REQUIRE(x && y, PC_ERR_ARG, "pc_l2_normalize: null buffer");
  return launch_l2norm(static_cast<const __half*>(x), static_cast<__half*>(y), rows, d,
                       static_cast<cudaStream_t>(stream));
}

size_t pc_adapter_fc_workspace_bytes(int Q, int D, int reduction) {
  if (Q <= 0 || D <= 0 || reduction <= 0) return 0;
  const int H = D / reduction;
  return align_up(static_cast<size_t>(Q) * align_up(H, 8) * 2, 256) * 2 + align_up(static_cast<size_t>(Q) * D * 2, 256);
}

int pc_adapter_fc_forward(const pc_adapter_fc_weights* w, const void* q, void* out, int Q, int D, void* workspace,
                          size_t workspace_bytes, void* stream) {
  PC_REQUIRE(w && q && out && Q > 0 && D > 0, PC_ERR_ARG, "pc_adapter_fc_forward: null buffer or empty batch");
  PC_REQUIRE(w->fc0_weight && w->fc1_weight && w->fc1_bias && w->fc2_weight && w->fc3_weight && w->fc3_bias,
             PC_ERR_ARG, "pc_adapter_fc_forward: null weight");
  const int red = w->reduction > 0 ? w->reduction : 4;
  PC_REQUIRE(D % red == 0 && (D / red) % 8 == 0 && D % 8 == 0, PC_ERR_ARG,
             "pc_adapter_fc_forward: D = %d with reduction %d is not supported", D, red);
  const int H = D / red;
  PC_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PC_ERR_ALIGN,
             "pc_adapter_fc_forward: workspace must be 256-byte aligned");
  PC_REQUIRE(workspace_bytes >= pc_adapter_fc_workspace_bytes(Q, D, red), PC_ERR_WORKSPACE,
             "pc_adapter_fc_forward: workspace %zu < %zu", workspace_bytes, pc_adapter_fc_workspace_bytes(Q, D, red));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* p = static_cast<uint8_t*>(workspace);
  const size_t a = align_up(static_cast<size_t>(Q) * H * 2, 256);
  __half* h1 = reinterpret_cast<__half*>(p);
  __half* h1n = reinterpret_cast<__half*>(p + a);
  __half* h2 = reinterpret_cast<__half*>(p + 2 * a);
  GemmArgs g{};
  g.M = Q; g.N = H; g.K = D;
  g.A = static_cast<const __half*>(q); g.lda = D;
  g.W = static_cast<const __half*>(w->fc0_weight); g.ldw = D;
  g.C = h1; g.ldc = H;
  PC_TRY(launch_gemm(g, EPI_BIAS, s));
  PC_TRY(launch_ln_f16(h1, h1n, static_cast<const __half*>(w->fc1_weight), static_cast<const __half*>(w->fc1_bias),
                       Q, H, s));
  g = GemmArgs{};
  g.M = Q; g.N = D; g.K = H;
  g.A = h1n; g.lda = H;
  g.W = static_cast<const __half*>(w->fc2_weight); g.ldw = H;
  g.C = h2; g.ldc = D;
  PC_TRY(launch_gemm(g, EPI_BIAS, s));
  PC_TRY(launch_ln_blend_f16(h2, static_cast<const __half*>(q), static_cast<__half*>(out),
                             static_cast<const __half*>(w->fc3_weight), static_cast<const __half*>(w->fc3_bias), 0.2f,
                             Q, D, s));
  return PC_OK;
}

int pc_adapter_conv_forward(const pc_adapter_conv_weights* w, int c_type, const void* q, void* out, int Q, int D,
                            void* stream) {
  PC_REQUIRE(w && q && out, PC_ERR_ARG, "pc_adapter_conv_forward: null buffer");
  PC_REQUIRE(c_type == 2 || c_type == 3, PC_ERR_ARG, "pc_adapter_conv_forward: c_type %d (2 = conv-2x, 3 = conv-3x)",
             c_type);
  PC_REQUIRE(w->conv1_weight && w->conv3_weight && w->bn1_weight && w->bn1_bias && w->bn3_weight && w->bn3_bias,
             PC_ERR_ARG, "pc_adapter_conv_forward: null weight");
  PC_REQUIRE(c_type == 2 || (w->conv2_weight && w->bn2_weight && w->bn2_bias), PC_ERR_ARG,
             "pc_adapter_conv_forward: conv-3x needs conv2 / bn2");
  AdapterConvW k;
  k.conv1 = static_cast<const __half*>(w->conv1_weight);
  k.conv2 = static_cast<const __half*>(w->conv2_weight);
  k.conv3 = static_cast<const __half*>(w->conv3_weight);
  k.bn1_w = static_cast<const __half*>(w->bn1_weight);
  k.bn1_b = static_cast<const __half*>(w->bn1_bias);
  k.bn2_w = static_cast<const __half*>(w->bn2_weight);
  k.bn2_b = static_cast<const __half*>(w->bn2_bias);
  k.bn3_w = static_cast<const __half*>(w->bn3_weight);
  k.bn3_b = static_cast<const __half*>(w->bn3_bias);
  return launch_adapter_conv(k, c_type == 3, static_cast<const __half*>(q), static_cast<__half*>(out), Q, D,
                             static_cast<cudaStream_t>(stream));
}

int pc_build_prototypes(const void* V, int N, int K, int D, int per_shot_norm, void* z, float* znorm2,
                        void* stream) {
  return launch_build_prototypes(static_cast<const __half*>(V), N, K, D, per_shot_norm, static_cast<__half*>(z),
                                 znorm2, static_cast<cudaStream_t>(stream));
}

size_t pc_proto_classify_workspace_bytes(int Q, int N) {
  if (Q <= 0 || N <= 0) return 0;
  const int rows = Q < kClassifyChunk ? Q : kClassifyChunk;
  return static_cast<size_t>(rows) * 2 * align_up(N, 4) * sizeof(float);
}

namespace {
// dots[q, 0:N] = q . z_img^T, dots[q, Np:Np+N] = q . z_txt^T (fp32). When the two banks are adjacent in memory (the packed
// head state, pipeline.HeadState.pack) and N needs no padding they are ONE [2N, D] matrix: one GEMM launch with N' = 2N
//  instead of one per bank.
int bank_dots(const __half* q, const __half* z_img, const __half* z_txt, float* dots, int n, int N, int Np, int D,
              cudaStream_t s) {
  GemmArgs g{};
  g.M = n; g.K = D;
  g.A = q; g.lda = D;
  g.ldw = D;
  g.ldc = 2 * Np;
  if (z_txt == z_img + static_cast<size_t>(N) * D && Np == N) {
    g.N = 2 * N;
    g.W = z_img;
    g.C = dots;
    return launch_gemm(g, EPI_F32, s);
  }
  for (int bank = 0; bank < 2; ++bank) {
    g.N = N;
    g.W = bank == 0 ? z_img : z_txt;
    g.C = dots + bank * Np;
    PC_TRY(launch_gemm(g, EPI_F32, s));
  }
  return PC_OK;
}
}  // namespace

int pc_proto_classify(const void* q, const void* z_img, const void* z_txt, const float* zi_n2, const float* zt_n2,
                      int Q, int N, int D, float alpha, float beta, float* p_out, int64_t* argmax, float* pmax,
                      void* workspace, size_t workspace_bytes, void* stream) {
  PC_REQUIRE(q && z_img && z_txt && zi_n2 && zt_n2 && Q > 0 && N > 0 && D > 0, PC_ERR_ARG,
             "pc_proto_classify: null buffer or empty problem");
  PC_REQUIRE(D % 8 == 0, PC_ERR_ARG, "pc_proto_classify: D = %d must be a multiple of 8", D);
  PC_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PC_ERR_ALIGN,
             "pc_proto_classify: workspace must be 256-byte aligned");
  PC_REQUIRE(workspace_bytes >= pc_proto_classify_workspace_bytes(Q, N), PC_ERR_WORKSPACE,
             "pc_proto_classify: workspace %zu < %zu", workspace_bytes, pc_proto_classify_workspace_bytes(Q, N));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Np = static_cast<int>(align_up(N, 4));
  float* dots = static_cast<float*>(workspace);
  const __half* qh = static_cast<const __half*>(q);
  for (int q0 = 0; q0 < Q; q0 += kClassifyChunk) {
    const int n = (Q - q0 < kClassifyChunk) ? (Q - q0) : kClassifyChunk;
    PC_TRY(bank_dots(qh + static_cast<size_t>(q0) * D, static_cast<const __half*>(z_img), static_cast<const __half*>(z_txt),
                     dots, n, N, Np, D, s));
    PC_TRY(launch_proto_softmax(dots, 2 * Np, qh + static_cast<size_t>(q0) * D, D, zi_n2, zt_n2, n, N, alpha, beta,
                                p_out ? p_out + static_cast<size_t>(q0) * N : nullptr,
                                argmax ? argmax + q0 : nullptr, pmax ? pmax + q0 : nullptr, s));
  }
  return PC_OK;
}

int pc_proto_grid_search(const void* q, const void* z_img, const void* z_txt, const float* zi_n2, const float* zt_n2,
                         const int64_t* labels, int Q, int N, int D, const float* alphas, int n_alpha,
                         const float* betas, int n_beta, int* counts, void* workspace, size_t workspace_bytes,
                         void* stream) {
  PC_REQUIRE(q && z_img && z_txt && zi_n2 && zt_n2 && labels && alphas && betas && counts && Q > 0 && N > 0 && D > 0 &&
                 n_alpha > 0 && n_beta > 0,
             PC_ERR_ARG, "pc_proto_grid_search: null buffer or empty problem");
  PC_REQUIRE(D % 8 == 0, PC_ERR_ARG, "pc_proto_grid_search: D = %d must be a multiple of 8", D);
  PC_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PC_ERR_ALIGN,
             "pc_proto_grid_search: workspace must be 256-byte aligned");
  PC_REQUIRE(workspace_bytes >= pc_proto_classify_workspace_bytes(Q, N), PC_ERR_WORKSPACE,
             "pc_proto_grid_search: workspace %zu < %zu", workspace_bytes, pc_proto_classify_workspace_bytes(Q, N));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PC_CHECK_CUDA(cudaMemsetAsync(counts, 0, static_cast<size_t>(n_alpha) * n_beta * sizeof(int), s));
  const int Np = static_cast<int>(align_up(N, 4));
  float* dots = static_cast<float*>(workspace);
  const __half* qh = static_cast<const __half*>(q);
  for (int q0 = 0; q0 < Q; q0 += kClassifyChunk) {
    const int n = (Q - q0 < kClassifyChunk) ? (Q - q0) : kClassifyChunk;
    PC_TRY(bank_dots(qh + static_cast<size_t>(q0) * D, static_cast<const __half*>(z_img), static_cast<const __half*>(z_txt),
                     dots, n, N, Np, D, s));
    PC_TRY(launch_proto_grid(dots, 2 * Np, qh + static_cast<size_t>(q0) * D, D, zi_n2, zt_n2, n, N, labels + q0, alphas,
                             n_alpha, betas, n_beta, counts, s));
  }
  return PC_OK;
}

}  // extern "C"
#pragma GCC visibility pop
