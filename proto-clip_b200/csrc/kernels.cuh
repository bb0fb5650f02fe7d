// Internal launcher declarations (host side). Every launcher is asynchronous on `stream`, allocates
// nothing, and returns a pc::PC_* code. Layouts: activations are token-major [rows, width] fp16 with the
// tokens of one image / prompt contiguous ([B, L, d]); weights keep the reference's [out, in] layout
// (nn.Linear.weight, clip/model.py:173-179) so a GEMM is always C[M,N] = A[M,K] * W[N,K]^T ("TN").
#pragma once
#include "common.cuh"

namespace pc {

// ---------------------------------------------------------------- gemm.cu (tcgen05 + TMA)
enum GemmEpilogue : int {
  EPI_BIAS = 0,        // C = f16(acc + bias)
  EPI_BIAS_QGELU = 1,  // C = quickgelu(f16(acc + bias))        (clip/model.py:164-166)
  EPI_BIAS_RES = 2,    // C = f16(acc + bias) + residual        (clip/model.py:188-189)
  EPI_F32 = 3,         // C(fp32) = acc (+ bias)
  // LayerNorm folded into the Linear that consumes it (clip/model.py:188-189: attn(ln_1(x)), mlp(ln_2(x))):
  //   LN(x) W^T + b = rstd_r * (x (W.gamma)^T)[r,n] - rstd_r * mean_r * s_n + c_n,
  //   s_n = sum_k (W.gamma)[n,k],  c_n = sum_k beta_k W[n,k] + b_n.
  // A is the un-normalised x, W the gamma-scaled weight; (sum x, sum x^2) per row come from `ln_stats`.
  EPI_LN_BIAS = 4,     // C = f16(LN-folded acc)
  EPI_LN_QGELU = 5     // C = quickgelu(f16(LN-folded acc))
};
struct GemmArgs {
  int M, N, K;
  const __half* A; int lda;   // [M, K] row-major, lda elements
  const __half* W; int ldw;   // [N, K] row-major
  void* C; int ldc;           // [M, N] fp16 (fp32 for EPI_F32)
  int c_planar;               // 1 (fp16 epilogues, N % 64 == 0): C is written as N/64 planes [N/64][M][64] -- column block j
                              // of the result is a contiguous [M, 64] matrix. The QKV projection uses it so that one
                              // head's q / k / v rows are contiguous 128-byte rows for the attention kernel's TMA loads.
  const __half* bias;         // [N] or nullptr
  const float* bias_f32;      // [N] fp32 bias used instead of `bias` when non-null (folded BatchNorm shift)
  // 3x3 / stride 1 / padding 1 convolution as an implicit GEMM over a ZERO-BORDERED NHWC activation
  // A = [B*(H+2)*(W+2), K = C] (conv_taps = 9, conv_pitch = W + 2): k-block (tap, c0) reads A rows shifted by
  // (tap/3 - 1) * conv_pitch + (tap%3 - 1) (TMA zero-fills rows outside the tensor) against W columns
  // [tap*C + c0, +64) of W [N, 9*C]; output rows are in the same bordered layout (border rows hold garbage).
  int conv_taps;              // 0 = plain GEMM, 9 = 3x3 taps
  int conv_pitch;             // W + 2
  // PATCH mode (conv_taps = 9, conv_w > 0; conv_patch_supported(h, w)): A = [n, h, w, K] and C = [n, h, w, N] are plain
  // (un-bordered) NHWC tensors. A tile's 128 rows are a rectangular patch bw x bh of bn images, loaded per tap with a
  // 4-D TMA box shifted by (dx, dy): the padding is TMA's out-of-bounds zero fill, no bordered copy, no frame rows in
  // the MMA work, and the result is stored through the same patch geometry. M is ignored (derived from the patches).
  int conv_h, conv_w, conv_n;
  int pw, ph, pn, px, py;     // filled by launch_gemm: patch box (bw, bh, bn) and patches per image row / column
  int relu;                   // 1: ReLU on the fp16 result (after the residual add for EPI_BIAS_RES): the conv + BN
                              // (+ identity) + ReLU of a Bottleneck (clip/model.py:43-52); EPI_BIAS / EPI_BIAS_RES only
  const __half* residual; int ldr;  // [M, N] fp16 (EPI_BIAS_RES); may alias C
  // LayerNorm folding (all fp32). Row statistics are (sum, sum of squares) pairs of the fp16 activations, kept as
  // `parts` partial pairs per row that the consumer adds in a fixed order (deterministic: no atomics).
  const float* ln_stats;      // [M][ln_parts][2] read by EPI_LN_*: statistics of A's rows (over K columns)
  int ln_parts;               // partial pairs per row in ln_stats
  const float* ln_s;          // [N] s_n (EPI_LN_*)
  const float* ln_c;          // [N] c_n (EPI_LN_*)
  float* stats_out;           // [M][gemm_stats_parts(M, N)][2] or nullptr: EPI_BIAS_RES writes the statistics of
                              // the row segments it produces (one partial pair per 64- or 128-column segment)
  int debug;                  // bring-up only (env PC_GEMM_DEBUG): 1 = skip epilogue body, 2 = skip TMA, 4 = skip MMA,
                              // 8 = per-tile cycle trace of CTA 0 (printed to stderr after a device sync)
  long long* trace;           // [tiles][16] clock64 samples when debug & 8
};
int launch_gemm(const GemmArgs& a, int epilogue, cudaStream_t stream);
// true when an h x w activation can be cut into 128-pixel patches (GemmArgs patch mode)
bool conv_patch_supported(int h, int w);
// partial (sum, sum^2) pairs per row that an EPI_BIAS_RES launch of this shape writes into stats_out
int gemm_stats_parts(int M, int N);

// ---------------------------------------------------------------- attention.cu (tcgen05 + TMA)
// qkv: [B*L, 3*d] fp16, columns [0,d) = q, [d,2d) = k, [2d,3d) = v, head h at h*64 (nn.MultiheadAttention
// packed in-proj order). out: [B*L, d] fp16 (heads merged). head_dim is fixed at 64 (all CLIP towers).
int launch_attention(const __half* qkv, __half* out, int B, int L, int heads, int causal, cudaStream_t stream);
// qkv_planar = 1: qkv is [3 * heads][B*L][64] (GemmArgs::c_planar); only attention6 shapes (attention6_supports(L))
int launch_attention_layout(const __half* qkv, int qkv_planar, __half* out, int B, int L, int heads, int causal,
                            cudaStream_t stream);
// query rows [row0, row0 + nrows) of every sequence only, all L keys (L <= 1024); out: compact [B * nrows, d]
int launch_attention_rows(const __half* qkv, int qkv_planar, __half* out, int B, int L, int heads, int row0, int nrows,
                          int causal, cudaStream_t stream);
// attention6.cu: the L <= 256 kernel (whole score row in TMEM, exact softmax, two query tiles in flight per SM);
// launch_attention dispatches to it. PC_ATTN_IMPL=5 selects the round-1 kernel (attention5.cu: 64-key blocks, online
// softmax, four tiles in flight), PC_ATTN_IMPL=2 keeps every shape on attention.cu's streaming kernel (A/B timing).
bool attention6_supports(int L);
bool attention6_supports_xkey(int L, int causal);  // L = 257: launch_attention6 + one tail query row
int launch_attention6(const __half* qkv, int qkv_planar, __half* out, int B, int L, int heads, int causal,
                      cudaStream_t stream);
// attention7.cu: unmasked L > 257 (ViT-L/14@336px: 577): 192-key blocks through the same softmax machinery, O resident
// in TMEM, online softmax with a lazy rescale, L % 192 == 1 through the extra-key trick
bool attention7_supports(int L, int causal);
int launch_attention7(const __half* qkv, __half* out, int B, int L, int heads, cudaStream_t stream);
bool attention5_supports(int L);
int launch_attention5(const __half* qkv, __half* out, int B, int L, int heads, int causal, cudaStream_t stream);

// ---------------------------------------------------------------- rowops.cu
// y[r,:] = f16(LN_fp32(x[src(r),:]) * gamma + beta); src(r) = r * row_stride_rows (gathers CLS rows when
// row_stride_rows = L), eps = 1e-5 (clip/model.py:155-161).
int launch_layernorm(const __half* x, __half* y, const float* gamma, const float* beta, int rows, int d,
                     int row_stride_rows, cudaStream_t stream);
// stats[r] = (sum x[r,:], sum x[r,:]^2), fp32: seeds the LayerNorm statistics of a block input (GemmArgs::ln_stats)
int launch_row_stats(const __half* x, float* stats, int rows, int d, cudaStream_t stream);
// bind-time folding of LayerNorm(gamma, beta) into the Linear (W [N,K], bias [N] or null) that consumes it:
// Wf = f16(W * gamma), s[n] = sum_k Wf[n,k], c[n] = sum_k beta[k] W[n,k] + bias[n]
int launch_fold_ln(const __half* W, const float* gamma, const float* beta, const __half* bias, __half* Wf, float* s,
                   float* c, int N, int K, cudaStream_t stream);
// images [B,3,R,R] (fp32 or fp16) -> patch rows [B*g*g, Kp] fp16, column = c*p*p + i*p + j (conv1 weight
// flattening order), zero-padded to Kp.
int launch_patchify(const void* images, int img_is_f16, __half* out, int B, int R, int p, int Kp,
                    cudaStream_t stream);
// x[b,0,:] = cls + pos[0]; x[b,1+t,:] = patch[b*g2+t,:] + pos[1+t]; then ln_pre (clip/model.py:222-227).
// stats (nullable): [B*L][2] fp32 (sum, sum of squares) of the stored rows = launch_row_stats(x) for free.
int launch_embed_ln_pre(const __half* patch, const float* cls, const float* pos, const float* gamma,
                        const float* beta, __half* x, float* stats, int B, int L, int d, cudaStream_t stream);
// text: x[p,t,:] = f16(tok_emb[tokens[p,t]]) + f16(pos[t]) (clip/model.py:342-344)
int launch_text_embed(const int64_t* tokens, const float* tok_emb, const float* pos, __half* x, int P, int L,
                      int d, int vocab, cudaStream_t stream);
// eot[p] = argmax_t tokens[p,t] (first max), as the row index p*L + eot into [P*L, d]
int launch_eot_index(const int64_t* tokens, int* rows, int P, int L, cudaStream_t stream);
// y[r,:] = LN(x[rows[r],:])
int launch_layernorm_gather(const __half* x, const int* rows, __half* y, const float* gamma, const float* beta,
                            int n, int d, cudaStream_t stream);
// in-place or out-of-place row L2 normalisation in fp16 storage / fp32 math: y = x / ||x||
int launch_l2norm(const __half* x, __half* y, int rows, int d, cudaStream_t stream);

// ---------------------------------------------------------------- convnet.cu (ModifiedResNet, clip/model.py:10-152)
// Activations are NHWC fp16 = pixel-major [B*H*W, C] GEMM operands.
// bind time: conv [Cout,Cin,k,k] fp16 + eval BatchNorm (eps 1e-5) -> wf [Cout,Kp] (column = tap*Cin + ci, scaled,
// zero-padded) and shift [Cout] fp32
int launch_fold_conv_bn(const __half* w, const float* gamma, const float* beta, const float* mean, const float* var,
                        __half* wf, float* shift, int Cout, int Cin, int k, int Kp, cudaStream_t stream);
// stem conv1 operand: images [B,3,R,R] -> [B*(R/2)^2, 32] (3x3, stride 2, pad 1; column = tap*3 + c, 27..31 zero)
int launch_stem_im2col(const void* images, int img_is_f16, __half* out, int B, int R, cudaStream_t stream);
// operand of the implicit 3x3 convolution (GemmArgs::conv_taps): x [B,H,W,C] -> xp [B,H+2,W+2,C], zero frame
int launch_pad_nhwc(const __half* x, __half* xp, int B, int H, int W, int C, cudaStream_t stream);
// interior of a bordered tensor: xp [B,H+2,W+2,C] -> x [B,H,W,C]
int launch_unpad_nhwc(const __half* xp, __half* x, int B, int H, int W, int C, cudaStream_t stream);
// zero the frame of a bordered tensor in place
int launch_zero_border(__half* xp, int B, int H, int W, int C, cudaStream_t stream);
// nn.AvgPool2d(s): [B,H,W,C] -> [B,H/s,W/s,C]; in_bordered: the input is [B,H+2,W+2,C], its interior is pooled
int launch_avgpool_nhwc(const __half* x, __half* y, int B, int H, int W, int C, int s, int in_bordered,
                        cudaStream_t stream);
// AttentionPool2d tokens: [B,HW,C] -> [B,HW+1,C] = [mean; pixels] + pos (clip/model.py:68-70)
int launch_attnpool_tokens(const __half* x, const float* pos, __half* tok, int B, int HW, int C, cudaStream_t stream);
// AttentionPool2d's single live query row (token 0, clip/model.py:72-92): q [B, E], kv [B*L, 2E] = k | v -> out [B, E]
int launch_attnpool_query0(const __half* q, const __half* kv, __half* out, int B, int L, int heads, cudaStream_t stream);

// ---------------------------------------------------------------- preprocess.cu (clip/clip.py:77-84 `_transform`)
// B same-size RGB uint8 images [B, H, W, 3] (device) -> [B, 3, n_px, n_px] fp32 / fp16: bicubic antialiased resize of the shorter side
// to n_px (byte-exact with Pillow), centre crop, /255, CLIP mean / std
size_t preprocess_workspace_bytes(int B, int H, int W, int n_px);
int launch_preprocess(const uint8_t* rgb, int B, int H, int W, int n_px, void* out, int out_f16, void* workspace,
                      size_t workspace_bytes, cudaStream_t stream);
// datasets/imagenet.py:8-23 `get_random_train_tfm` for a given box / flip: the ch x cw box at (top, left) of one image
// resampled (as a cropped image of its own) to n_px x n_px, mirrored when flip != 0, /255, CLIP mean / std
size_t preprocess_train_workspace_bytes(int ch, int cw, int n_px);
int launch_preprocess_train(const uint8_t* rgb, int H, int W, int top, int left, int ch, int cw, int flip, int n_px,
                            void* out, int out_f16, void* workspace, size_t workspace_bytes, cudaStream_t stream);

// ---------------------------------------------------------------- head.cu
int launch_build_prototypes(const __half* V, int N, int K, int D, int per_shot_norm, __half* z, float* zn2,
                            cudaStream_t stream);
int launch_text_prototypes(const __half* T, int N, int D, __half* z, float* zn2, cudaStream_t stream);
// Adapter_FC tail: y = LN(h) row-wise in fp16 semantics (model.py:84-88); blend: out = 0.2*LN(h2) + 0.8*x
int launch_ln_f16(const __half* x, __half* y, const __half* gamma, const __half* beta, int rows, int d,
                  cudaStream_t stream);
int launch_ln_blend_f16(const __half* h, const __half* x_in, __half* y, const __half* gamma,
                        const __half* beta, float ratio, int rows, int d, cudaStream_t stream);
struct AdapterConvW {
  const __half *conv1, *conv2, *conv3;            // [16], [16*16*9], [16]
  const __half *bn1_w, *bn1_b, *bn2_w, *bn2_b;    // [16,S,S]
  const __half *bn3_w, *bn3_b;                    // [S,S]
};
int launch_adapter_conv(const AdapterConvW& w, int three_x, const __half* q, __half* out, int Q, int D,
                        cudaStream_t stream);
// logits fp32 [Q, 2N] (= q . [z_img; z_txt]^T) -> p, argmax (utils.py:225-244)
int launch_proto_softmax(const float* dots, int ld, const __half* q, int D, const float* zi_n2,
                         const float* zt_n2, int Q, int N, float alpha, float beta, float* p_out,
                         int64_t* argmax, float* pmax, cudaStream_t stream);

// (alpha, beta) grid search over the same logits: counts[a * n_beta + b] += (argmax P(alpha_a, beta_b) == label)
// (main.py:187-199, 419-430). alphas / betas are device arrays; counts must be zero on entry.
int launch_proto_grid(const float* dots, int ld, const __half* q, int D, const float* zi_n2, const float* zt_n2, int Q,
                      int N, const int64_t* labels, const float* alphas, int n_alpha, const float* betas, int n_beta,
                      int* counts, cudaStream_t stream);

}  // namespace pc
