/* libprotoclip_b200 — C ABI of the B200-native Proto-CLIP few-shot inference hot path.
 *
 * The reference (IRVLUTD/Proto-CLIP) has no FFI: its seams are Python call sites. Every entry point below
 * names the reference interface it stands in for (file:line relative to the reference tree); the Python
 * shells in proto-clip_b200/ (clip/, model.py, utils.py, main.py) bind these with ctypes and keep the
 * reference's names, argument meaning and error behaviour. INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - return 0 on success, a negative PC_ERR_* otherwise; pc_last_error() gives a thread-local message.
 *     Nothing throws or exits across this boundary.
 *   - every pointer is a DEVICE pointer unless noted; the caller owns all buffers including workspaces
 *     (16-byte aligned; cudaMalloc / torch allocations qualify). The library owns only the context
 *     (weight views plus two small re-laid-out weight copies made at bind time).
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); no hidden
 *     synchronisation or allocation on the hot path.
 *   - fp16 storage, fp32 accumulation / statistics, with the reference's fp16 rounding points
 *     (clip/model.py:373-394 convert_weights semantics).
 *   - a context is bound to one device and is not thread-safe: one context per rank.
 *   - sm_100 only: pc_ctx_create refuses any other device (no fallback path exists).
 */
#ifndef PROTOCLIP_B200_H_
#define PROTOCLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PC_VERSION 100

enum {
  PC_OK = 0,
  PC_ERR_ARG = -1,       /* bad shape / null pointer / unsupported size */
  PC_ERR_ALIGN = -2,     /* pointer or leading dimension not aligned for TMA / vector access */
  PC_ERR_ARCH = -3,      /* device is not sm_100 */
  PC_ERR_CUDA = -4,      /* CUDA runtime / driver error */
  PC_ERR_STATE = -5,     /* weights not bound, wrong device, ... */
  PC_ERR_WORKSPACE = -6  /* workspace too small */
};

enum { PC_TOWER_VISUAL = 0, PC_TOWER_TEXT = 1 };
enum { PC_IMG_F32 = 0, PC_IMG_F16 = 1 };
/* nn.Linear epilogues (pc_linear_forward) */
enum { PC_EPI_BIAS = 0, PC_EPI_BIAS_QUICKGELU = 1, PC_EPI_BIAS_RESIDUAL = 2, PC_EPI_F32 = 3 };

typedef struct pc_ctx pc_ctx;

int pc_version(void);
const char* pc_last_error(void);

/* One context per device / rank. Replaces the module-global CUDA state the reference gets from
 * `clip.load(...)` + `.cuda()` (clip/clip.py:92-139, main.py:495-496). */
int pc_ctx_create(int device, pc_ctx** out);
void pc_ctx_destroy(pc_ctx* ctx);
/* VisionTransformer.forward returns x[:, 0, :] of the last block (clip/model.py:232-236). By default pc_encode_image runs
 * that block's attention query, out_proj, ln_2 and MLP on the CLS rows only (K / V still come from every token): same
 * features, 6 % fewer flop for ViT-B/16. full = 1 computes every token of the last block like the reference does
 * (A/B timing; bench.py reports both), 0 the CLS rows only, -1 follows the environment (PC_FULL_LAST_BLOCK=1). */
int pc_ctx_set_full_last_block(pc_ctx* ctx, int full);

/* Parameters of one ResidualAttentionBlock, in the reference state-dict layout and dtype
 * (clip/model.py:169-181; fp16 Linear/MHA tensors, fp32 LayerNorm tensors after convert_weights). */
typedef struct {
  const void* ln_1_weight;      /* f32 [d]     */
  const void* ln_1_bias;        /* f32 [d]     */
  const void* in_proj_weight;   /* f16 [3d, d] rows = [Wq; Wk; Wv] */
  const void* in_proj_bias;     /* f16 [3d]    */
  const void* out_proj_weight;  /* f16 [d, d]  */
  const void* out_proj_bias;    /* f16 [d]     */
  const void* ln_2_weight;      /* f32 [d]     */
  const void* ln_2_bias;        /* f32 [d]     */
  const void* c_fc_weight;      /* f16 [4d, d] */
  const void* c_fc_bias;        /* f16 [4d]    */
  const void* c_proj_weight;    /* f16 [d, 4d] */
  const void* c_proj_bias;      /* f16 [d]     */
} pc_resblock_weights;

/* VisionTransformer (clip/model.py:204-238). heads = width / 64 (clip/model.py:273). */
typedef struct {
  int image_resolution, patch_size, width, layers, heads, embed_dim;
  const void* conv1_weight;          /* f16 [width, 3, p, p], no bias */
  const void* class_embedding;       /* f32 [width] */
  const void* positional_embedding;  /* f32 [(res/p)^2 + 1, width] */
  const void* ln_pre_weight;         /* f32 [width] */
  const void* ln_pre_bias;
  const void* ln_post_weight;        /* f32 [width] */
  const void* ln_post_bias;
  const void* proj;                  /* f16 [width, embed_dim] */
  const pc_resblock_weights* blocks; /* HOST array of `layers` entries (device pointers inside) */
} pc_vit_weights;

/* Text tower (clip/model.py:277-291, 341-354). heads = width / 64 (clip/model.py:418). */
typedef struct {
  int context_length, vocab_size, width, layers, heads, embed_dim;
  const void* token_embedding;       /* f32 [vocab, width] */
  const void* positional_embedding;  /* f32 [context_length, width] */
  const void* ln_final_weight;       /* f32 [width] */
  const void* ln_final_bias;
  const void* text_projection;       /* f16 [width, embed_dim] */
  const pc_resblock_weights* blocks; /* HOST array */
} pc_text_weights;

/* ModifiedResNet (clip/model.py:95-152; RN50 / RN101 / RN50x4 / RN50x16 / RN50x64 visual towers, config C5).
 * One bias-free Conv2d + eval-mode BatchNorm2d pair in the reference's state-dict layout and dtypes after
 * convert_weights (conv fp16, BatchNorm fp32; clip/model.py:17-26, 109-114, 373-394). */
typedef struct {
  const void* conv_weight;      /* f16 [Cout, Cin, k, k], k = 1 or 3 */
  const void* bn_weight;        /* f32 [Cout] */
  const void* bn_bias;          /* f32 [Cout] */
  const void* bn_running_mean;  /* f32 [Cout] */
  const void* bn_running_var;   /* f32 [Cout] */
} pc_conv_bn_weights;

/* Bottleneck (clip/model.py:10-53): conv1 1x1 inplanes->planes, conv2 3x3, avgpool(stride), conv3 1x1 planes->4*planes,
 * downsample = avgpool(stride) + 1x1 conv + BN when stride > 1 or inplanes != 4*planes. */
typedef struct {
  int inplanes, planes, stride;
  pc_conv_bn_weights conv1, conv2, conv3;
  pc_conv_bn_weights downsample; /* `downsample.0.weight` / `downsample.1.*`; conv_weight == NULL: no such branch */
} pc_bottleneck_weights;

typedef struct {
  int image_resolution, width, output_dim, heads; /* heads = width * 32 / 64 (clip/model.py:260) */
  int layers[4];                                  /* bottlenecks in layer1..layer4 */
  pc_conv_bn_weights stem[3];                     /* conv1/bn1 (3x3 stride 2), conv2/bn2, conv3/bn3 */
  const pc_bottleneck_weights* blocks;            /* HOST array, layer1 blocks first (device pointers inside) */
  const void* attnpool_positional_embedding;      /* f32 [(res/32)^2 + 1, 32*width] */
  const void *q_proj_weight, *q_proj_bias;        /* f16 [E, E], [E], E = 32*width (AttentionPool2d, :56-92) */
  const void *k_proj_weight, *k_proj_bias;
  const void *v_proj_weight, *v_proj_bias;
  const void *c_proj_weight, *c_proj_bias;        /* f16 [output_dim, E], [output_dim] */
} pc_rn_weights;

/* Bind a ModifiedResNet as the context's visual tower (replaces a bound ViT and vice versa). BatchNorm is folded
 * into re-laid-out fp16 weight copies + fp32 shifts owned by the context; pc_encode_image then runs this tower. */
int pc_rn_bind_weights(pc_ctx* ctx, const pc_rn_weights* w);

/* Bind tower weights (views; the tensors must outlive the context). Stands in for build_model +
 * load_state_dict (clip/model.py:397-434). Synchronous; may allocate (bind time only). */
int pc_vit_bind_weights(pc_ctx* ctx, const pc_vit_weights* w);
int pc_text_bind_weights(pc_ctx* ctx, const pc_text_weights* w);

/* CLIP.encode_image (clip/model.py:338-339 -> 221-238 for a ViT, -> 137-152 for a ModifiedResNet)
 * [+ the `/= norm` of utils.py:352 when l2norm != 0].
 * images: [B, 3, res, res] f32 or f16 (NCHW, already normalised); feat_out: f16 [B, embed_dim].
 * The batch is walked in passes of `micro_batch` images; 0 = library default: equal passes of at most twelve row-block
 * waves of the GEMM for a ViT (ViT-B/16: <= 1152 images), of at most 256 images for a ModifiedResNet. The workspace must
 * hold pc_encode_image_workspace_bytes(ctx, micro_batch) (0: the largest default pass). Features do not depend on it. */
size_t pc_encode_image_workspace_bytes(const pc_ctx* ctx, int micro_batch);
int pc_encode_image(pc_ctx* ctx, const void* images, int img_dtype, int B, void* feat_out, int l2norm,
                    int micro_batch, void* workspace, size_t workspace_bytes, void* stream);

/* `_transform(n_px)` of clip/clip.py:77-84 for ONE image: Resize(n_px, BICUBIC) -> CenterCrop(n_px) -> ToTensor ->
 * Normalize(CLIP mean / std). rgb: uint8 [H, W, 3] in DEVICE memory (what `image.convert("RGB")` holds); out: [3, n_px,
 * n_px] f32 or f16 (out_dtype = PC_IMG_F32 / PC_IMG_F16), e.g. one slot of the batch handed to pc_encode_image. Byte-
 * exact with Pillow's 8-bit antialiased resampler and torchvision's size / crop arithmetic (the third-party code the
 * reference calls); fp32 results identical to ToTensor + Normalize. The only host work is building the filter tables
 * (double precision) and one small host-to-device copy of them on `stream`. */
size_t pc_preprocess_workspace_bytes(int H, int W, int n_px);
int pc_preprocess_image(const uint8_t* rgb, int H, int W, int n_px, void* out, int out_dtype, void* workspace,
                        size_t workspace_bytes, void* stream);
/* The same for B images of one size, rgb: uint8 [B, H, W, 3] -> out [B, 3, n_px, n_px], in two launches (what a
 * DataLoader's default collate of decoded, same-size frames gives; the filter tables are built once per size). */
size_t pc_preprocess_batch_workspace_bytes(int B, int H, int W, int n_px);
int pc_preprocess_batch(const uint8_t* rgb, int B, int H, int W, int n_px, void* out, int out_dtype, void* workspace,
                        size_t workspace_bytes, void* stream);
/* `get_random_train_tfm()` of datasets/imagenet.py:8-23 (the augmentation build_cache_model's loader applies to the
 * support images, utils.py:303-310) for ONE image and a GIVEN draw: RandomResizedCrop(n_px, BICUBIC) with the box
 * (top, left, crop_h, crop_w) -> RandomHorizontalFlip (flip != 0) -> ToTensor -> Normalize(CLIP mean / std). The
 * random draws stay on the host (torchvision takes them from torch's global generator; the Python shell
 * `GPUTrainTransform` consumes it identically, so a seed gives the same box / flip as the reference). Byte-exact with
 * torchvision F.resized_crop + F.hflip on a PIL image: the box is resampled as an image of its own (taps stop at its
 * edges), both axes scaled independently. rgb: uint8 [H, W, 3] in device memory; out: [3, n_px, n_px] f32 / f16. */
size_t pc_preprocess_train_workspace_bytes(int crop_h, int crop_w, int n_px);
int pc_preprocess_train_image(const uint8_t* rgb, int H, int W, int top, int left, int crop_h, int crop_w, int flip,
                              int n_px, void* out, int out_dtype, void* workspace, size_t workspace_bytes, void* stream);

/* CLIP.encode_text (clip/model.py:341-354). tokens: int64 [P, context_length] (clip.tokenize output,
 * clip/clip.py:194-230); out: f16 [P, embed_dim]. */
size_t pc_encode_text_workspace_bytes(const pc_ctx* ctx, int micro_batch);
int pc_encode_text(pc_ctx* ctx, const int64_t* tokens, int P, void* out, int l2norm, int micro_batch,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ResidualAttentionBlock.forward (clip/model.py:187-190) of layer `layer` of a bound tower, in place on
 * x: f16 token-major [B*L, d] (the reference's [L, B, d] permuted to batch-first; the Python shell does the
 * permute). workspace >= pc_resblock_workspace_bytes(ctx, tower, B, L). */
size_t pc_resblock_workspace_bytes(const pc_ctx* ctx, int tower, int B, int L);
int pc_resblock_forward(pc_ctx* ctx, int tower, int layer, void* x, int B, int L, int causal, void* workspace,
                        size_t workspace_bytes, void* stream);
/* Measurement entry (bench.py's roofline legs): the same block, restricted to the launches named by `parts` (a mask of
 * PC_PART_*), each with the epilogue the towers use -- LayerNorm-folded QKV / c_fc, residual + row statistics for
 * out_proj / c_proj (clip/model.py:187-190), the attention core (clip/model.py:183-185). chained = 1: the LayerNorm
 * statistics left in the workspace by the previous call's c_proj are used (as between the blocks of a tower) instead
 * of a fresh row-statistics pass. Results are only meaningful for parts = PC_PART_ALL. */
#define PC_PART_QKV 1
#define PC_PART_ATTN 2
#define PC_PART_OUT 4
#define PC_PART_FC 8
#define PC_PART_PROJ 16
#define PC_PART_GEMMS 29
#define PC_PART_ALL 31
int pc_resblock_forward_parts(pc_ctx* ctx, int tower, int layer, void* x, int B, int L, int causal, int parts,
                              int chained, void* workspace, size_t workspace_bytes, void* stream);

/* Building blocks, exposed because the reference's nn.Modules are (and the parity tests probe them):
 * nn.Linear / F.linear: out[M,N] = x[M,K] @ w[N,K]^T (+ bias) with the epilogues above. */
int pc_linear_forward(const void* x, int ldx, const void* w, int ldw, const void* bias, const void* residual,
                      int ldr, void* out, int ldo, int M, int N, int K, int epilogue, void* stream);
/* 1x1 Conv2d + eval BatchNorm2d [+ identity] [+ ReLU] of a Bottleneck (clip/model.py:43-52) on NHWC pixels, i.e.
 * pc_linear_forward with an fp32 per-channel shift (the folded BatchNorm bias, nullable) and an optional ReLU applied
 * to the fp16 result (after the fp16 residual add for PC_EPI_BIAS_RESIDUAL). epilogue: PC_EPI_BIAS or
 * PC_EPI_BIAS_RESIDUAL. */
int pc_linear_shift_relu_forward(const void* x, int ldx, const void* w, int ldw, const float* shift,
                                 const void* residual, int ldr, void* out, int ldo, int M, int N, int K, int epilogue,
                                 int relu, void* stream);
/* 3x3 / stride 1 / padding 1 Conv2d + eval BatchNorm2d + ReLU of a Bottleneck or of the stem (clip/model.py:21-23,
 * 109-114, 44 / 139-141) on plain NHWC tensors: x f16 [n, h, w, cin], w f16 [cout, 9 * cin (ldw)] with column
 * (ky * 3 + kx) * cin + c (the BatchNorm scale already folded in), shift f32 [cout] (nullable), out f16 [n, h, w, cout].
 * Implicit GEMM: a tile's 128 rows are a rectangular pixel patch, loaded per tap through a 4-D TMA box shifted by
 * (kx - 1, ky - 1); the zero padding is TMA's out-of-bounds fill. cin % 8 == 0, cout % 8 == 0; h x w must be cut by
 * one of the patch shapes (32x4, 16x8, 8x8 of 2 images, 4x4 of 8 images), else PC_ERR_ARG (the towers then fall back
 * to a zero-bordered copy). */
int pc_conv3x3_shift_relu_forward(const void* x, const void* w, int ldw, const float* shift, void* out, int n, int h,
                                  int width, int cin, int cout, int relu, void* stream);
/* clip.model.LayerNorm.forward (clip/model.py:155-161): f16 in/out, f32 gamma/beta, eps 1e-5. */
int pc_layernorm_forward(const void* x, void* y, const void* gamma, const void* beta, int rows, int d,
                         void* stream);
/* nn.MultiheadAttention core (clip/model.py:173,183-185): packed qkv f16 [B*L, 3d] -> f16 [B*L, d]. */
int pc_attention_forward(const void* qkv, void* out, int B, int L, int heads, int causal, void* stream);
/* The same attention for query rows [row0, row0 + nrows) of every sequence only (keys / values: all L tokens);
 * out is compact f16 [B*nrows, d]. This is what x[:, 0, :] after the last block needs (clip/model.py:232-236, row0 = 0,
 * nrows = 1) and what the attention pool's query token needs (clip/model.py:70-92). nrows < L <= 1024. */
int pc_attention_rows_forward(const void* qkv, void* out, int B, int L, int heads, int row0, int nrows, int causal,
                              void* stream);
/* x / x.norm(dim=-1, keepdim=True) on f16 rows (utils.py:267,319,352; main.py:400-409). */
int pc_l2_normalize(const void* x, void* y, int rows, int d, void* stream);

/* Adapter_FC.forward (model.py:81-95). state-dict tensors, all f16: fc.0.weight [D/4, D], fc.1.{weight,bias}
 * [D/4], fc.2.weight [D, D/4], fc.3.{weight,bias} [D]. workspace >= pc_adapter_fc_workspace_bytes(Q, D). */
typedef struct {
  const void* fc0_weight;
  const void* fc1_weight;
  const void* fc1_bias;
  const void* fc2_weight;
  const void* fc3_weight;
  const void* fc3_bias;
  int reduction; /* 4 */
} pc_adapter_fc_weights;
size_t pc_adapter_fc_workspace_bytes(int Q, int D, int reduction);
int pc_adapter_fc_forward(const pc_adapter_fc_weights* w, const void* q, void* out, int Q, int D,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Adapter.forward (model.py:49-78), c_type 2 = 'conv-2x', 3 = 'conv-3x'. f16 tensors: conv1.weight [16,1,1,1],
 * conv2.weight [16,16,3,3], conv3.weight [1,16,1,1], bn{1,2}.{weight,bias} [16,S,S], bn3.{weight,bias} [1,S,S],
 * S = ceil(sqrt(D)). */
typedef struct {
  const void *conv1_weight, *conv2_weight, *conv3_weight;
  const void *bn1_weight, *bn1_bias, *bn2_weight, *bn2_bias, *bn3_weight, *bn3_bias;
} pc_adapter_conv_weights;
int pc_adapter_conv_forward(const pc_adapter_conv_weights* w, int c_type, const void* q, void* out, int Q,
                            int D, void* stream);

/* Prototype construction (main.py:399-405; zero-shot variant main.py:173-178 with per_shot_norm = 0).
 * V: f16 [N*K, D] class-contiguous; z: f16 [N, D]; znorm2: f32 [N] = |z|^2 of the stored fp16 values. */
int pc_build_prototypes(const void* V, int N, int K, int D, int per_shot_norm, void* z, float* znorm2,
                        void* stream);

/* P() + argmax (utils.py:225-244, main.py:436-438): p = a*softmax(-b*|q-z_img|^2) + (1-a)*softmax(-b*|q-z_txt|^2).
 * q: f16 [Q, D]; z_img, z_txt: f16 [N, D]; zi_n2 / zt_n2: f32 [N] (pc_build_prototypes). Outputs (each
 * nullable): p_out f32 [Q, N], argmax int64 [Q], pmax f32 [Q]. */
size_t pc_proto_classify_workspace_bytes(int Q, int N);
int pc_proto_classify(const void* q, const void* z_img, const void* z_txt, const float* zi_n2,
                      const float* zt_n2, int Q, int N, int D, float alpha, float beta, float* p_out,
                      int64_t* argmax, float* pmax, void* workspace, size_t workspace_bytes, void* stream);

/* Fused (alpha, beta) grid search (main.py:187-199 and 419-430 call P() 319 x 3 times on identical matmuls):
 * the similarities are computed once, then counts[a * n_beta + b] = #{rows : argmax_n P(alpha_a, beta_b)[row, n]
 * == labels[row]} (lowest-index tie-break, like Tensor.max). alphas / betas: f32 DEVICE arrays; labels: int64 [Q];
 * counts: int32 [n_alpha * n_beta], cleared by the call. Workspace: pc_proto_classify_workspace_bytes(Q, N). */
int pc_proto_grid_search(const void* q, const void* z_img, const void* z_txt, const float* zi_n2,
                         const float* zt_n2, const int64_t* labels, int Q, int N, int D, const float* alphas,
                         int n_alpha, const float* betas, int n_beta, int* counts, void* workspace,
                         size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROTOCLIP_B200_H_ */
