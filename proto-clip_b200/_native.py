"""ctypes binding of libprotoclip_b200.so (C ABI declared in include/protoclip_b200.h).

There is no CPU or PyTorch fallback behind this module: if the shared library is missing, or the device is
not sm_100, every entry point raises. torch is used only for device memory (tensors are passed as
``data_ptr()`` + shapes) and for the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprotoclip_b200.so")

PC_TOWER_VISUAL, PC_TOWER_TEXT = 0, 1
PC_IMG_F32, PC_IMG_F16 = 0, 1
EPI_BIAS, EPI_BIAS_QUICKGELU, EPI_BIAS_RESIDUAL, EPI_F32 = 0, 1, 2, 3

# every symbol include/protoclip_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "pc_version", "pc_last_error", "pc_ctx_create", "pc_ctx_destroy", "pc_ctx_set_full_last_block", "pc_vit_bind_weights",
    "pc_text_bind_weights", "pc_rn_bind_weights", "pc_linear_shift_relu_forward",
    "pc_conv3x3_shift_relu_forward",
    "pc_preprocess_workspace_bytes", "pc_preprocess_image", "pc_preprocess_batch_workspace_bytes", "pc_preprocess_batch",
    "pc_preprocess_train_workspace_bytes", "pc_preprocess_train_image", "pc_encode_image_workspace_bytes", "pc_encode_image",
    "pc_encode_text_workspace_bytes", "pc_encode_text", "pc_resblock_workspace_bytes", "pc_resblock_forward",
    "pc_resblock_forward_parts",
    "pc_linear_forward", "pc_layernorm_forward", "pc_attention_forward", "pc_attention_rows_forward",
    "pc_l2_normalize",
    "pc_adapter_fc_workspace_bytes", "pc_adapter_fc_forward", "pc_adapter_conv_forward", "pc_build_prototypes",
    "pc_proto_classify_workspace_bytes", "pc_proto_classify", "pc_proto_grid_search",
]


class nvtx_range:
    """NVTX range around a library call (towers, head): visible in Nsight Systems / `ncu --nvtx`; a no-op cost of two
    calls when no tool is attached (the reference has no profiler hooks)."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        try:
            torch.cuda.nvtx.range_push(self.name)
            self._on = True
        except Exception:  # nvtx unavailable (CPU-only import)
            self._on = False
        return self

    def __exit__(self, *exc):
        if self._on:
            torch.cuda.nvtx.range_pop()
        return False


class NativeError(RuntimeError):
    """A libprotoclip_b200 call returned a negative PC_ERR_* code."""


class ResblockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln_1_weight", "ln_1_bias", "in_proj_weight", "in_proj_bias", "out_proj_weight", "out_proj_bias",
        "ln_2_weight", "ln_2_bias", "c_fc_weight", "c_fc_bias", "c_proj_weight", "c_proj_bias")]


class VitWeights(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("image_resolution", "patch_size", "width", "layers", "heads", "embed_dim")] + \
               [(n, C.c_void_p) for n in ("conv1_weight", "class_embedding", "positional_embedding", "ln_pre_weight",
                                          "ln_pre_bias", "ln_post_weight", "ln_post_bias", "proj")] + \
               [("blocks", C.POINTER(ResblockWeights))]


class TextWeights(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("context_length", "vocab_size", "width", "layers", "heads", "embed_dim")] + \
               [(n, C.c_void_p) for n in ("token_embedding", "positional_embedding", "ln_final_weight",
                                          "ln_final_bias", "text_projection")] + \
               [("blocks", C.POINTER(ResblockWeights))]


class ConvBnWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("conv_weight", "bn_weight", "bn_bias", "bn_running_mean", "bn_running_var")]


class BottleneckWeights(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("inplanes", "planes", "stride")] + \
               [(n, ConvBnWeights) for n in ("conv1", "conv2", "conv3", "downsample")]


class RnWeights(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("image_resolution", "width", "output_dim", "heads")] + \
               [("layers", C.c_int * 4), ("stem", ConvBnWeights * 3), ("blocks", C.POINTER(BottleneckWeights))] + \
               [(n, C.c_void_p) for n in ("attnpool_positional_embedding", "q_proj_weight", "q_proj_bias",
                                          "k_proj_weight", "k_proj_bias", "v_proj_weight", "v_proj_bias",
                                          "c_proj_weight", "c_proj_bias")]


class AdapterFCWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("fc0_weight", "fc1_weight", "fc1_bias", "fc2_weight", "fc3_weight",
                                          "fc3_bias")] + [("reduction", C.c_int)]


class AdapterConvWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("conv1_weight", "conv2_weight", "conv3_weight", "bn1_weight", "bn1_bias",
                                          "bn2_weight", "bn2_bias", "bn3_weight", "bn3_bias")]


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """dlopen the in-tree library and declare signatures. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} is missing: build it with `make -C proto-clip_b200/csrc` (or __graft_entry__.build()). "
            "There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    lib.pc_version.restype = i
    lib.pc_last_error.restype = C.c_char_p
    lib.pc_ctx_create.argtypes = [i, C.POINTER(vp)]
    lib.pc_ctx_destroy.argtypes = [vp]
    lib.pc_ctx_destroy.restype = None
    lib.pc_ctx_set_full_last_block.argtypes = [vp, i]
    lib.pc_vit_bind_weights.argtypes = [vp, C.POINTER(VitWeights)]
    lib.pc_text_bind_weights.argtypes = [vp, C.POINTER(TextWeights)]
    lib.pc_rn_bind_weights.argtypes = [vp, C.POINTER(RnWeights)]
    lib.pc_encode_image_workspace_bytes.argtypes = [vp, i]
    lib.pc_encode_image_workspace_bytes.restype = sz
    lib.pc_encode_image.argtypes = [vp, vp, i, i, vp, i, i, vp, sz, vp]
    lib.pc_preprocess_workspace_bytes.argtypes = [i, i, i]
    lib.pc_preprocess_workspace_bytes.restype = sz
    lib.pc_preprocess_image.argtypes = [vp, i, i, i, vp, i, vp, sz, vp]
    lib.pc_preprocess_batch_workspace_bytes.argtypes = [i, i, i, i]
    lib.pc_preprocess_batch_workspace_bytes.restype = sz
    lib.pc_preprocess_batch.argtypes = [vp, i, i, i, i, vp, i, vp, sz, vp]
    lib.pc_preprocess_train_workspace_bytes.argtypes = [i, i, i]
    lib.pc_preprocess_train_workspace_bytes.restype = sz
    lib.pc_preprocess_train_image.argtypes = [vp, i, i, i, i, i, i, i, i, vp, i, vp, sz, vp]
    lib.pc_encode_text_workspace_bytes.argtypes = [vp, i]
    lib.pc_encode_text_workspace_bytes.restype = sz
    lib.pc_encode_text.argtypes = [vp, vp, i, vp, i, i, vp, sz, vp]
    lib.pc_resblock_workspace_bytes.argtypes = [vp, i, i, i]
    lib.pc_resblock_workspace_bytes.restype = sz
    lib.pc_resblock_forward.argtypes = [vp, i, i, vp, i, i, i, vp, sz, vp]
    lib.pc_resblock_forward_parts.argtypes = [vp, i, i, vp, i, i, i, i, i, vp, sz, vp]
    lib.pc_linear_forward.argtypes = [vp, i, vp, i, vp, vp, i, vp, i, i, i, i, i, vp]
    lib.pc_linear_shift_relu_forward.argtypes = [vp, i, vp, i, vp, vp, i, vp, i, i, i, i, i, i, vp]
    lib.pc_conv3x3_shift_relu_forward.argtypes = [vp, vp, i, vp, vp, i, i, i, i, i, i, vp]
    lib.pc_layernorm_forward.argtypes = [vp, vp, vp, vp, i, i, vp]
    lib.pc_attention_forward.argtypes = [vp, vp, i, i, i, i, vp]
    lib.pc_attention_rows_forward.argtypes = [vp, vp, i, i, i, i, i, i, vp]
    lib.pc_l2_normalize.argtypes = [vp, vp, i, i, vp]
    lib.pc_adapter_fc_workspace_bytes.argtypes = [i, i, i]
    lib.pc_adapter_fc_workspace_bytes.restype = sz
    lib.pc_adapter_fc_forward.argtypes = [C.POINTER(AdapterFCWeights), vp, vp, i, i, vp, sz, vp]
    lib.pc_adapter_conv_forward.argtypes = [C.POINTER(AdapterConvWeights), i, vp, vp, i, i, vp]
    lib.pc_build_prototypes.argtypes = [vp, i, i, i, i, vp, vp, vp]
    lib.pc_proto_classify_workspace_bytes.argtypes = [i, i]
    lib.pc_proto_classify_workspace_bytes.restype = sz
    lib.pc_proto_classify.argtypes = [vp, vp, vp, vp, vp, i, i, i, f, f, vp, vp, vp, vp, sz, vp]
    lib.pc_proto_grid_search.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, vp, i, vp, i, vp, vp, sz, vp]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().pc_last_error()
        raise NativeError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def stream_ptr(device: Optional[torch.device] = None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise NativeError(f"{name} must be a CUDA tensor (got {t.device}); this path has no CPU implementation")
    if t.dtype != dtype:
        raise NativeError(f"{name} must be {dtype} (got {t.dtype})")
    return t if t.is_contiguous() else t.contiguous()


class Workspace:
    """Caller-owned scratch (the C ABI never allocates on the hot path): a growable uint8 CUDA tensor."""

    def __init__(self, device: torch.device):
        self.device = device
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
        return self.buf


_workspaces: dict = {}


def workspace(device: torch.device, tag: str, nbytes: int) -> torch.Tensor:
    key = (str(device), tag)
    if key not in _workspaces:
        _workspaces[key] = Workspace(device)
    return _workspaces[key].get(nbytes)


# ----------------------------------------------------------------------------- primitive ops
def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, epilogue: int = EPI_BIAS,
           residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
           bias_f32: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """F.linear on the tcgen05 GEMM: x [M,K] f16, w [N,K] f16 -> [M,N] f16 (f32 for EPI_F32).
    `out`, if given, must be a contiguous [M, ldo >= N] tensor of the output dtype (ldo a multiple of 8, 4 for f32).
    bias_f32 / relu select the conv + folded-BatchNorm (+ identity) + ReLU form (pc_linear_shift_relu_forward)."""
    lib = load_library()
    x = require_cuda(x, torch.float16, "x")
    w = require_cuda(w, torch.float16, "w")
    M, K = x.shape
    N = w.shape[0]
    out_dtype = torch.float32 if epilogue == EPI_F32 else torch.float16
    pad = 4 if epilogue == EPI_F32 else 8
    ldo = (N + pad - 1) // pad * pad
    if out is None:
        out = torch.empty((M, ldo), dtype=out_dtype, device=x.device)
    else:
        out = require_cuda(out, out_dtype, "out")
        if out.shape[0] != M or out.shape[1] < N or out.stride(0) % pad:
            raise ValueError(f"out must be [M={M}, >= {N}] with a row pitch that is a multiple of {pad}")
        ldo = out.stride(0)
    if bias is not None:
        bias = require_cuda(bias, torch.float16, "bias")
    if residual is not None:
        residual = require_cuda(residual, torch.float16, "residual")
    if bias_f32 is not None or relu:
        if bias is not None:
            raise ValueError("pass either bias (f16) or bias_f32, not both")
        if bias_f32 is not None:
            bias_f32 = require_cuda(bias_f32, torch.float32, "bias_f32")
        with torch.cuda.device(x.device):
            check(lib.pc_linear_shift_relu_forward(
                x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), ptr(bias_f32), ptr(residual),
                residual.stride(0) if residual is not None else 0, out.data_ptr(), ldo, M, N, K, epilogue, int(relu),
                stream_ptr(x.device)), "pc_linear_shift_relu_forward")
        return out[:, :N] if out.shape[1] != N else out
    with torch.cuda.device(x.device):
        check(lib.pc_linear_forward(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), ptr(bias), ptr(residual),
                                    residual.stride(0) if residual is not None else 0, out.data_ptr(), ldo, M, N, K,
                                    epilogue, stream_ptr(x.device)), "pc_linear_forward")
    return out[:, :N] if out.shape[1] != N else out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    lib = load_library()
    x = require_cuda(x, torch.float16, "x")
    gamma = require_cuda(gamma, torch.float32, "gamma")
    beta = require_cuda(beta, torch.float32, "beta")
    y = torch.empty_like(x)
    rows, d = x.numel() // x.shape[-1], x.shape[-1]
    with torch.cuda.device(x.device):
        check(lib.pc_layernorm_forward(x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rows, d,
                                       stream_ptr(x.device)), "pc_layernorm_forward")
    return y


def attention(qkv: torch.Tensor, B: int, L: int, heads: int, causal: bool) -> torch.Tensor:
    lib = load_library()
    qkv = require_cuda(qkv, torch.float16, "qkv")
    d = heads * 64
    assert qkv.shape == (B * L, 3 * d), f"qkv must be [B*L, 3d] = {(B * L, 3 * d)}, got {tuple(qkv.shape)}"
    out = torch.empty((B * L, d), dtype=torch.float16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        check(lib.pc_attention_forward(qkv.data_ptr(), out.data_ptr(), B, L, heads, int(causal),
                                       stream_ptr(qkv.device)), "pc_attention_forward")
    return out


def conv3x3_shift_relu(x: torch.Tensor, weight: torch.Tensor, shift: Optional[torch.Tensor] = None,
                       relu: bool = True) -> torch.Tensor:
    """3x3 / pad 1 Conv2d (+ per-channel fp32 shift = folded eval BatchNorm) (+ ReLU) on an NHWC fp16 tensor
    (clip/model.py:21-23,44,109-114). weight: the module's [cout, cin, 3, 3] fp16 tensor (re-laid out here to the
    [cout, (ky, kx, c)] matrix the implicit GEMM reads). Returns NHWC fp16 [n, h, w, cout]."""
    lib = load_library()
    if not x.is_cuda:
        raise NativeError("conv3x3_shift_relu: CUDA tensors only; there is no CPU path")
    n, h, w, cin = x.shape
    cout = weight.shape[0]
    wm = weight.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous().half()
    x = x.contiguous().half()
    out = torch.empty((n, h, w, cout), dtype=torch.float16, device=x.device)
    sh = shift.float().contiguous() if shift is not None else None
    with torch.cuda.device(x.device):
        check(lib.pc_conv3x3_shift_relu_forward(x.data_ptr(), wm.data_ptr(), 9 * cin, sh.data_ptr() if sh is not None else None,
                                                out.data_ptr(), n, h, w, cin, cout, int(relu), stream_ptr(x.device)),
              "pc_conv3x3_shift_relu_forward")
    return out


def attention_rows(qkv: torch.Tensor, B: int, L: int, heads: int, row0: int, nrows: int, causal: bool = False) -> torch.Tensor:
    """Attention output of query rows [row0, row0 + nrows) of every sequence, compact [B*nrows, d]
    (pc_attention_rows_forward; clip/model.py:232-236 keeps row 0 of the last block)."""
    lib = load_library()
    d = heads * 64
    out = torch.empty((B * nrows, d), dtype=torch.float16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        check(lib.pc_attention_rows_forward(qkv.data_ptr(), out.data_ptr(), B, L, heads, row0, nrows, int(causal),
                                            stream_ptr(qkv.device)), "pc_attention_rows_forward")
    return out


def preprocess_image(rgb: torch.Tensor, n_px: int, out: Optional[torch.Tensor] = None,
                     dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """`_transform(n_px)` (clip/clip.py:77-84) on the GPU: one RGB uint8 image [H, W, 3] -> [3, n_px, n_px], or a batch of
    same-size images [B, H, W, 3] -> [B, 3, n_px, n_px] (two launches for the whole batch)."""
    lib = load_library()
    if not rgb.is_cuda:
        raise NativeError("preprocess_image: the image must be a CUDA uint8 tensor; there is no CPU path")
    if rgb.dtype != torch.uint8 or rgb.dim() not in (3, 4) or rgb.shape[-1] != 3:
        raise ValueError(f"preprocess_image: expected uint8 [H, W, 3] or [B, H, W, 3], got {rgb.dtype} {tuple(rgb.shape)}")
    rgb = rgb.contiguous()
    batched = rgb.dim() == 4
    B = int(rgb.shape[0]) if batched else 1
    H, W = int(rgb.shape[-3]), int(rgb.shape[-2])
    shape = (B, 3, n_px, n_px) if batched else (3, n_px, n_px)
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=rgb.device)
    elif tuple(out.shape) != shape or not out.is_contiguous() or out.dtype not in (torch.float32, torch.float16):
        raise ValueError(f"preprocess_image: out must be a contiguous {shape} f32 / f16 tensor")
    ws = workspace(rgb.device, "preprocess", lib.pc_preprocess_batch_workspace_bytes(B, H, W, n_px))
    with torch.cuda.device(rgb.device):
        check(lib.pc_preprocess_batch(rgb.data_ptr(), B, H, W, n_px, out.data_ptr(),
                                      PC_IMG_F16 if out.dtype == torch.float16 else PC_IMG_F32, ws.data_ptr(), ws.numel(),
                                      stream_ptr(rgb.device)), "pc_preprocess_batch")
    return out


def preprocess_train_image(rgb: torch.Tensor, box, flip: bool, n_px: int = 224, out: Optional[torch.Tensor] = None,
                           dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """`get_random_train_tfm()` (datasets/imagenet.py:8-23) on the GPU for a given draw: the `box` = (top, left, h, w)
    of one RGB uint8 image [H, W, 3] resampled to [3, n_px, n_px] (RandomResizedCrop), mirrored when `flip`
    (RandomHorizontalFlip), ToTensor, Normalize."""
    lib = load_library()
    if not rgb.is_cuda:
        raise NativeError("preprocess_train_image: the image must be a CUDA uint8 tensor; there is no CPU path")
    if rgb.dtype != torch.uint8 or rgb.dim() != 3 or rgb.shape[-1] != 3:
        raise ValueError(f"preprocess_train_image: expected uint8 [H, W, 3], got {rgb.dtype} {tuple(rgb.shape)}")
    rgb = rgb.contiguous()
    H, W = int(rgb.shape[0]), int(rgb.shape[1])
    top, left, ch, cw = (int(v) for v in box)
    shape = (3, n_px, n_px)
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=rgb.device)
    elif tuple(out.shape) != shape or not out.is_contiguous() or out.dtype not in (torch.float32, torch.float16):
        raise ValueError(f"preprocess_train_image: out must be a contiguous {shape} f32 / f16 tensor")
    ws = workspace(rgb.device, "preprocess", lib.pc_preprocess_train_workspace_bytes(ch, cw, n_px))
    with torch.cuda.device(rgb.device):
        check(lib.pc_preprocess_train_image(rgb.data_ptr(), H, W, top, left, ch, cw, int(bool(flip)), n_px, out.data_ptr(),
                                            PC_IMG_F16 if out.dtype == torch.float16 else PC_IMG_F32, ws.data_ptr(),
                                            ws.numel(), stream_ptr(rgb.device)), "pc_preprocess_train_image")
    return out


def l2_normalize(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = load_library()
    x = require_cuda(x, torch.float16, "x")
    y = torch.empty_like(x) if out is None else out
    rows, d = x.numel() // x.shape[-1], x.shape[-1]
    with torch.cuda.device(x.device):
        check(lib.pc_l2_normalize(x.data_ptr(), y.data_ptr(), rows, d, stream_ptr(x.device)), "pc_l2_normalize")
    return y


def build_prototypes(V: torch.Tensor, N: int, K: int, per_shot_norm: bool = True):
    """V f16 [N*K, D] -> (z f16 [N, D], |z|^2 f32 [N]) (main.py:399-403)."""
    lib = load_library()
    V = require_cuda(V, torch.float16, "V")
    D = V.shape[-1]
    assert V.numel() == N * K * D
    z = torch.empty((N, D), dtype=torch.float16, device=V.device)
    zn2 = torch.empty((N,), dtype=torch.float32, device=V.device)
    with torch.cuda.device(V.device):
        check(lib.pc_build_prototypes(V.data_ptr(), N, K, D, int(per_shot_norm), z.data_ptr(), zn2.data_ptr(),
                                      stream_ptr(V.device)), "pc_build_prototypes")
    return z, zn2


def proto_classify(q: torch.Tensor, z_img: torch.Tensor, z_txt: torch.Tensor, zi_n2: torch.Tensor,
                   zt_n2: torch.Tensor, alpha: float, beta: float, want_p: bool = True, want_argmax: bool = True):
    """P() + argmax (utils.py:225-244). Returns (p f32 [Q,N] or None, argmax int64 [Q] or None, pmax f32 [Q])."""
    lib = load_library()
    q = require_cuda(q, torch.float16, "q")
    z_img = require_cuda(z_img, torch.float16, "z_img")
    z_txt = require_cuda(z_txt, torch.float16, "z_txt")
    zi_n2 = require_cuda(zi_n2, torch.float32, "zi_n2")
    zt_n2 = require_cuda(zt_n2, torch.float32, "zt_n2")
    Q, D = q.shape
    N = z_img.shape[0]
    p = torch.empty((Q, N), dtype=torch.float32, device=q.device) if want_p else None
    am = torch.empty((Q,), dtype=torch.int64, device=q.device) if want_argmax else None
    pm = torch.empty((Q,), dtype=torch.float32, device=q.device)
    nbytes = lib.pc_proto_classify_workspace_bytes(Q, N)
    ws = workspace(q.device, "classify", nbytes)
    with torch.cuda.device(q.device):
        check(lib.pc_proto_classify(q.data_ptr(), z_img.data_ptr(), z_txt.data_ptr(), zi_n2.data_ptr(),
                                    zt_n2.data_ptr(), Q, N, D, float(alpha), float(beta), ptr(p), ptr(am),
                                    pm.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr(q.device)),
              "pc_proto_classify")
    return p, am, pm


def proto_grid_search(q: torch.Tensor, z_img: torch.Tensor, z_txt: torch.Tensor, zi_n2: torch.Tensor,
                      zt_n2: torch.Tensor, labels: torch.Tensor, alphas, betas) -> torch.Tensor:
    """Fused (alpha, beta) grid of P() + argmax + accuracy (main.py:187-199, 419-430). Returns the int32 matrix of
    correct predictions, [len(alphas), len(betas)]."""
    lib = load_library()
    q = require_cuda(q, torch.float16, "q")
    z_img = require_cuda(z_img, torch.float16, "z_img")
    z_txt = require_cuda(z_txt, torch.float16, "z_txt")
    zi_n2 = require_cuda(zi_n2, torch.float32, "zi_n2")
    zt_n2 = require_cuda(zt_n2, torch.float32, "zt_n2")
    labels = require_cuda(labels, torch.int64, "labels")
    Q, D = q.shape
    N = z_img.shape[0]
    if labels.numel() != Q:
        raise NativeError(f"proto_grid_search: {labels.numel()} labels for {Q} queries")
    a = torch.as_tensor(alphas, dtype=torch.float32).to(q.device).contiguous()
    b = torch.as_tensor(betas, dtype=torch.float32).to(q.device).contiguous()
    counts = torch.empty((a.numel(), b.numel()), dtype=torch.int32, device=q.device)
    nbytes = lib.pc_proto_classify_workspace_bytes(Q, N)
    ws = workspace(q.device, "classify", nbytes)
    with torch.cuda.device(q.device):
        check(lib.pc_proto_grid_search(q.data_ptr(), z_img.data_ptr(), z_txt.data_ptr(), zi_n2.data_ptr(),
                                       zt_n2.data_ptr(), labels.data_ptr(), Q, N, D, a.data_ptr(), a.numel(),
                                       b.data_ptr(), b.numel(), counts.data_ptr(), ws.data_ptr(), ws.numel(),
                                       stream_ptr(q.device)), "pc_proto_grid_search")
    return counts


# ----------------------------------------------------------------------------- context (one per device)
class Context:
    """Owns a pc_ctx and keeps the bound weight tensors alive (the library only stores views)."""

    def __init__(self, device: torch.device):
        self.lib = load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise NativeError(f"libprotoclip_b200 needs a CUDA (sm_100) device, got {self.device}")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        h = C.c_void_p()
        check(self.lib.pc_ctx_create(idx, C.byref(h)), "pc_ctx_create")
        self.handle = h
        self._keep: list = []
        self.vis_desc = None
        self.txt_desc = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.pc_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _blocks(self, sd: dict, prefix: str, layers: int):
        arr = (ResblockWeights * layers)()
        for i in range(layers):
            p = f"{prefix}{i}."
            names = {
                "ln_1_weight": ("ln_1.weight", torch.float32), "ln_1_bias": ("ln_1.bias", torch.float32),
                "in_proj_weight": ("attn.in_proj_weight", torch.float16),
                "in_proj_bias": ("attn.in_proj_bias", torch.float16),
                "out_proj_weight": ("attn.out_proj.weight", torch.float16),
                "out_proj_bias": ("attn.out_proj.bias", torch.float16),
                "ln_2_weight": ("ln_2.weight", torch.float32), "ln_2_bias": ("ln_2.bias", torch.float32),
                "c_fc_weight": ("mlp.c_fc.weight", torch.float16), "c_fc_bias": ("mlp.c_fc.bias", torch.float16),
                "c_proj_weight": ("mlp.c_proj.weight", torch.float16),
                "c_proj_bias": ("mlp.c_proj.bias", torch.float16),
            }
            for field, (key, dt) in names.items():
                setattr(arr[i], field, self._dev(sd[p + key], dt).data_ptr())
        return arr

    def _dev(self, t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self._keep.append(t)
        return t

    def _conv_bn(self, dst: ConvBnWeights, sd: dict, conv_key: str, bn_prefix: str) -> None:
        dst.conv_weight = self._dev(sd[conv_key], torch.float16).data_ptr()
        dst.bn_weight = self._dev(sd[bn_prefix + "weight"], torch.float32).data_ptr()
        dst.bn_bias = self._dev(sd[bn_prefix + "bias"], torch.float32).data_ptr()
        dst.bn_running_mean = self._dev(sd[bn_prefix + "running_mean"], torch.float32).data_ptr()
        dst.bn_running_var = self._dev(sd[bn_prefix + "running_var"], torch.float32).data_ptr()

    def bind_visual_rn(self, sd: dict) -> dict:
        """ModifiedResNet state dict (clip/model.py:95-152 key names; architecture inferred like build_model,
        clip/model.py:407-414). Conv / attention-pool Linear tensors -> f16, BatchNorm and pos-emb stay f32."""
        counts = [len(set(k.split(".")[2] for k in sd if k.startswith(f"visual.layer{b}."))) for b in (1, 2, 3, 4)]
        width = sd["visual.layer1.0.conv1.weight"].shape[0]
        grid = round((sd["visual.attnpool.positional_embedding"].shape[0] - 1) ** 0.5)
        out_dim = sd["visual.attnpool.c_proj.weight"].shape[0]
        w = RnWeights()
        w.image_resolution, w.width, w.output_dim, w.heads = grid * 32, width, out_dim, width * 32 // 64
        for i in range(4):
            w.layers[i] = counts[i]
        for i in range(3):
            self._conv_bn(w.stem[i], sd, f"visual.conv{i + 1}.weight", f"visual.bn{i + 1}.")
        blocks = (BottleneckWeights * sum(counts))()
        inpl, nb = width, 0
        for li in range(4):
            planes = width * (2 ** li)
            for b in range(counts[li]):
                p = f"visual.layer{li + 1}.{b}."
                blk = blocks[nb]
                blk.inplanes, blk.planes, blk.stride = inpl, planes, (2 if (li > 0 and b == 0) else 1)
                for j in (1, 2, 3):
                    self._conv_bn(getattr(blk, f"conv{j}"), sd, f"{p}conv{j}.weight", f"{p}bn{j}.")
                if p + "downsample.0.weight" in sd:
                    self._conv_bn(blk.downsample, sd, p + "downsample.0.weight", p + "downsample.1.")
                inpl, nb = planes * 4, nb + 1
        w.blocks = C.cast(blocks, C.POINTER(BottleneckWeights))
        pre = "visual.attnpool."
        w.attnpool_positional_embedding = self._dev(sd[pre + "positional_embedding"], torch.float32).data_ptr()
        for nm in ("q", "k", "v", "c"):
            setattr(w, f"{nm}_proj_weight", self._dev(sd[f"{pre}{nm}_proj.weight"], torch.float16).data_ptr())
            setattr(w, f"{nm}_proj_bias", self._dev(sd[f"{pre}{nm}_proj.bias"], torch.float16).data_ptr())
        with torch.cuda.device(self.device):
            check(self.lib.pc_rn_bind_weights(self.handle, C.byref(w)), "pc_rn_bind_weights")
        self.vis_desc = dict(image_resolution=grid * 32, patch_size=None, width=width, layers=tuple(counts),
                             heads=width * 32 // 64, embed_dim=out_dim, L=grid * grid + 1)
        return self.vis_desc

    def bind_visual(self, sd: dict) -> dict:
        """sd: OpenAI-CLIP state dict (clip/model.py:397-434 key names). fp16 conversion follows
        convert_weights (clip/model.py:373-394): Linear/conv/MHA/proj -> f16, LayerNorm/embeddings stay f32.
        State dicts without `visual.proj` hold a ModifiedResNet (clip/model.py:398) -> bind_visual_rn."""
        if "visual.proj" not in sd:
            return self.bind_visual_rn(sd)
        width = sd["visual.conv1.weight"].shape[0]
        patch = sd["visual.conv1.weight"].shape[-1]
        layers = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
        grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
        embed = sd["visual.proj"].shape[1]
        w = VitWeights()
        w.image_resolution, w.patch_size, w.width, w.layers = grid * patch, patch, width, layers
        w.heads, w.embed_dim = width // 64, embed
        w.conv1_weight = self._dev(sd["visual.conv1.weight"], torch.float16).data_ptr()
        w.class_embedding = self._dev(sd["visual.class_embedding"], torch.float32).data_ptr()
        w.positional_embedding = self._dev(sd["visual.positional_embedding"], torch.float32).data_ptr()
        w.ln_pre_weight = self._dev(sd["visual.ln_pre.weight"], torch.float32).data_ptr()
        w.ln_pre_bias = self._dev(sd["visual.ln_pre.bias"], torch.float32).data_ptr()
        w.ln_post_weight = self._dev(sd["visual.ln_post.weight"], torch.float32).data_ptr()
        w.ln_post_bias = self._dev(sd["visual.ln_post.bias"], torch.float32).data_ptr()
        w.proj = self._dev(sd["visual.proj"], torch.float16).data_ptr()
        blocks = self._blocks(sd, "visual.transformer.resblocks.", layers)
        w.blocks = C.cast(blocks, C.POINTER(ResblockWeights))
        with torch.cuda.device(self.device):
            check(self.lib.pc_vit_bind_weights(self.handle, C.byref(w)), "pc_vit_bind_weights")
        self.vis_desc = dict(image_resolution=grid * patch, patch_size=patch, width=width, layers=layers,
                             heads=width // 64, embed_dim=embed, L=grid * grid + 1)
        return self.vis_desc

    def bind_text(self, sd: dict) -> dict:
        width = sd["ln_final.weight"].shape[0]
        layers = len(set(k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")))
        w = TextWeights()
        w.context_length = sd["positional_embedding"].shape[0]
        w.vocab_size = sd["token_embedding.weight"].shape[0]
        w.width, w.layers, w.heads = width, layers, width // 64
        w.embed_dim = sd["text_projection"].shape[1]
        w.token_embedding = self._dev(sd["token_embedding.weight"], torch.float32).data_ptr()
        w.positional_embedding = self._dev(sd["positional_embedding"], torch.float32).data_ptr()
        w.ln_final_weight = self._dev(sd["ln_final.weight"], torch.float32).data_ptr()
        w.ln_final_bias = self._dev(sd["ln_final.bias"], torch.float32).data_ptr()
        w.text_projection = self._dev(sd["text_projection"], torch.float16).data_ptr()
        blocks = self._blocks(sd, "transformer.resblocks.", layers)
        w.blocks = C.cast(blocks, C.POINTER(ResblockWeights))
        with torch.cuda.device(self.device):
            check(self.lib.pc_text_bind_weights(self.handle, C.byref(w)), "pc_text_bind_weights")
        self.txt_desc = dict(context_length=w.context_length, vocab_size=w.vocab_size, width=width, layers=layers,
                             heads=width // 64, embed_dim=w.embed_dim)
        return self.txt_desc

    def encode_image(self, images: torch.Tensor, l2norm: bool = False, micro_batch: int = 0,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.vis_desc is None:
            raise NativeError("encode_image: visual weights are not bound")
        if not images.is_cuda:
            raise NativeError("encode_image: images must be a CUDA tensor; there is no CPU path")
        if images.dtype not in (torch.float32, torch.float16):
            images = images.float()
        images = images.contiguous()
        R = self.vis_desc["image_resolution"]
        if tuple(images.shape[1:]) != (3, R, R):
            raise NativeError(f"encode_image: expected [B, 3, {R}, {R}], got {tuple(images.shape)}")
        B = images.shape[0]
        if out is None:
            out = torch.empty((B, self.vis_desc["embed_dim"]), dtype=torch.float16, device=self.device)
        nbytes = self.lib.pc_encode_image_workspace_bytes(self.handle, micro_batch)
        ws = workspace(self.device, "tower", nbytes)
        with torch.cuda.device(self.device), nvtx_range("protoclip.encode_image"):
            check(self.lib.pc_encode_image(self.handle, images.data_ptr(),
                                           PC_IMG_F16 if images.dtype == torch.float16 else PC_IMG_F32, B,
                                           out.data_ptr(), int(l2norm), micro_batch, ws.data_ptr(), ws.numel(),
                                           stream_ptr(self.device)), "pc_encode_image")
        return out

    def encode_text(self, tokens: torch.Tensor, l2norm: bool = False, micro_batch: int = 0) -> torch.Tensor:
        if self.txt_desc is None:
            raise NativeError("encode_text: text weights are not bound")
        tokens = require_cuda(tokens.to(torch.int64), torch.int64, "tokens")
        P, L = tokens.shape
        if L != self.txt_desc["context_length"]:
            raise NativeError(f"encode_text: expected context length {self.txt_desc['context_length']}, got {L}")
        out = torch.empty((P, self.txt_desc["embed_dim"]), dtype=torch.float16, device=self.device)
        nbytes = self.lib.pc_encode_text_workspace_bytes(self.handle, micro_batch)
        ws = workspace(self.device, "tower", nbytes)
        with torch.cuda.device(self.device), nvtx_range("protoclip.encode_text"):
            check(self.lib.pc_encode_text(self.handle, tokens.data_ptr(), P, out.data_ptr(), int(l2norm), micro_batch,
                                          ws.data_ptr(), ws.numel(), stream_ptr(self.device)), "pc_encode_text")
        return out

    def set_full_last_block(self, full: Optional[bool]) -> None:
        """True: the last visual block computes every token like the reference (clip/model.py:232-233 then keeps the
        CLS row); False: CLS rows only (default); None: follow PC_FULL_LAST_BLOCK."""
        check(self.lib.pc_ctx_set_full_last_block(self.handle, -1 if full is None else int(bool(full))),
              "pc_ctx_set_full_last_block")

    def resblock_forward(self, tower: int, layer: int, x: torch.Tensor, B: int, L: int, causal: bool) -> torch.Tensor:
        """In place on x: f16 [B*L, d] token-major."""
        x = require_cuda(x, torch.float16, "x")
        nbytes = self.lib.pc_resblock_workspace_bytes(self.handle, tower, B, L)
        ws = workspace(self.device, "tower", nbytes)
        with torch.cuda.device(self.device):
            check(self.lib.pc_resblock_forward(self.handle, tower, layer, x.data_ptr(), B, L, int(causal),
                                               ws.data_ptr(), ws.numel(), stream_ptr(self.device)),
                  "pc_resblock_forward")
        return x

    def resblock_forward_parts(self, tower: int, layer: int, x: torch.Tensor, B: int, L: int, causal: bool,
                               parts: int, chained: bool) -> torch.Tensor:
        """Measurement entry (bench.py roofline legs): the block restricted to the launches in `parts` (mask: 1 QKV,
        2 attention, 4 out_proj, 8 c_fc, 16 c_proj; 29 = the four Linears with the towers' own epilogues, 31 = all)."""
        x = require_cuda(x, torch.float16, "x")
        nbytes = self.lib.pc_resblock_workspace_bytes(self.handle, tower, B, L)
        ws = workspace(self.device, "tower", nbytes)
        with torch.cuda.device(self.device):
            check(self.lib.pc_resblock_forward_parts(self.handle, tower, layer, x.data_ptr(), B, L, int(causal),
                                                     int(parts), int(chained), ws.data_ptr(), ws.numel(),
                                                     stream_ptr(self.device)), "pc_resblock_forward_parts")
        return x


# ----------------------------------------------------------------------------- adapters (functional form)
def adapter_fc_forward(params: dict, q: torch.Tensor, reduction: int = 4) -> torch.Tensor:
    """Adapter_FC.forward (model.py:91-95). params: state-dict tensors (f16, CUDA) keyed fc.0.weight ..."""
    lib = load_library()
    q = require_cuda(q, torch.float16, "q")
    Q, D = q.shape
    keep = [require_cuda(params[k], torch.float16, k) for k in
            ("fc.0.weight", "fc.1.weight", "fc.1.bias", "fc.2.weight", "fc.3.weight", "fc.3.bias")]
    w = AdapterFCWeights(*[t.data_ptr() for t in keep], reduction)
    out = torch.empty_like(q)
    nbytes = lib.pc_adapter_fc_workspace_bytes(Q, D, reduction)
    ws = workspace(q.device, "adapter", nbytes)
    with torch.cuda.device(q.device):
        check(lib.pc_adapter_fc_forward(C.byref(w), q.data_ptr(), out.data_ptr(), Q, D, ws.data_ptr(), ws.numel(),
                                        stream_ptr(q.device)), "pc_adapter_fc_forward")
    return out


def adapter_conv_forward(params: dict, c_type: str, q: torch.Tensor) -> torch.Tensor:
    """Adapter.forward (model.py:49-78); c_type 'conv-2x' | 'conv-3x'."""
    lib = load_library()
    q = require_cuda(q, torch.float16, "q")
    Q, D = q.shape
    names = ("conv1.weight", "conv2.weight", "conv3.weight", "bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias",
             "bn3.weight", "bn3.bias")
    keep = [require_cuda(params[k], torch.float16, k) for k in names]
    w = AdapterConvWeights(*[t.data_ptr() for t in keep])
    out = torch.empty_like(q)
    with torch.cuda.device(q.device):
        check(lib.pc_adapter_conv_forward(C.byref(w), 3 if c_type == "conv-3x" else 2, q.data_ptr(), out.data_ptr(),
                                          Q, D, stream_ptr(q.device)), "pc_adapter_conv_forward")
    return out
