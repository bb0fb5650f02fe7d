"""Drop-in for the reference's clip/model.py on the ViT path: a ``CLIP`` nn.Module with the same attribute
surface (``encode_image``, ``encode_text``, ``dtype``, ``visual.input_resolution``, ``state_dict()`` keys) whose
forwards run on libprotoclip_b200 (sm_100a CUDA) instead of torch.nn. There is no eager fallback: calling an
encoder on a CPU copy raises.

build_model / convert_weights keep the reference's semantics (clip/model.py:373-434): architecture inferred from
tensor shapes, Linear / conv / MHA / projection tensors converted to fp16, LayerNorm and embeddings kept fp32.
ModifiedResNet checkpoints (no `visual.proj` key, clip/model.py:398) bind the RN tower of the library
(pc_rn_bind_weights: NHWC activations, eval BatchNorm folded into tcgen05 GEMM operands).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional

import torch
from torch import nn

from .. import _native as nat

_FP16_SUFFIXES = ("attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias",
                  "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias")
_FP16_KEYS = ("visual.conv1.weight", "visual.proj", "text_projection")
_RN_FP16_MARKERS = (".conv1.weight", ".conv2.weight", ".conv3.weight", ".downsample.0.weight", "_proj.weight", "_proj.bias")
_META_KEYS = ("input_resolution", "context_length", "vocab_size")


def _is_fp16_key(key: str) -> bool:
    if key in _FP16_KEYS or key.endswith(_FP16_SUFFIXES):
        return True
    # ModifiedResNet: every nn.Conv2d and the attention pool's nn.Linear modules (clip/model.py:377-380)
    return key.startswith("visual.") and key.endswith(_RN_FP16_MARKERS)


def convert_weights(state_dict: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """State-dict form of the reference's convert_weights (clip/model.py:373-394)."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in state_dict.items():
        if k in _META_KEYS:
            continue
        if k.endswith("num_batches_tracked"):
            out[k] = v.detach().clone()  # BatchNorm's int64 counter keeps its dtype
            continue
        out[k] = v.detach().half() if _is_fp16_key(k) else v.detach().float()
    return out


class _VisualInfo(nn.Module):
    """Carries the attributes callers read off ``model.visual`` (clip/clip.py:139, main.py)."""

    def __init__(self, input_resolution: int, output_dim: int, patch_size: int, width: int, layers: int):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.patch_size, self.width, self.layers = patch_size, width, layers


class CLIP(nn.Module):
    def __init__(self, state_dict: Dict[str, torch.Tensor]):
        super().__init__()
        sd = convert_weights(state_dict)
        self._keys = list(sd.keys())
        for k, v in sd.items():
            self.register_buffer(k.replace(".", "__"), v, persistent=False)
        if "visual.proj" in sd:  # clip/model.py:398-405
            width = sd["visual.conv1.weight"].shape[0]
            patch = sd["visual.conv1.weight"].shape[-1]
            grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
            layers = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
            self.visual = _VisualInfo(grid * patch, sd["visual.proj"].shape[1], patch, width, layers)
        else:  # ModifiedResNet, clip/model.py:406-414
            layers = tuple(len(set(k.split(".")[2] for k in sd if k.startswith(f"visual.layer{b}."))) for b in (1, 2, 3, 4))
            width = sd["visual.layer1.0.conv1.weight"].shape[0]
            grid = round((sd["visual.attnpool.positional_embedding"].shape[0] - 1) ** 0.5)
            assert grid ** 2 + 1 == sd["visual.attnpool.positional_embedding"].shape[0]
            self.visual = _VisualInfo(grid * 32, sd["visual.attnpool.c_proj.weight"].shape[0], None, width, layers)
        self.context_length = sd["positional_embedding"].shape[0]
        self.vocab_size = sd["token_embedding.weight"].shape[0]
        self._ctx: Optional["nat.Context"] = None
        self._ctx_device: Optional[torch.device] = None

    # ---- reference surface -------------------------------------------------------------------------
    @property
    def dtype(self) -> torch.dtype:
        return torch.float16  # clip/model.py:334-336 after convert_weights

    def state_dict(self, *args, **kwargs):  # reference key names
        return OrderedDict((k, getattr(self, k.replace(".", "__"))) for k in self._keys)

    def float(self):
        raise nat.NativeError("this CLIP runs fp16 storage / fp32 accumulation on sm_100a only; there is no fp32 "
                              "(CPU) execution path (reference: clip/clip.py:137-138)")

    def _context(self) -> "nat.Context":
        dev = getattr(self, self._keys[0].replace(".", "__")).device
        if dev.type != "cuda":
            raise nat.NativeError(f"CLIP weights are on {dev}; move the model to a CUDA (sm_100) device — "
                                  "there is no CPU fallback")
        if self._ctx is None or self._ctx_device != dev:
            self._ctx = nat.Context(dev)
            sd = self.state_dict()
            self._ctx.bind_visual(sd)
            self._ctx.bind_text(sd)
            self._ctx_device = self._ctx.device
        return self._ctx

    @torch.no_grad()
    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        """clip/model.py:338-339: [B,3,R,R] (any float dtype) -> fp16 [B, embed_dim], un-normalised."""
        ctx = self._context()
        return ctx.encode_image(image.to(ctx.device))

    @torch.no_grad()
    def encode_text(self, text: torch.Tensor) -> torch.Tensor:
        """clip/model.py:341-354: int tokens [P, context_length] -> fp16 [P, embed_dim]."""
        ctx = self._context()
        return ctx.encode_text(text.to(ctx.device))

    @torch.no_grad()
    def forward(self, image, text):
        """clip/model.py:356-371 (logit_scale.exp() * cosine similarities)."""
        fi = nat.l2_normalize(self.encode_image(image)).float()
        ft = nat.l2_normalize(self.encode_text(text)).float()
        scale = getattr(self, "logit_scale").float().exp() if "logit_scale" in self._keys else 1.0
        logits = scale * fi @ ft.t()
        return logits, logits.t()


def build_model(state_dict: Dict[str, torch.Tensor]) -> CLIP:
    """clip/model.py:397-434."""
    return CLIP(state_dict).eval()
