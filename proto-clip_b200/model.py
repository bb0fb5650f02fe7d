"""Drop-in for the reference's model.py: ``Adapter`` (conv-2x / conv-3x) and ``Adapter_FC`` nn.Modules with the
same constructor arguments, parameter names (checkpoints in pretrained_ckpt/ load unchanged) and default
initialisation, whose forward runs on libprotoclip_b200 (model.py:12-95 of the reference). Inference only:
the kernels produce no autograd graph (training is outside the hot path).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

try:  # imported as proto_clip_b200.model (package) or as top-level `model` (main.py run as a script)
    from . import _native as nat
except ImportError:  # pragma: no cover
    from proto_clip_b200 import _native as nat


def _half_cuda_params(module: nn.Module) -> dict:
    out = {}
    for k, v in module.state_dict().items():
        if not v.is_cuda:
            raise nat.NativeError(f"{type(module).__name__}.{k} is on {v.device}: call .cuda() first — the adapter "
                                  "kernels are sm_100a-only and there is no CPU fallback")
        out[k] = v.detach().half().contiguous()
    return out


class Adapter(nn.Module):
    """Zero-pad D -> S*S (S = ceil(sqrt(D))), 1x1 conv 1->width, LN, [3x3 conv, LN if 'conv-3x'], 1x1 conv
    width->1, LN, + identity, crop to D. No activation is applied (the reference builds a ReLU it never calls)."""

    def __init__(self, c_in, c_type, width=16, dtype=None):
        super().__init__()
        if width != 16:
            raise ValueError("the sm_100a adapter kernel is specialised for width = 16 (the reference's value)")
        self.c_in, self.c_type = c_in, c_type
        size = int(math.ceil(math.sqrt(self.c_in)))
        self.conv1 = nn.Conv2d(1, width, kernel_size=1, stride=1, bias=False, dtype=dtype)
        self.bn1 = nn.LayerNorm([width, size, size], dtype=dtype)
        self.conv2 = nn.Conv2d(width, width, kernel_size=3, stride=1, padding=1, bias=False, dtype=dtype)
        self.bn2 = nn.LayerNorm([width, size, size], dtype=dtype)
        self.conv3 = nn.Conv2d(width, 1, kernel_size=1, stride=1, bias=False, dtype=dtype)
        self.bn3 = nn.LayerNorm([1, size, size], dtype=dtype)
        self.relu = nn.ReLU(inplace=True)  # kept for state/attribute parity; unused, as in the reference

    @torch.no_grad()
    def forward(self, x):
        x2 = x.reshape(-1, self.c_in)
        kind = "conv-3x" if self.c_type == "conv-3x" else "conv-2x"
        out = nat.adapter_conv_forward(_half_cuda_params(self), kind, x2.half())
        return out.to(x.dtype)


class Adapter_FC(nn.Module):
    """Linear(D -> D/r, no bias) -> LN -> Linear(D/r -> D, no bias) -> LN; out = 0.2 * that + 0.8 * input."""

    def __init__(self, c_in, reduction=4, dtype=None):
        super().__init__()
        self.reduction = reduction
        self.fc = nn.Sequential(
            nn.Linear(c_in, c_in // reduction, bias=False, dtype=dtype),
            nn.LayerNorm(c_in // reduction, dtype=dtype),
            nn.Linear(c_in // reduction, c_in, bias=False, dtype=dtype),
            nn.LayerNorm(c_in, dtype=dtype),
        )

    @torch.no_grad()
    def forward(self, image_features):
        c_in = self.fc[0].in_features
        x2 = image_features.reshape(-1, c_in)
        out = nat.adapter_fc_forward(_half_cuda_params(self), x2.half(), self.reduction)
        return out.to(image_features.dtype).reshape(image_features.shape)
