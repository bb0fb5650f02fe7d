"""The few-shot query path as one object: encode_image -> /norm -> adapter -> /norm -> P -> argmax.

This is the reference's test-time sequence (utils.py:349-352 pre_load_features, main.py:399-409 prototypes +
adapter, utils.py:225-244 P, main.py:438 argmax) with the head state (prototypes + adapter weights) packed
into ONE flat buffer so that multi-GPU runs need exactly one NCCL broadcast (dist.py).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch

from . import _native as nat

FC_KEYS = ("fc.0.weight", "fc.1.weight", "fc.1.bias", "fc.2.weight", "fc.3.weight", "fc.3.bias")
CONV_KEYS = ("conv1.weight", "conv2.weight", "conv3.weight", "bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias",
             "bn3.weight", "bn3.bias")


class HeadState:
    """z_img [N,D] f16, z_txt [N,D] f16, their squared norms [N] f32, adapter parameters (f16)."""

    def __init__(self, z_img, z_txt, zi_n2, zt_n2, adapter_kind: str, adapter: Dict[str, torch.Tensor],
                 alpha: float, beta: float):
        self.z_img, self.z_txt, self.zi_n2, self.zt_n2 = z_img, z_txt, zi_n2, zt_n2
        self.adapter_kind, self.adapter = adapter_kind, adapter
        self.alpha, self.beta = float(alpha), float(beta)

    # ---- flat packing: [z_img | z_txt | adapter tensors...] as f16, then zn2s as f32 viewed as f16 pairs
    def layout(self) -> "OrderedDict[str, Tuple[int, ...]]":
        lay: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
        lay["z_img"] = tuple(self.z_img.shape)
        lay["z_txt"] = tuple(self.z_txt.shape)
        for k in (FC_KEYS if self.adapter_kind == "fc" else CONV_KEYS):
            lay[k] = tuple(self.adapter[k].shape)
        return lay

    def pack(self) -> torch.Tensor:
        parts = [self.z_img.reshape(-1), self.z_txt.reshape(-1)]
        parts += [self.adapter[k].reshape(-1) for k in (FC_KEYS if self.adapter_kind == "fc" else CONV_KEYS)]
        parts += [self.zi_n2.view(torch.float16), self.zt_n2.view(torch.float16)]
        return torch.cat(parts).contiguous()

    @staticmethod
    def packed_numel(N: int, D: int, adapter_kind: str) -> int:
        import math
        n = 2 * N * D + 2 * N * 2
        if adapter_kind == "fc":
            H = D // 4
            n += H * D + 2 * H + D * H + 2 * D
        else:
            S = int(math.ceil(math.sqrt(D)))
            n += 16 + 16 * 16 * 9 + 16 + 4 * 16 * S * S + 2 * S * S
        return n

    @staticmethod
    def unpack(flat: torch.Tensor, N: int, D: int, adapter_kind: str, alpha: float, beta: float) -> "HeadState":
        import math
        off = 0

        def take(shape):
            nonlocal off
            n = 1
            for s in shape:
                n *= s
            t = flat[off:off + n].view(*shape)
            off += n
            return t

        z_img, z_txt = take((N, D)), take((N, D))
        adapter = {}
        if adapter_kind == "fc":
            H = D // 4
            shapes = {"fc.0.weight": (H, D), "fc.1.weight": (H,), "fc.1.bias": (H,), "fc.2.weight": (D, H),
                      "fc.3.weight": (D,), "fc.3.bias": (D,)}
            for k in FC_KEYS:
                adapter[k] = take(shapes[k])
        else:
            S = int(math.ceil(math.sqrt(D)))
            shapes = {"conv1.weight": (16, 1, 1, 1), "conv2.weight": (16, 16, 3, 3), "conv3.weight": (1, 16, 1, 1),
                      "bn1.weight": (16, S, S), "bn1.bias": (16, S, S), "bn2.weight": (16, S, S),
                      "bn2.bias": (16, S, S), "bn3.weight": (1, S, S), "bn3.bias": (1, S, S)}
            for k in CONV_KEYS:
                adapter[k] = take(shapes[k])
        zi_n2 = take((2 * N,)).view(torch.float32)
        zt_n2 = take((2 * N,)).view(torch.float32)
        return HeadState(z_img, z_txt, zi_n2, zt_n2, adapter_kind, adapter, alpha, beta)


def build_head_state(V: torch.Tensor, T: torch.Tensor, N: int, K: int, adapter_kind: str,
                     adapter: Dict[str, torch.Tensor], alpha: float, beta: float) -> HeadState:
    """V: visual memory [N*K, D] f16 (class-contiguous), T: textual memory [N, D] f16 (main.py:399-405)."""
    z_img, zi_n2 = nat.build_prototypes(V, N, K, per_shot_norm=True)
    z_txt, zt_n2 = nat.build_prototypes(T, N, 1, per_shot_norm=False)
    dev = V.device
    adapter = {k: v.to(device=dev, dtype=torch.float16).contiguous() for k, v in adapter.items()}
    if adapter_kind != "fc" and "conv2.weight" not in adapter:
        raise ValueError("conv adapter state needs conv2 / bn2 tensors (present even for conv-2x, model.py:39-41)")
    return HeadState(z_img, z_txt, zi_n2, zt_n2, adapter_kind, adapter, alpha, beta)


class FewShotClassifier:
    """Query images -> class predictions on one GPU."""

    def __init__(self, ctx: "nat.Context", head: HeadState, micro_batch: int = 0):
        self.ctx, self.head, self.micro_batch = ctx, head, micro_batch

    def adapt(self, feats: torch.Tensor) -> torch.Tensor:
        h = self.head
        if h.adapter_kind == "fc":
            q = nat.adapter_fc_forward(h.adapter, feats)
        else:
            q = nat.adapter_conv_forward(h.adapter, h.adapter_kind, feats)
        return nat.l2_normalize(q, out=q)

    def classify_features(self, feats: torch.Tensor, want_p: bool = False):
        """feats: L2-normalised image features f16 [Q, D] (what pre_load_features caches)."""
        h = self.head
        with nat.nvtx_range("protoclip.head"):
            q = self.adapt(feats)
            return nat.proto_classify(q, h.z_img, h.z_txt, h.zi_n2, h.zt_n2, h.alpha, h.beta, want_p=want_p)

    def classify(self, images: torch.Tensor, want_p: bool = False):
        """images: [B,3,R,R] f32/f16 on the context's device. Returns (p or None, argmax int64 [B], pmax)."""
        feats = self.ctx.encode_image(images, l2norm=True, micro_batch=self.micro_batch)
        return self.classify_features(feats, want_p=want_p)


def build_memory_sharded(ctx: "nat.Context", support_images, prompt_tokens: Optional[torch.Tensor],
                         micro_batch: int = 0, num_support: Optional[int] = None,
                         chunk: int = 1024) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Memory-bank construction sharded across ranks (utils.py:284-332 and 256-273 run it on one
    GPU): every rank encodes its contiguous slice of the label-sorted support set and of the prompt rows
    ([N*T, ctx] int64), then ONE all-gather per bank. `support_images` is either a tensor [N*K, 3, R, R] (the same on
    every rank, or at least the rank's own slice filled) or a callable `(lo, hi) -> images [hi - lo, 3, R, R]` that
    produces a slice on demand (bench.py draws its synthetic support set that way; `num_support` = N*K then).
    Returns (V [N*K, D] f16 L2-normalised, text features [N*T, D] f16 L2-normalised or None), identical on every rank
    and identical to the single-GPU result (per-row arithmetic does not depend on the shard or on `chunk`)."""
    from . import dist as pdist
    rank, _, world = pdist.env_rank()
    if not pdist.active():
        rank, world = 0, 1
    total = int(num_support) if callable(support_images) else support_images.shape[0]
    lo, hi = pdist.shard_bounds(total, rank, world)
    dev = ctx.device
    D = ctx.vis_desc["embed_dim"]
    feats = []
    for c0 in range(lo, hi, chunk):
        c1 = min(c0 + chunk, hi)
        imgs = support_images(c0, c1) if callable(support_images) else support_images[c0:c1]
        feats.append(ctx.encode_image(imgs.to(dev), l2norm=True, micro_batch=micro_batch))
    local = torch.cat(feats) if feats else torch.empty((0, D), dtype=torch.float16, device=dev)
    V = pdist.all_gather_rows(local, total)
    T = None
    if prompt_tokens is not None:
        tp = prompt_tokens.shape[0]
        lo, hi = pdist.shard_bounds(tp, rank, world)
        local_t = ctx.encode_text(prompt_tokens[lo:hi].to(dev), l2norm=True) if hi > lo else \
            torch.empty((0, D), dtype=torch.float16, device=dev)
        T = pdist.all_gather_rows(local_t, tp)
    return V, T
