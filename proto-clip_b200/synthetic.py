"""Deterministic synthetic weights and inputs (there is no network for checkpoints or datasets).

* ``make_state_dict`` builds a CLIP state dict with the OpenAI key names and the dtypes the reference ends
  up with after ``convert_weights`` (clip/model.py:373-394): fp16 for conv / Linear / MHA / projections, fp32
  for LayerNorm and embeddings. Values come from a seeded CPU generator, so the same (arch, seed) gives the
  same bytes on every box with this torch build; the reference (``build_model``) and this repo consume
  identical tensors. Scales follow clip/model.py:297-324 so activations have realistic magnitudes.
* ``class_structured_images`` follows SURVEY.md §8(d): per-class base pattern + 0.5 * noise, so that argmax
  over prototypes is meaningful even with random weights.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import torch

# name -> (embed_dim, image_resolution, vision_layers, vision_width, patch, ctx, vocab, text_width, text_heads, text_layers)
ARCHS: Dict[str, Tuple[int, ...]] = {
    "ViT-B/32": (512, 224, 12, 768, 32, 77, 49408, 512, 8, 12),
    "ViT-B/16": (512, 224, 12, 768, 16, 77, 49408, 512, 8, 12),
    "ViT-L/14": (768, 224, 24, 1024, 14, 77, 49408, 768, 12, 12),
    "ViT-L/14@336px": (768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12),
    # small towers for fast CPU oracles / golden fixtures (head_dim stays 64)
    "tiny": (64, 32, 2, 128, 8, 77, 512, 64, 1, 2),
    "small": (128, 64, 3, 256, 16, 77, 1024, 128, 2, 2),
    # ModifiedResNet towers (clip/model.py:95-152): vision_layers is the 4-tuple of bottleneck counts, vision_width
    # the stem width (embed_dim of the attention pool = 32 * width, heads = width / 2), no patch size
    "RN50": (1024, 224, (3, 4, 6, 3), 64, None, 77, 49408, 512, 8, 12),
    "RN50x16": (768, 384, (6, 8, 18, 8), 96, None, 77, 49408, 768, 12, 12),
    "rn_tiny": (64, 64, (1, 1, 1, 1), 16, None, 77, 512, 64, 1, 2),
    "rn_small": (128, 96, (2, 1, 2, 1), 32, None, 77, 1024, 128, 2, 2),
}


def arch_config(name: str) -> dict:
    e, r, vl, vw, p, ctx, vocab, tw, th, tl = ARCHS[name]
    return dict(embed_dim=e, image_resolution=r, vision_layers=vl, vision_width=vw, vision_patch_size=p,
                context_length=ctx, vocab_size=vocab, transformer_width=tw, transformer_heads=th,
                transformer_layers=tl)


def vit_flops_per_image(name: str) -> float:
    """2*MAC count of one encode_image (layers*(24 L d^2 + 4 L^2 d) + 2 g^2 3p^2 d + 2 d D)."""
    c = arch_config(name)
    g = c["image_resolution"] // c["vision_patch_size"]
    L, d, p = g * g + 1, c["vision_width"], c["vision_patch_size"]
    return c["vision_layers"] * (24.0 * L * d * d + 4.0 * L * L * d) + 2.0 * g * g * 3 * p * p * d + 2.0 * d * c["embed_dim"]


def vit_executed_flops_per_image(name: str, full_last_block: bool = False) -> float:
    """2*MAC count this library executes for one encode_image: the reference's count minus what the last block never
    computes when it runs on the CLS rows alone (out_proj, c_fc, c_proj and the attention queries of the other L - 1
    tokens: 18 (L - 1) d^2 + 4 L (L - 1) d; csrc/api.cu resblock_cls_only)."""
    full = vit_flops_per_image(name)
    if full_last_block:
        return full
    c = arch_config(name)
    g = c["image_resolution"] // c["vision_patch_size"]
    L, d = g * g + 1, c["vision_width"]
    return full - (18.0 * (L - 1) * d * d + 4.0 * L * (L - 1) * d)


def _blocks(sd, prefix: str, width: int, layers: int, gen: torch.Generator):
    attn_std = width ** -0.5
    proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
    fc_std = (2 * width) ** -0.5

    def n(*shape, std=1.0):
        return torch.randn(*shape, generator=gen) * std

    for i in range(layers):
        p = f"{prefix}{i}."
        sd[p + "attn.in_proj_weight"] = n(3 * width, width, std=attn_std).half()
        sd[p + "attn.in_proj_bias"] = n(3 * width, std=0.02).half()
        sd[p + "attn.out_proj.weight"] = n(width, width, std=proj_std).half()
        sd[p + "attn.out_proj.bias"] = n(width, std=0.02).half()
        sd[p + "ln_1.weight"] = 1.0 + n(width, std=0.05)
        sd[p + "ln_1.bias"] = n(width, std=0.05)
        sd[p + "mlp.c_fc.weight"] = n(4 * width, width, std=fc_std).half()
        sd[p + "mlp.c_fc.bias"] = n(4 * width, std=0.02).half()
        sd[p + "mlp.c_proj.weight"] = n(width, 4 * width, std=proj_std).half()
        sd[p + "mlp.c_proj.bias"] = n(width, std=0.02).half()
        sd[p + "ln_2.weight"] = 1.0 + n(width, std=0.05)
        sd[p + "ln_2.bias"] = n(width, std=0.05)


def rn_flops_per_image(name: str) -> float:
    """2*MAC count of one ModifiedResNet encode_image (convs + attention pool as the reference computes it)."""
    c = arch_config(name)
    w, r = c["vision_width"], c["image_resolution"]
    f = 0.0
    h = r // 2
    f += 2.0 * h * h * (27 * (w // 2) + 9 * (w // 2) * (w // 2) + 9 * (w // 2) * w)
    h //= 2
    inpl = w
    for li, nb in enumerate(c["vision_layers"]):
        planes = w * (2 ** li)
        for b in range(nb):
            stride = 2 if (li > 0 and b == 0) else 1
            f += 2.0 * h * h * (inpl * planes + 9 * planes * planes)
            ho = h // stride
            f += 2.0 * ho * ho * planes * planes * 4
            if stride > 1 or inpl != planes * 4:
                f += 2.0 * ho * ho * inpl * planes * 4
            h, inpl = ho, planes * 4
    L, E = h * h + 1, w * 32
    f += 2.0 * L * 3 * E * E + 4.0 * L * L * E + 2.0 * L * E * c["embed_dim"]
    return f


def _conv_bn(sd, prefix: str, idx: str, cout: int, cin: int, k: int, gen: torch.Generator, bn_name: str = None):
    """conv{idx}.weight fp16 + bn{idx}.{weight,bias,running_mean,running_var,num_batches_tracked} fp32 (eval BN)."""
    def n(*shape, std=1.0):
        return torch.randn(*shape, generator=gen) * std

    cname = f"{prefix}conv{idx}.weight" if bn_name is None else f"{prefix}0.weight"
    bname = f"{prefix}bn{idx}." if bn_name is None else f"{prefix}{bn_name}."
    sd[cname] = n(cout, cin, k, k, std=(2.0 / (cin * k * k)) ** 0.5).half()
    sd[bname + "weight"] = 1.0 + n(cout, std=0.1)
    sd[bname + "bias"] = n(cout, std=0.1)
    sd[bname + "running_mean"] = n(cout, std=0.1)
    sd[bname + "running_var"] = 1.0 + 0.2 * torch.rand(cout, generator=gen)
    sd[bname + "num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)


def _make_rn_visual(sd, c: dict, gen: torch.Generator):
    """ModifiedResNet parameters with the reference's key names (clip/model.py:10-152). BN statistics are random
    (not the identity) so that the eval-mode BatchNorm folding is exercised; bn3 gains are damped so that the
    residual stream keeps a stable magnitude over up to 40 bottlenecks."""
    def n(*shape, std=1.0):
        return torch.randn(*shape, generator=gen) * std

    w = c["vision_width"]
    _conv_bn(sd, "visual.", "1", w // 2, 3, 3, gen)
    _conv_bn(sd, "visual.", "2", w // 2, w // 2, 3, gen)
    _conv_bn(sd, "visual.", "3", w, w // 2, 3, gen)
    inpl = w
    for li, nb in enumerate(c["vision_layers"]):
        planes = w * (2 ** li)
        for b in range(nb):
            stride = 2 if (li > 0 and b == 0) else 1
            p = f"visual.layer{li + 1}.{b}."
            _conv_bn(sd, p, "1", planes, inpl, 1, gen)
            _conv_bn(sd, p, "2", planes, planes, 3, gen)
            _conv_bn(sd, p, "3", planes * 4, planes, 1, gen)
            sd[p + "bn3.weight"] = sd[p + "bn3.weight"] * 0.3
            if stride > 1 or inpl != planes * 4:
                _conv_bn(sd, p + "downsample.", "", planes * 4, inpl, 1, gen, bn_name="1")
            inpl = planes * 4
    E = w * 32
    g = c["image_resolution"] // 32
    std = E ** -0.5
    sd["visual.attnpool.positional_embedding"] = n(g * g + 1, E, std=std)
    for nm in ("k_proj", "q_proj", "v_proj"):
        sd[f"visual.attnpool.{nm}.weight"] = n(E, E, std=std).half()
        sd[f"visual.attnpool.{nm}.bias"] = n(E, std=0.02).half()
    sd["visual.attnpool.c_proj.weight"] = n(c["embed_dim"], E, std=std).half()
    sd["visual.attnpool.c_proj.bias"] = n(c["embed_dim"], std=0.02).half()


def make_state_dict(name: str, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    c = arch_config(name)
    gen = torch.Generator().manual_seed(1_000_003 * seed + 17)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    vw, p, e = c["vision_width"], c["vision_patch_size"], c["embed_dim"]

    def n(*shape, std=1.0):
        return torch.randn(*shape, generator=gen) * std

    if isinstance(c["vision_layers"], tuple):
        _make_rn_visual(sd, c, gen)
    else:
        g = c["image_resolution"] // p
        scale = vw ** -0.5
        sd["visual.class_embedding"] = n(vw, std=scale)
        sd["visual.positional_embedding"] = n(g * g + 1, vw, std=scale)
        sd["visual.proj"] = n(vw, e, std=scale).half()
        sd["visual.conv1.weight"] = n(vw, 3, p, p, std=(3 * p * p) ** -0.5).half()
        sd["visual.ln_pre.weight"] = 1.0 + n(vw, std=0.05)
        sd["visual.ln_pre.bias"] = n(vw, std=0.05)
        _blocks(sd, "visual.transformer.resblocks.", vw, c["vision_layers"], gen)
        sd["visual.ln_post.weight"] = 1.0 + n(vw, std=0.05)
        sd["visual.ln_post.bias"] = n(vw, std=0.05)
    tw = c["transformer_width"]
    sd["positional_embedding"] = n(c["context_length"], tw, std=0.01)
    sd["text_projection"] = n(tw, e, std=tw ** -0.5).half()
    sd["logit_scale"] = torch.tensor(2.6592)
    sd["token_embedding.weight"] = n(c["vocab_size"], tw, std=0.02)
    _blocks(sd, "transformer.resblocks.", tw, c["transformer_layers"], gen)
    sd["ln_final.weight"] = 1.0 + n(tw, std=0.05)
    sd["ln_final.bias"] = n(tw, std=0.05)
    return sd


def make_adapter_state_dict(kind: str, D: int, seed: int = 4, out_gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """fp16 adapter parameters with the reference's state-dict keys (model.py:19-47, 84-89).

    out_gain scales the last LayerNorm's affine (fc.3 / bn3). A trained adapter is a small correction on top of
    its residual branch; out_gain = trained_like_gain(D) makes the random adapter one too (unit-variance LN output
    would otherwise swamp the L2-normalised features, whose elements are ~D^-0.5), so synthetic queries stay
    classifiable. The golden fixtures use out_gain = 1 to exercise the arithmetic at full magnitude."""
    gen = torch.Generator().manual_seed(1_000_003 * seed + 29)

    def n(*shape, std=1.0):
        return torch.randn(*shape, generator=gen) * std

    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    if kind == "fc":
        H = D // 4
        sd["fc.0.weight"] = n(H, D, std=D ** -0.5).half()
        sd["fc.1.weight"] = (1.0 + n(H, std=0.05)).half()
        sd["fc.1.bias"] = n(H, std=0.05).half()
        sd["fc.2.weight"] = n(D, H, std=H ** -0.5).half()
        sd["fc.3.weight"] = (out_gain * (1.0 + n(D, std=0.05))).half()
        sd["fc.3.bias"] = (out_gain * n(D, std=0.05)).half()
    else:
        import math
        S = int(math.ceil(math.sqrt(D)))
        sd["conv1.weight"] = n(16, 1, 1, 1).half()
        sd["bn1.weight"] = (1.0 + n(16, S, S, std=0.05)).half()
        sd["bn1.bias"] = n(16, S, S, std=0.05).half()
        sd["conv2.weight"] = n(16, 16, 3, 3, std=(16 * 9) ** -0.5).half()
        sd["bn2.weight"] = (1.0 + n(16, S, S, std=0.05)).half()
        sd["bn2.bias"] = n(16, S, S, std=0.05).half()
        sd["conv3.weight"] = n(1, 16, 1, 1, std=0.25).half()
        sd["bn3.weight"] = (out_gain * (1.0 + n(1, S, S, std=0.05))).half()
        sd["bn3.bias"] = (out_gain * n(1, S, S, std=0.05)).half()
    return sd


def trained_like_gain(D: int) -> float:
    """Adapter output gain that keeps the random adapter a ~10 % perturbation of a unit-norm feature."""
    return 0.5 * D ** -0.5


def class_bases(num_classes: int, resolution: int, seed: int = 1, device="cpu") -> torch.Tensor:
    """Per-class base patterns c_n ~ N(0,1), fp32 [N, 3, R, R]."""
    gen = torch.Generator(device=device).manual_seed(1_000_003 * seed + 41)
    return torch.randn(num_classes, 3, resolution, resolution, generator=gen, device=device)


def class_structured_images(bases: torch.Tensor, labels: torch.Tensor, seed: int, noise: float = 0.5) -> torch.Tensor:
    """image_i = c_{label_i} + noise * eps_i (eps from a generator on the bases' device), fp32 [B,3,R,R]."""
    gen = torch.Generator(device=bases.device).manual_seed(1_000_003 * seed + 43)
    eps = torch.randn((labels.numel(),) + tuple(bases.shape[1:]), generator=gen, device=bases.device)
    return bases[labels.to(bases.device)] + noise * eps


def aligned_text_memory(V: torch.Tensor, num_classes: int, shots: int, seed: int = 6, rel_noise: float = 0.02) -> torch.Tensor:
    """Synthetic textual memory bank [N, D] (fp16) that is ALIGNED with the visual memory, as a trained
    Proto-CLIP-F textual memory is (main.py:352-369 saves it as a learned embedding): the per-class mean of the
    visual memory V [N*K, D] plus a small perturbation of norm rel_noise. A random-init text tower cannot
    produce class-aligned features, so its output is exercised by the tower tests, not by the classifier."""
    D = V.shape[-1]
    gen = torch.Generator(device=V.device).manual_seed(1_000_003 * seed + 47)
    mean = V.float().view(num_classes, shots, D).mean(dim=1)
    mean = mean / mean.norm(dim=-1, keepdim=True)
    noise = torch.randn(num_classes, D, generator=gen, device=V.device) * (rel_noise * D ** -0.5)
    return (mean + noise).half()
