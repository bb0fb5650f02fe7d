"""CPU oracle for the CLIP image preprocessing of the reference (clip/clip.py:77-84).  TEST INFRASTRUCTURE ONLY.

    Compose([Resize(n_px, interpolation=BICUBIC), CenterCrop(n_px), convert("RGB"), ToTensor(), Normalize(mean, std)])

The arithmetic lives in third-party code that is not under /root/reference: Pillow's 8-bit antialiased resampler
(`src/libImaging/Resample.c`; Pillow 12.2.0 here, unpinned by the reference's requirements.txt) and torchvision's
functional transforms (0.26). It is integer / byte work, restated below in plain numpy + Python loops, each function
naming the routine it follows. Pinned by tests/test_preprocess.py against the live PIL + torchvision pipeline on random
images of many sizes (bit-exact on the uint8 stage, bit-exact fp32 after ToTensor / Normalize).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c: fixed-point fraction of the 8 bpc coefficient tables
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)  # clip/clip.py:83
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def bicubic_filter(x: float) -> float:
    """Resample.c bicubic_filter (a = -0.5), support 2.0."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, in0: float, in1: float, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc: per output index the first source index, the tap count and
    `ksize` fixed-point weights (int32, PRECISION_BITS fractional bits)."""
    scale = (in1 - in0) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _clip8(acc: np.ndarray) -> np.ndarray:
    """Resample.c clip8: (acc >> PRECISION_BITS) clamped to a byte."""
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """PIL.Image.resize((new_w, new_h), BICUBIC) on an RGB uint8 image [H, W, 3]: ImagingResampleInner — horizontal
    pass over the source rows the vertical pass needs, uint8 intermediate, then the vertical pass; a pass whose size
    does not change is skipped."""
    H, W, _ = img.shape
    out = img
    if (new_w, new_h) == (W, H):
        return img.copy()
    bv, kv, _ = precompute_coeffs(H, 0.0, float(H), new_h)
    if new_w != W:
        bh, kh, _ = precompute_coeffs(W, 0.0, float(W), new_w)
        y0 = int(bv[0, 0])
        y1 = int(bv[new_h - 1, 0] + bv[new_h - 1, 1])
        src = img[y0:y1].astype(np.int64)
        tmp = np.empty((y1 - y0, new_w, 3), dtype=np.uint8)
        for xx in range(new_w):
            xmin, n = int(bh[xx, 0]), int(bh[xx, 1])
            acc = (src[:, xmin:xmin + n, :] * kh[xx, :n].astype(np.int64)[None, :, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
            tmp[:, xx, :] = _clip8(acc)
        out = tmp
        bv = bv.copy()
        bv[:, 0] -= y0
    if new_h != H:
        src = out.astype(np.int64)
        res = np.empty((new_h, out.shape[1], 3), dtype=np.uint8)
        for yy in range(new_h):
            ymin, n = int(bv[yy, 0]), int(bv[yy, 1])
            acc = (src[ymin:ymin + n] * kv[yy, :n].astype(np.int64)[:, None, None]).sum(axis=0) + (1 << (PRECISION_BITS - 1))
            res[yy] = _clip8(acc)
        out = res
    return out


def resized_size(h: int, w: int, n_px: int) -> Tuple[int, int]:
    """torchvision F.resize with a single int: the SHORTER side becomes n_px, the longer one int(n_px * long / short)
    (_compute_resized_output_size). Returns (new_h, new_w)."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = n_px, int(n_px * long_ / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def crop_offsets(h: int, w: int, n_px: int) -> Tuple[int, int]:
    """torchvision F.center_crop: int(round((size - crop) / 2.0)) with Python's round-half-to-even."""
    return int(round((h - n_px) / 2.0)), int(round((w - n_px) / 2.0))


def clip_preprocess_u8(img: np.ndarray, n_px: int) -> np.ndarray:
    """Resize + CenterCrop of clip/clip.py:79-80 on an RGB uint8 image [H, W, 3] (both sides >= 1) -> [n_px, n_px, 3]."""
    H, W, _ = img.shape
    new_h, new_w = resized_size(H, W, n_px)
    r = resize_bicubic_u8(img, new_w, new_h)
    top, left = crop_offsets(new_h, new_w, n_px)
    return np.ascontiguousarray(r[top:top + n_px, left:left + n_px])


def clip_preprocess(img: np.ndarray, n_px: int) -> np.ndarray:
    """The whole `_transform(n_px)` of clip/clip.py:77-84 -> float32 [3, n_px, n_px]: ToTensor (byte / 255 in fp32),
    Normalize ((x - mean) / std in fp32, in that order)."""
    u8 = clip_preprocess_u8(img, n_px)
    x = u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    mean = np.asarray(CLIP_MEAN, dtype=np.float32)[:, None, None]
    std = np.asarray(CLIP_STD, dtype=np.float32)[:, None, None]
    return ((x - mean) / std).astype(np.float32)


# ------------------------------------------------------------------------------------------------------------------
# Training-time augmentation of the support images (reference datasets/imagenet.py:8-23 `get_random_train_tfm`):
#
#     Compose([RandomResizedCrop(224, scale=(0.5, 1), interpolation=BICUBIC), RandomHorizontalFlip(0.5),
#              ToTensor(), Normalize(mean, std)])
#
# Third-party arithmetic again: torchvision 0.26 draws the crop box and the flip from torch's global generator
# (`RandomResizedCrop.get_params`, `RandomHorizontalFlip.forward`), crops the PIL image, and Pillow resamples the
# CROPPED image (so taps stop at the box edge) to size x size. Pinned by tests/test_preprocess.py against the live
# transform under the same seed.

TRAIN_SCALE = (0.5, 1.0)              # datasets/imagenet.py:16-17
TRAIN_RATIO = (3.0 / 4.0, 4.0 / 3.0)  # torchvision default


def random_resized_crop_params(height: int, width: int, scale=TRAIN_SCALE, ratio=TRAIN_RATIO):
    """RandomResizedCrop.get_params: ten tries of (area fraction, log-uniform aspect ratio) from torch's global
    generator, the first box that fits wins (two more draws place it); else the central fallback. -> (top, left, h, w).
    Consumes the generator exactly like torchvision, so the same seed gives the same box."""
    import torch
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,)).item()
            j = torch.randint(0, width - w + 1, size=(1,)).item()
            return i, j, h, w
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


def random_flip(p: float = 0.5) -> bool:
    """RandomHorizontalFlip.forward: one `torch.rand(1)` draw."""
    import torch
    return bool(torch.rand(1) < p)


def train_transform_u8(img: np.ndarray, top: int, left: int, h: int, w: int, flip: bool, size: int) -> np.ndarray:
    """F.resized_crop on a PIL image (crop, then Image.resize((size, size), BICUBIC) of the crop) + F.hflip."""
    crop = np.ascontiguousarray(img[top:top + h, left:left + w])
    r = resize_bicubic_u8(crop, size, size)
    return np.ascontiguousarray(r[:, ::-1]) if flip else r


def train_transform(img: np.ndarray, top: int, left: int, h: int, w: int, flip: bool, size: int = 224) -> np.ndarray:
    """`get_random_train_tfm()` for a given box / flip -> float32 [3, size, size]."""
    u8 = train_transform_u8(img, top, left, h, w, flip, size)
    x = u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    mean = np.asarray(CLIP_MEAN, dtype=np.float32)[:, None, None]
    std = np.asarray(CLIP_STD, dtype=np.float32)[:, None, None]
    return ((x - mean) / std).astype(np.float32)
