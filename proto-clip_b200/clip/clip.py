"""Drop-in for the reference's clip/clip.py: ``available_models``, ``load``, ``tokenize`` (clip/clip.py:87-230).

Differences, all forced by the environment or the scope:
  * no download: there is no network, so a model name resolves to ``<download_root or ~/.cache/clip>/<file>.pt``
    and raises RuntimeError when the file is absent (the reference would fetch it, clip/clip.py:41-70);
  * ``name`` may also be ``synthetic:<arch>[:seed]`` (random-init weights from proto_clip_b200.synthetic), or a
    path to a torch.save'd state dict / TorchScript archive exactly like the reference (clip/clip.py:119-133);
  * the returned model runs only on a CUDA sm_100 device (no `model.float()` CPU branch, clip/clip.py:137-138).
"""
from __future__ import annotations

import os
import warnings
from typing import List, Union

import torch

from .bpe_tokenizer import BPETokenizer
from .model import build_model

_MODEL_FILES = {
    "RN50": "RN50.pt", "RN101": "RN101.pt", "RN50x4": "RN50x4.pt", "RN50x16": "RN50x16.pt",
    "ViT-B/32": "ViT-B-32.pt", "ViT-B/16": "ViT-B-16.pt", "ViT-L/14": "ViT-L-14.pt",
}
_tokenizer = None


def available_models() -> List[str]:
    """Names accepted by ``load`` (clip/clip.py:87-89)."""
    return list(_MODEL_FILES.keys())


def _transform(n_px: int):
    """Host-side preprocessing, clip/clip.py:77-84: bicubic resize, centre crop, RGB, tensor, CLIP mean/std."""
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor
    return Compose([
        Resize(n_px, interpolation=InterpolationMode.BICUBIC),
        CenterCrop(n_px),
        lambda image: image.convert("RGB"),
        ToTensor(),
        Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)),
    ])


class GPUTransform:
    """The same `_transform(n_px)` on the GPU (pc_preprocess_image): PIL image or HxWx3 uint8 array -> CUDA tensor
    [3, n_px, n_px], bit-identical to the host pipeline above (Pillow's 8-bit bicubic resampler restated in integer
    arithmetic). For callers that already hold decoded pixels (the toolkit's cropped objects, video frames): the
    DataLoader workers of main.py keep the host transform, because JPEG decoding lives there anyway."""

    def __init__(self, n_px: int, device: Union[str, torch.device] = "cuda", dtype: torch.dtype = torch.float32):
        self.n_px, self.device, self.dtype = n_px, torch.device(device), dtype

    def __call__(self, image, out: torch.Tensor = None) -> torch.Tensor:
        import numpy as np
        from .. import _native as nat
        if isinstance(image, torch.Tensor):
            rgb = image
        else:
            if hasattr(image, "convert"):  # PIL.Image
                if image.mode not in ("RGB", "L"):
                    # the reference resizes BEFORE convert("RGB"): Pillow resamples palette / bilevel images with NEAREST
                    # and alpha images in premultiplied form, which this RGB kernel does not reproduce
                    raise ValueError(f"GPUTransform handles RGB / L images; use the host `preprocess` for mode {image.mode!r}")
                image = image.convert("RGB")  # for L: resize-then-replicate == replicate-then-resize, channel by channel
            arr = np.ascontiguousarray(np.asarray(image))
            if arr.ndim == 2:
                arr = np.repeat(arr[:, :, None], 3, axis=2)
            rgb = torch.from_numpy(arr[:, :, :3].copy())
        return nat.preprocess_image(rgb.to(self.device, non_blocking=True), self.n_px, out=out, dtype=self.dtype)


def _read_state_dict(path: str):
    try:
        return torch.jit.load(path, map_location="cpu").eval().state_dict()
    except RuntimeError:
        sd = torch.load(path, map_location="cpu", weights_only=False)
        return sd.state_dict() if hasattr(sd, "state_dict") else sd


def load(name: str, device: Union[str, torch.device] = "cuda", jit: bool = False, download_root: str = None):
    """Returns (model, preprocess) like clip/clip.py:92-139."""
    if jit:
        warnings.warn("jit=True is ignored: the encoders run on libprotoclip_b200, not TorchScript")
    if name.startswith("synthetic:"):
        from .. import synthetic
        parts = name.split(":")
        state_dict = synthetic.make_state_dict(parts[1], int(parts[2]) if len(parts) > 2 else 0)
    elif name in _MODEL_FILES:
        path = os.path.join(download_root or os.path.expanduser("~/.cache/clip"), _MODEL_FILES[name])
        if not os.path.isfile(path):
            raise RuntimeError(f"Model {name}: checkpoint {path} not found and this build cannot download it; "
                               f"place the OpenAI checkpoint there or pass a state-dict path")
        state_dict = _read_state_dict(path)
    elif os.path.isfile(name):
        state_dict = _read_state_dict(name)
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    model = build_model(state_dict).to(device)
    model.preprocess_gpu = GPUTransform(model.visual.input_resolution, device)  # same arithmetic, on the device
    return model, _transform(model.visual.input_resolution)


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False) -> torch.LongTensor:
    """clip/clip.py:194-230: [SOT] + BPE(text) + [EOT], zero-padded to context_length."""
    global _tokenizer
    if _tokenizer is None:
        _tokenizer = BPETokenizer()
    if isinstance(texts, str):
        texts = [texts]
    sot, eot = _tokenizer.encoder["<|startoftext|>"], _tokenizer.encoder["<|endoftext|>"]
    result = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        ids = [sot] + _tokenizer.encode(text) + [eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {text} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = eot
        result[i, : len(ids)] = torch.tensor(ids)
    return result
