"""Dataset front door of main.py: ``build_dataset``, ``build_data_loader``, ``get_random_train_tfm``, ``ImageNet``
(reference datasets/__init__.py:16-34, datasets/utils.py build_data_loader, datasets/imagenet.py:8-23,201-236).

Reading datasets (split files, JPEG decode, PIL / torchvision augmentation, 8 DataLoader workers) is host-side code
outside the hot path this repo rebuilds, so it is not re-implemented:

* ``synthetic[:N[:Q]]`` is a built-in alias: N classes (default 16), ``shots`` support images per class and Q queries
  per split (default 256) of class-structured synthetic images (``synthetic.class_structured_images``),
  generated on the GPU at the bound model's input resolution. It lets ``main.py`` run end to end with no data on disk
  (``--backbone synthetic:<arch>``), which is how the CLI is exercised on the B200 box.
* ``get_random_train_tfm()`` is the reference's five-line torchvision Compose; ``get_random_train_tfm(device=...)`` /
  ``GPUTrainTransform`` runs the same augmentation on the GPU for decoded images, bit-identical under the same seed.
* the reference's eleven aliases (caltech101 ... ucf101, imagenet, fewsol) are delegated, unchanged, to the reference's
  own ``datasets`` package when ``PROTOCLIP_REFERENCE_ROOT`` points at a checkout of it; without one a RuntimeError says so.
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
from typing import Iterator, List, Tuple, Union

import torch

_resolution = 224


def configure(input_resolution: int) -> None:
    """main.py tells the synthetic loaders which resolution the bound CLIP expects (clip/clip.py:139)."""
    global _resolution
    _resolution = int(input_resolution)


# ------------------------------------------------------------------------------------------ synthetic alias
class _SyntheticSplit:
    """A list-like split: labels plus the seed its images are drawn with."""

    def __init__(self, labels: torch.Tensor, seed: int, num_classes: int):
        self.labels, self.seed, self.num_classes = labels, seed, num_classes

    def __len__(self) -> int:
        return int(self.labels.numel())


class SyntheticDataset:
    template = ["a photo of a {}."]

    def __init__(self, spec: str, num_shots: int):
        parts = spec.split(":")
        n = int(parts[1]) if len(parts) > 1 else 16
        q = int(parts[2]) if len(parts) > 2 else 256
        self.classnames = [f"class_{i}" for i in range(n)]
        self.train_x = _SyntheticSplit(torch.arange(n).repeat_interleave(num_shots), 2, n)
        self.train = self.train_x
        self.val = _SyntheticSplit(torch.arange(q) % n, 5, n)
        self.test = _SyntheticSplit(torch.arange(q) % n, 3, n)


class _SyntheticLoader:
    """Yields (images fp32 [b,3,R,R] on the current CUDA device, labels int64) batches; the whole split is drawn from
    one seeded generator so that batch size does not change the images."""

    def __init__(self, split: _SyntheticSplit, batch_size: int):
        self.split, self.batch_size = split, batch_size

    def __len__(self) -> int:
        return (len(self.split) + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        return self.iter_range(0, len(self))

    def iter_range(self, lo: int, hi: int) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Batches [lo, hi) only (dist.sharded_batches): a batch's images depend on its index alone, so a rank that
        starts in the middle of the split draws exactly what the single-process run draws there."""
        try:
            from .. import synthetic
        except ImportError:  # pragma: no cover
            from proto_clip_b200 import synthetic
        dev = torch.device("cuda", torch.cuda.current_device())
        bases = synthetic.class_bases(self.split.num_classes, _resolution, seed=1, device=dev)
        for i in range(lo * self.batch_size, min(hi * self.batch_size, len(self.split)), self.batch_size):
            labels = self.split.labels[i:i + self.batch_size]
            images = synthetic.class_structured_images(bases, labels.to(dev), seed=1000 * self.split.seed + i)
            yield images, labels


# ------------------------------------------------------------------------------------------ reference delegation
_ref_pkg = None


def _reference_datasets():
    global _ref_pkg
    if _ref_pkg is not None:
        return _ref_pkg
    root = os.environ.get("PROTOCLIP_REFERENCE_ROOT", "")
    init = os.path.join(root, "datasets", "__init__.py")
    if not root or not os.path.isfile(init):
        raise RuntimeError(
            "dataset readers are host-side code outside this build: set PROTOCLIP_REFERENCE_ROOT to a checkout of "
            "IRVLUTD/Proto-CLIP to use its datasets/ package unchanged, or use the built-in `--dataset synthetic[:N[:Q]]`")
    spec = importlib.util.spec_from_file_location("protoclip_reference_datasets", init,
                                                  submodule_search_locations=[os.path.dirname(init)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    _ref_pkg = mod
    return mod


def _ref_sub(name: str):
    """Submodule of the reference's datasets package (datasets/utils.py, datasets/imagenet.py)."""
    pkg = _reference_datasets()
    return importlib.import_module(f"{pkg.__name__}.{name}")


def build_dataset(dataset: str, root_path: str, shots: int):
    """datasets/__init__.py:31-34."""
    if dataset.startswith("synthetic"):
        return SyntheticDataset(dataset, shots)
    return _reference_datasets().build_dataset(dataset, root_path, shots)


def build_data_loader(data_source=None, batch_size=64, input_size=224, tfm=None, is_train=True, shuffle=False,
                      **kwargs):
    """datasets/utils.py build_data_loader (same keywords; extra DataLoader keywords pass through)."""
    if isinstance(data_source, _SyntheticSplit):
        return _SyntheticLoader(data_source, batch_size)
    return _ref_sub("utils").build_data_loader(data_source=data_source, batch_size=batch_size, input_size=input_size,
                                               tfm=tfm, is_train=is_train, shuffle=shuffle, **kwargs)


TRAIN_SCALE = (0.5, 1.0)              # datasets/imagenet.py:16-17
TRAIN_RATIO = (3.0 / 4.0, 4.0 / 3.0)  # torchvision's RandomResizedCrop default


class GPUTrainTransform:
    """`get_random_train_tfm()` on the GPU (pc_preprocess_train_image): PIL image or HxWx3 uint8 array / tensor -> CUDA
    tensor [3, size, size], bit-identical to the host Compose under the same torch seed. The draws are taken on the
    host from torch's global generator in torchvision's order (RandomResizedCrop.get_params: up to ten (area, log-ratio)
    pairs, two randint for the position; RandomHorizontalFlip: one rand), the pixels never leave the device: Pillow's
    resampler of the cropped box, the mirror, ToTensor and Normalize are two launches. For callers that hold decoded
    support images (build_cache_model re-augments the same N*K images `augment_epoch` times, utils.py:303-310)."""

    def __init__(self, size: int = 224, scale=TRAIN_SCALE, ratio=TRAIN_RATIO, p_flip: float = 0.5,
                 device: Union[str, torch.device] = "cuda", dtype: torch.dtype = torch.float32):
        self.size, self.scale, self.ratio, self.p_flip = size, tuple(scale), tuple(ratio), p_flip
        self.device, self.dtype = torch.device(device), dtype
        lr = torch.log(torch.tensor(self.ratio))  # fp32 like torchvision's; no generator state involved
        self._log_ratio = (float(lr[0]), float(lr[1]))
        self._draw = torch.empty(1)

    def get_params(self, height: int, width: int):
        """torchvision RandomResizedCrop.get_params -> (top, left, h, w)."""
        area = height * width
        log_ratio, draw = self._log_ratio, self._draw
        for _ in range(10):
            target_area = area * draw.uniform_(self.scale[0], self.scale[1]).item()
            aspect_ratio = torch.exp(draw.uniform_(log_ratio[0], log_ratio[1])).item()
            w = int(round(math.sqrt(target_area * aspect_ratio)))
            h = int(round(math.sqrt(target_area / aspect_ratio)))
            if 0 < w <= width and 0 < h <= height:
                top = torch.randint(0, height - h + 1, size=(1,)).item()
                left = torch.randint(0, width - w + 1, size=(1,)).item()
                return top, left, h, w
        in_ratio = float(width) / float(height)  # central fallback
        if in_ratio < min(self.ratio):
            w = width
            h = int(round(w / min(self.ratio)))
        elif in_ratio > max(self.ratio):
            h = height
            w = int(round(h * max(self.ratio)))
        else:
            w, h = width, height
        return (height - h) // 2, (width - w) // 2, h, w

    def __call__(self, image, out: torch.Tensor = None) -> torch.Tensor:
        import numpy as np
        from .. import _native as nat
        if isinstance(image, torch.Tensor):
            rgb = image
        else:
            if hasattr(image, "convert"):  # PIL.Image; the reference's loader hands over convert("RGB") images
                if image.mode != "RGB":
                    raise ValueError(f"GPUTrainTransform handles RGB images; got mode {image.mode!r}")
            rgb = torch.from_numpy(np.array(image))  # a writable copy
        box = self.get_params(int(rgb.shape[0]), int(rgb.shape[1]))
        flip = bool(torch.rand(1) < self.p_flip)
        return nat.preprocess_train_image(rgb.to(self.device, non_blocking=True), box, flip, self.size, out=out,
                                          dtype=self.dtype)


class GPUAugmentedLoader:
    """The support-image loader of build_cache_model (utils.py:284-332; main.py builds it with `get_random_train_tfm()`,
    shuffle=False) for a few-shot split that FITS IN HBM: every image is decoded once (PIL on the host, like the
    reference's `read_image`: `Image.open(path).convert("RGB")`), kept on the device as uint8 [H, W, 3] (ImageNet
    16-shot: 16 000 images, about 9 GB of the 180), and every pass over the loader -- one per `augment_epoch` -- re-draws
    the augmentation and runs it on the GPU (`GPUTrainTransform`). Yields (images [b, 3, size, size] on the device,
    labels int64 [b]) like the reference's loader. The draws come from torch's global generator in image order, so a
    pass reproduces, bit for bit, the reference's loader at num_workers=0 under the same seed.

    `data_source`: a sequence of Datum-like items (`.impath`, `.label`: datasets/utils.py), or of (image, label) pairs
    with image a PIL image / HxWx3 uint8 array / tensor."""

    def __init__(self, data_source, batch_size: int = 64, tfm: "GPUTrainTransform" = None,
                 device: Union[str, torch.device] = "cuda"):
        self.data_source, self.batch_size = data_source, int(batch_size)
        self.device = torch.device(device)
        self.tfm = tfm if tfm is not None else GPUTrainTransform(224, device=self.device)
        self._pixels: List[torch.Tensor] = []
        self._labels: List[int] = []

    def __len__(self) -> int:
        return (len(self.data_source) + self.batch_size - 1) // self.batch_size

    def _decode(self) -> None:
        import numpy as np
        for item in self.data_source:
            if hasattr(item, "impath"):
                from PIL import Image
                image, label = Image.open(item.impath).convert("RGB"), item.label
            else:
                image, label = item
            if hasattr(image, "convert"):
                image = image.convert("RGB")
            if not isinstance(image, torch.Tensor):
                image = torch.from_numpy(np.array(image))  # a writable copy
            if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[-1] != 3:
                raise ValueError(f"GPUAugmentedLoader: expected RGB uint8 [H, W, 3] images, got {image.dtype} {tuple(image.shape)}")
            self._pixels.append(image.to(self.device))
            self._labels.append(int(label))

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        return self.iter_range(0, len(self))

    def iter_range(self, lo: int, hi: int) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Batches [lo, hi) (dist.sharded_batches). The draws of the batches before `lo` are not consumed: ranks of a
        torchrun job draw independently (like DataLoader workers do), a single process reproduces the reference."""
        if not self._pixels:
            self._decode()
        n = len(self._pixels)
        for i in range(lo * self.batch_size, min(hi * self.batch_size, n), self.batch_size):
            j = min(i + self.batch_size, n)
            out = torch.empty((j - i, 3, self.tfm.size, self.tfm.size), dtype=self.tfm.dtype, device=self.device)
            for k in range(i, j):
                self.tfm(self._pixels[k], out=out[k - i])
            yield out, torch.tensor(self._labels[i:j], dtype=torch.int64)


def get_random_train_tfm(device: Union[str, torch.device, None] = None):
    """datasets/imagenet.py:8-23: RandomResizedCrop(224, scale (0.5, 1), BICUBIC) + RandomHorizontalFlip + ToTensor + CLIP
    normalisation. Without `device` the host Compose the reference hands its DataLoader workers; with one, the same
    transform on that GPU (`GPUTrainTransform`)."""
    if device is not None:
        return GPUTrainTransform(224, device=device)
    import torchvision.transforms as T
    return T.Compose([
        T.RandomResizedCrop(size=224, scale=TRAIN_SCALE, interpolation=T.InterpolationMode.BICUBIC),
        T.RandomHorizontalFlip(p=0.5),
        T.ToTensor(),
        T.Normalize(mean=(0.48145466, 0.4578275, 0.40821073), std=(0.26862954, 0.26130258, 0.27577711)),
    ])


def ImageNet(root_path: str, shots: int, preprocess):
    """datasets/imagenet.py:201-236."""
    return _ref_sub("imagenet").ImageNet(root_path, shots, preprocess)
