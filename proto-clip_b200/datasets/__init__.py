"""Dataset front door of main.py: ``build_dataset``, ``build_data_loader``, ``get_random_train_tfm``, ``ImageNet``
(reference datasets/__init__.py:16-34, datasets/utils.py build_data_loader, datasets/imagenet.py:8-23,201-236).

Reading datasets (split files, JPEG decode, PIL / torchvision augmentation, 8 DataLoader workers) is host-side code
outside the hot path this repo rebuilds, so it is not re-implemented:

* ``synthetic[:N[:Q]]`` is a built-in alias: N classes (default 16), ``shots`` support images per class and Q queries
  per split (default 256) of class-structured synthetic images (``synthetic.class_structured_images``),
  generated on the GPU at the bound model's input resolution. It lets ``main.py`` run end to end with no data on disk
  (``--backbone synthetic:<arch>``), which is how the CLI is exercised on the B200 box.
* the reference's eleven aliases (caltech101 ... ucf101, imagenet, fewsol) are delegated, unchanged, to the reference's
  own ``datasets`` package when ``PROTOCLIP_REFERENCE_ROOT`` points at a checkout of it; without one a RuntimeError says so.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from typing import Iterator, List, Tuple

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


def get_random_train_tfm():
    """datasets/imagenet.py:8-23 (RandomResizedCrop + flip + CLIP normalisation); unused by the synthetic alias."""
    if "PROTOCLIP_REFERENCE_ROOT" not in os.environ:
        return None
    return _ref_sub("imagenet").get_random_train_tfm()


def ImageNet(root_path: str, shots: int, preprocess):
    """datasets/imagenet.py:201-236."""
    return _ref_sub("imagenet").ImageNet(root_path, shots, preprocess)
