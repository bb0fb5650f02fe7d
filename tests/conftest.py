import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100) device; run on the B200 box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def rel_err(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def mean_rel_err(got, ref):
    """mean |got - ref| / mean |ref|: unlike rel_err (max error over the reference's RANGE) it does not let a few
    large components hide errors on the many small ones."""
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).abs().mean() / ref.abs().mean().clamp_min(1e-12)).item()


def golden_images(fx):
    """Regenerate the images a tower fixture was made with (tests/golden/make_golden.py:images_for)."""
    from proto_clip_b200 import synthetic
    c = synthetic.arch_config(fx["arch"])
    bases = synthetic.class_bases(5, c["image_resolution"], seed=fx["image_seed"])
    return synthetic.class_structured_images(bases, torch.arange(fx["B"]) % 5, seed=fx["image_seed"] + 1)
