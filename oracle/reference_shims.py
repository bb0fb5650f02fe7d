"""Import the REAL reference (read-only tree at /root/reference, else the byte-for-byte snapshot of its hot-path
files that oracle/make_ref.py leaves in oracle/_ref/) on a CPU-only box.  TEST INFRASTRUCTURE ONLY.

Used by tests/golden/make_golden.py (fixture generation), by tests that pin oracle/protoclip_oracle.py against
the live reference, and by bench.py --impl reference / cpu_baseline. /root/reference does not exist on the GPU box;
the snapshot (git-ignored, not gpurun-ignored) travels there, so the CPU arm times the reference's own modules.

The shims follow SURVEY.md §8(c): four stub modules for packages that are not installed here (ftfy,
matplotlib, info_nce, gdown) — none of them touches the hot path's arithmetic (ftfy.fix_text is the identity
on ASCII class names; the others are plotting / training-loss / download helpers).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

SNAPSHOT_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root() -> str:
    full = os.environ.get("PROTOCLIP_REFERENCE_ROOT", "/root/reference")
    if os.path.isfile(os.path.join(full, "clip", "model.py")):
        return full
    return SNAPSHOT_ROOT


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "clip", "model.py"))


def is_snapshot() -> bool:
    """True when the modules come from oracle/_ref/ (hot-path files only) rather than a full checkout."""
    return os.path.abspath(REFERENCE_ROOT) == os.path.abspath(SNAPSHOT_ROOT)


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def _install_stubs() -> None:
    import torch

    for name in ("ftfy", "matplotlib", "info_nce", "gdown"):
        try:
            importlib.import_module(name)
        except Exception:
            if name == "ftfy":
                _stub("ftfy", fix_text=lambda s: s)
            elif name == "matplotlib":
                class _Anything:
                    def __getattr__(self, _):
                        return _Anything()

                    def __call__(self, *a, **k):
                        return _Anything()

                    def __iter__(self):
                        return iter(())

                mpl = _stub("matplotlib", use=lambda *a, **k: None)
                plt = _stub("matplotlib.pyplot")

                def _plt_getattr(n):
                    if n.startswith("__"):
                        raise AttributeError(n)
                    return _Anything()

                plt.__getattr__ = _plt_getattr  # type: ignore[attr-defined]
                mpl.pyplot = plt
            elif name == "info_nce":
                class InfoNCE(torch.nn.Module):
                    def __init__(self, *a, **k):
                        super().__init__()

                    def forward(self, *a, **k):
                        raise RuntimeError("info_nce stub: training losses are outside the oracle's scope")

                _stub("info_nce", InfoNCE=InfoNCE)
            else:
                _stub(name)


def load_clip_model_module():
    """reference clip/model.py loaded by file path (needs only numpy + torch)."""
    spec = importlib.util.spec_from_file_location("_ref_clip_model", os.path.join(REFERENCE_ROOT, "clip", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ReferenceModules:
    """Handles to the reference's own modules: clip_model (clip/model.py), clip (package), model (model.py),
    utils (utils.py). Imported with REFERENCE_ROOT temporarily at the front of sys.path and removed from
    sys.modules afterwards so the drop-in shells of this repo (same module names) are not shadowed."""

    def __init__(self):
        if not available():
            raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
        _install_stubs()
        saved = {k: sys.modules.pop(k) for k in list(sys.modules)
                 if k in ("clip", "model", "utils", "datasets") or k.startswith(("clip.", "datasets."))}
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            self.clip = importlib.import_module("clip")
            self.clip_model = importlib.import_module("clip.model")
            self.utils = importlib.import_module("utils")
            self.model = importlib.import_module("model")
        finally:
            sys.path.remove(REFERENCE_ROOT)
            for k in list(sys.modules):
                if k in ("clip", "model", "utils", "datasets") or k.startswith(("clip.", "datasets.")):
                    sys.modules.pop(k)
            sys.modules.update(saved)


_cached = None


def reference() -> ReferenceModules:
    global _cached
    if _cached is None:
        _cached = ReferenceModules()
    return _cached
