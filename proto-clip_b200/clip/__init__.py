"""Drop-in for the reference's vendored ``clip`` package (clip/__init__.py: `from .clip import *`)."""
from .clip import available_models, load, tokenize  # noqa: F401
from .model import CLIP, build_model, convert_weights  # noqa: F401
