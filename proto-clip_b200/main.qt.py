"""Drop-in for the reference's main.qt.py CLI (the Q^T variant named by config C5: sun397, RN50x16, conv-2x).

On the inference path main.qt.py is main.py with an un-rounded alpha grid (main.qt.py:109-111) and trained
`_v/_t/_a.pt` files under `best-alpha-beta/` instead of `alpha-beta/` (main.qt.py:292,327); its episodic training
loop (main.qt.py:186-316) is outside the inference hot path, as in main.py. Same flags, YAML keys and cache layout.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("proto_clip_b200_main", os.path.join(_HERE, "main.py"))
_main = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_main)
_main.VARIANT = "qt"

if __name__ == "__main__":
    sys.exit(_main.main())
