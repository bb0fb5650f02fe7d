"""Snapshot the reference's own hot-path modules into oracle/_ref/.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_ref.py            # needs /root/reference (or $PROTOCLIP_REFERENCE_ROOT)

The reference (IRVLUTD/Proto-CLIP) is pure Python: there is nothing to compile. What the CPU baseline, the
`--impl reference` arm of bench.py and the tokenizer / prompt tests need on the GPU box (where /root/reference
does not exist) is the reference's OWN code, unmodified:

    clip/__init__.py clip/clip.py clip/model.py clip/simple_tokenizer.py clip/bpe_simple_vocab_16e6.txt.gz
    model.py utils.py datasets/imagenet.py   (class names + prompt templates, datasets/imagenet.py:26-199)

They are copied byte for byte (sha256 recorded in MANIFEST.json) into oracle/_ref/, which is git-ignored -- the
reference's sources never enter this repository's history -- but NOT gpurun-ignored, so the snapshot travels to the
GPU box like a built .so. oracle/reference_shims.py imports from it when /root/reference is absent; bench.py then
reports `cpu_baseline.kind = "reference"`. __graft_entry__.build() runs this recipe whenever the reference tree is
mounted.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["clip/__init__.py", "clip/clip.py", "clip/model.py", "clip/simple_tokenizer.py",
         "clip/bpe_simple_vocab_16e6.txt.gz", "model.py", "utils.py", "datasets/imagenet.py"]


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def snapshot(src_root: str = None, dest: str = DEST) -> dict:
    src_root = src_root or os.environ.get("PROTOCLIP_REFERENCE_ROOT", "/root/reference")
    if not os.path.isfile(os.path.join(src_root, "clip", "model.py")):
        raise FileNotFoundError(f"reference tree not found at {src_root}")
    manifest = {"source": src_root, "files": {}}
    for rel in FILES:
        s, d = os.path.join(src_root, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        os.chmod(d, 0o644)
        manifest["files"][rel] = sha256(d)
    head = os.path.join(src_root, ".git", "HEAD")
    manifest["note"] = "byte-for-byte copies of the reference's files; never edited, never committed"
    if os.path.isfile(head):
        manifest["git_head"] = open(head).read().strip()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


def verify(dest: str = DEST) -> bool:
    """True when oracle/_ref/ holds every file of the manifest with the recorded hash."""
    mp = os.path.join(dest, "MANIFEST.json")
    if not os.path.isfile(mp):
        return False
    files = json.load(open(mp)).get("files", {})
    return set(files) == set(FILES) and all(
        os.path.isfile(os.path.join(dest, r)) and sha256(os.path.join(dest, r)) == h for r, h in files.items())


if __name__ == "__main__":
    m = snapshot(sys.argv[1] if len(sys.argv) > 1 else None)
    print(f"oracle/_ref: {len(m['files'])} files from {m['source']}")
