"""CPU-side checks of the C ABI: the shared library loads, exports every symbol include/protoclip_b200.h
declares, and fails loudly (never silently falls back) when there is no sm_100 device."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from proto_clip_b200 import _native as nat

HEADER = os.path.join(ROOT, "include", "protoclip_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = nat.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(nat.SYMBOLS) == syms, "ctypes binding list and header drifted apart"


def test_version_and_error_string():
    lib = nat.load_library()
    assert lib.pc_version() == 100
    assert isinstance(lib.pc_last_error(), bytes)


def test_struct_layouts_match_header():
    # 12 pointers per resblock; vit: 6 ints + 8 pointers + blocks pointer; text: 6 ints + 5 pointers + blocks
    assert ctypes.sizeof(nat.ResblockWeights) == 12 * 8
    assert ctypes.sizeof(nat.VitWeights) == 6 * 4 + 9 * 8
    assert ctypes.sizeof(nat.TextWeights) == 6 * 4 + 6 * 8
    assert ctypes.sizeof(nat.AdapterFCWeights) == 6 * 8 + 8  # int + padding
    assert ctypes.sizeof(nat.AdapterConvWeights) == 9 * 8
    # ModifiedResNet: conv+bn = 5 pointers; bottleneck = 3 ints (+ pad) + 4 conv+bn; rn = 4 + 4 ints, 3 stem conv+bn,
    # blocks pointer, 9 attention-pool pointers
    assert ctypes.sizeof(nat.ConvBnWeights) == 5 * 8
    assert ctypes.sizeof(nat.BottleneckWeights) == 16 + 4 * 40
    assert ctypes.sizeof(nat.RnWeights) == 8 * 4 + 3 * 40 + 8 + 9 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    lib = nat.load_library()
    h = ctypes.c_void_p()
    rc = lib.pc_ctx_create(0, ctypes.byref(h))
    assert rc < 0 and not h.value
    assert len(lib.pc_last_error()) > 0
    with pytest.raises(nat.NativeError):
        nat.Context(torch.device("cpu"))
    with pytest.raises(nat.NativeError):
        nat.linear(torch.zeros(8, 8, dtype=torch.float16), torch.zeros(8, 8, dtype=torch.float16))
    with pytest.raises(nat.NativeError):
        nat.l2_normalize(torch.zeros(2, 8, dtype=torch.float16))


def test_argument_validation_without_compute():
    """Workspace-size queries and argument checks run on the host and must not need a device."""
    lib = nat.load_library()
    assert lib.pc_adapter_fc_workspace_bytes(1000, 512, 4) >= 1000 * (128 * 2 * 2 + 512 * 2)
    assert lib.pc_proto_classify_workspace_bytes(50000, 1000) == 8192 * 2000 * 4
    assert lib.pc_proto_classify_workspace_bytes(10, 47) == 10 * 2 * 48 * 4
    assert lib.pc_encode_image_workspace_bytes(None, 96) == 0  # no context -> 0, not a crash
    rc = lib.pc_linear_forward(None, 0, None, 0, None, None, 0, None, 0, 4, 4, 4, 0, None)
    assert rc == -1 and b"null" in lib.pc_last_error()
    rc = lib.pc_attention_forward(None, None, 1, 600, 1, 0, None)
    assert rc == -1


def test_missing_library_message(monkeypatch):
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", "/nonexistent/libprotoclip_b200.so")
    with pytest.raises(nat.NativeError, match="no fallback"):
        nat.load_library()
