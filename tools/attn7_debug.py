"""Bring-up helper: which rows of the long-sequence attention differ from the fp32 reference, for a list of lengths."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proto_clip_b200 import _native as nat  # noqa: E402

for (B, L, heads) in [(2, 288, 1), (2, 384, 1), (2, 385, 1), (2, 392, 1), (2, 480, 1), (2, 300, 2), (2, 577, 2), (40, 577, 16)]:
    torch.manual_seed(L)
    d = heads * 64
    qkv = torch.randn(B * L, 3 * d, device="cuda").half()
    try:
        got = nat.attention(qkv, B, L, heads, False)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(B, L, heads, "ERROR", str(e)[:200])
        break
    q, k, v = [t.reshape(B, L, heads, 64).permute(0, 2, 1, 3).float() for t in qkv.split(d, dim=1)]
    ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, d)
    err = (got.float() - ref).abs().amax(dim=1)
    bad = torch.nonzero(~(err < 2e-3)).flatten().tolist()
    print(f"B={B} L={L} h={heads}: blocks={-(-(L - (L % 96 == 1)) // 96)} bad rows {len(bad)}: {[(r // L, r % L) for r in bad[:12]]}", flush=True)
