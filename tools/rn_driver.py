"""RN50x16 encode_image a few times (target of ncu launch lists).  python tools/rn_driver.py [batch] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proto_clip_b200 import _native as nat  # noqa: E402
from proto_clip_b200 import synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = nat.Context(torch.device("cuda:0"))
ctx.bind_visual(synthetic.make_state_dict("RN50x16", 0))
x = torch.randn(B, 3, 384, 384, device="cuda")
torch.cuda.synchronize()
print("BOUND", flush=True)
for _ in range(iters):
    ctx.encode_image(x, l2norm=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ctx.encode_image(x, l2norm=True)
e1.record()
torch.cuda.synchronize()
print(f"RN50x16 B={B}: {B * iters / e0.elapsed_time(e1) * 1e3:.0f} img/s")
