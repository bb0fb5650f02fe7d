"""RN50x16 encode_image at several images-per-pass values on one box (same-box A/B of `kRnMicroBatch`).
python tools/rn_pass_sweep.py [batch] [iters]  -> one line per pass size, interleaved twice (clock drift under the cap)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proto_clip_b200 import _native as nat  # noqa: E402
from proto_clip_b200 import synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = nat.Context(torch.device("cuda:0"))
ctx.bind_visual(synthetic.make_state_dict("RN50x16", 0))
x = torch.randn(B, 3, 384, 384, device="cuda")
ref = None
for rnd in range(2):
    for mb in (128, 192, 256, 512, 64):
        if mb > B:
            continue
        f = ctx.encode_image(x, l2norm=True, micro_batch=mb)
        torch.cuda.synchronize()
        if ref is None:
            ref = f.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            f = ctx.encode_image(x, l2norm=True, micro_batch=mb)
        e1.record()
        torch.cuda.synchronize()
        print(f"RN50x16 B={B} pass={mb} round {rnd}: {B * iters / e0.elapsed_time(e1) * 1e3:.0f} img/s  "
              f"bit-identical to pass=128: {bool(torch.equal(f, ref))}", flush=True)
