"""One ViT-B/16 ResidualAttentionBlock (QKV, attention, out_proj, c_fc, c_proj with the towers' epilogues) at the
benchmark's micro-batch, a few times in a row: the target of `ncu --set full -k regex:gemm_tn_kernel|attention6`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proto_clip_b200 import _native as nat  # noqa: E402
from proto_clip_b200 import synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
c = synthetic.arch_config("ViT-B/16")
d, L = c["vision_width"], 197
ctx = nat.Context(torch.device("cuda:0"))
ctx.bind_visual(synthetic.make_state_dict("ViT-B/16", 0))
x = (torch.randn(B * L, d, device="cuda") * 0.5).half()
ctx.resblock_forward_parts(nat.PC_TOWER_VISUAL, 0, x, B, L, False, 31, False)
for _ in range(iters):
    ctx.resblock_forward_parts(nat.PC_TOWER_VISUAL, 0, x, B, L, False, 31, True)
torch.cuda.synchronize()
print("block driver done")
