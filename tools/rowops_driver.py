"""Launches every non-contraction kernel of the query path once per iteration at the benchmark's shapes, for
`ncu --set full -k regex:...` (tools/ncu_rowops.sh) and for a CUDA-event timing of each on its own:
patchify / embed_ln_pre / row_stats / layernorm (ln_post) / l2norm inside one ViT-B/16 encode_image of 512 images,
ln_f16 + proto_softmax (fc adapter + P at Q = 1024, N = 1000), adapter_conv (conv-3x, D = 768), prototypes
(N = 1000, K = 16, D = 512), preprocess (256 frames 480 x 640 -> 224)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proto_clip_b200 import _native as nat  # noqa: E402
from proto_clip_b200 import pipeline, synthetic  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
sd = synthetic.make_state_dict("ViT-B/16", 0)
ctx = nat.Context(dev)
ctx.bind_visual(sd)
imgs = torch.randn(512, 3, 224, 224, device=dev)
N, K, D, Q = 1000, 16, 512, 1024
V = nat.l2_normalize(torch.randn(N * K, D, device=dev).half())
T = nat.l2_normalize(torch.randn(N, D, device=dev).half())
afc = synthetic.make_adapter_state_dict("fc", D, seed=4, out_gain=synthetic.trained_like_gain(D))
head = pipeline.build_head_state(V, T, N, K, "fc", afc, 0.5, 12.0)
clf = pipeline.FewShotClassifier(ctx, head)
ac = {k: v.to(dev) for k, v in synthetic.make_adapter_state_dict("conv", 768, seed=4).items()}
q768 = nat.l2_normalize(torch.randn(Q, 768, device=dev).half())
frames = torch.randint(0, 256, (256, 480, 640, 3), device=dev, dtype=torch.uint8)
out_pp = torch.empty(256, 3, 224, 224, device=dev)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    f = ctx.encode_image(imgs, l2norm=True)
    clf.classify_features(f[:Q] if f.shape[0] >= Q else f.repeat(2, 1)[:Q], want_p=True)
    nat.adapter_conv_forward(ac, "conv-3x", q768)
    nat.build_prototypes(V, N, K, True)
    nat.preprocess_image(frames, 224, out=out_pp)
torch.cuda.synchronize()
print("rowops driver done")
