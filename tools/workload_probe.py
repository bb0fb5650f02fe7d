"""How classifiable is the synthetic 1000-way workload through a random-init ViT-B/16? (GPU, CUDA path)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proto_clip_b200 import _native as nat, synthetic

dev = torch.device("cuda:0")
arch = "ViT-B/16"; c = synthetic.arch_config(arch); D = c["embed_dim"]
sd = synthetic.make_state_dict(arch, 0)
ctx = nat.Context(dev); ctx.bind_visual(sd)
N, K, Q = 1000, 16, 1024
bases = synthetic.class_bases(N, 224, seed=1, device=dev)
for noise in (0.5, 0.25, 0.1):
    feats = []
    for n0 in range(0, N, 64):
        labels = torch.arange(n0, min(n0 + 64, N), device=dev).repeat_interleave(K)
        feats.append(ctx.encode_image(synthetic.class_structured_images(bases, labels, seed=2 + n0, noise=noise), l2norm=True))
    V = torch.cat(feats)
    zi, zn = nat.build_prototypes(V, N, K, True)
    labels = (torch.arange(Q, device=dev) * 7 + 13) % N
    f = ctx.encode_image(synthetic.class_structured_images(bases, labels, seed=1000, noise=noise), l2norm=True)
    sim = f.float() @ zi.float().t()
    same = sim[torch.arange(Q), labels]
    other_max = sim.scatter(1, labels[:, None], -1).max(1).values
    p, am, pm = nat.proto_classify(f, zi, zi, zn, zn, 0.5, 12.0)
    t2 = p.topk(2, 1).values
    print(f"noise {noise}: acc(sim) {(sim.argmax(1) == labels).float().mean():.4f} acc(P) {(am == labels).float().mean():.4f} "
          f"same {same.mean():.5f} other_max {other_max.mean():.5f} other_mean {((sim.sum(1) - same) / (N - 1)).mean():.5f} "
          f"margin min/med {(t2[:, 0] - t2[:, 1]).min():.3e} {(t2[:, 0] - t2[:, 1]).median():.3e}", flush=True)
