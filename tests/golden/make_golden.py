"""Generate the golden fixtures under tests/golden/ by running the REAL reference (imported from
/root/reference through oracle/reference_shims.py) on seeded inputs. Run in the authoring container only:

    python tests/golden/make_golden.py

Weights and inputs are regenerable from seeds (proto-clip_b200/synthetic.py), so only the reference's
OUTPUTS (plus small non-regenerable inputs such as checkpoint subsets) are stored. Every fixture records
two reference runs: "fp32" = `clip.load(..., device="cpu")` semantics (model.float(), clip/clip.py:137-138)
and "fp16" = the reference's GPU dtype (convert_weights fp16 modules) executed with CPU half kernels.
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shims  # noqa: E402
from proto_clip_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = reference_shims.reference()
torch.set_grad_enabled(False)


def build_ref_clip(arch: str, seed: int, fp32: bool):
    sd = synthetic.make_state_dict(arch, seed)
    model = ref.clip_model.build_model({k: v.clone() for k, v in sd.items()})  # clip/model.py:397-434
    return (model.float() if fp32 else model), sd


def synth_tokens(P: int, ctx: int, vocab: int, seed: int) -> torch.Tensor:
    """Token rows shaped like clip.tokenize output: SOT, words, EOT (= max id), zero padding."""
    gen = torch.Generator().manual_seed(seed)
    t = torch.zeros(P, ctx, dtype=torch.int64)
    for i in range(P):
        n = int(torch.randint(3, min(ctx - 2, 20), (1,), generator=gen))
        t[i, 0] = vocab - 2
        t[i, 1:1 + n] = torch.randint(1, vocab - 2, (n,), generator=gen)
        t[i, 1 + n] = vocab - 1
    return t


def images_for(arch: str, B: int, seed: int) -> torch.Tensor:
    c = synthetic.arch_config(arch)
    bases = synthetic.class_bases(5, c["image_resolution"], seed=seed)
    return synthetic.class_structured_images(bases, torch.arange(B) % 5, seed=seed + 1)


def tower_fixture(arch: str, B: int, P: int, with_fp16: bool = True, with_blocks: bool = True):
    c = synthetic.arch_config(arch)
    out = {"arch": arch, "seed": 0, "B": B, "P": P, "image_seed": 7, "token_seed": 11}
    images = images_for(arch, B, 7)
    tokens = synth_tokens(P, c["context_length"], c["vocab_size"], 11)
    out["tokens"] = tokens
    for mode in (["fp32", "fp16"] if with_fp16 else ["fp32"]):
        model, _ = build_ref_clip(arch, 0, fp32=(mode == "fp32"))
        out[f"image_features_{mode}"] = model.encode_image(images).float()   # clip/model.py:338
        out[f"text_features_{mode}"] = model.encode_text(tokens).float()     # clip/model.py:341
        if with_blocks:
            # one ResidualAttentionBlock in the reference's [L, B, d] layout (clip/model.py:187-190)
            gen = torch.Generator().manual_seed(23)
            L = (c["image_resolution"] // c["vision_patch_size"]) ** 2 + 1
            x = torch.randn(L, 3, c["vision_width"], generator=gen).half().float()
            xt = torch.randn(c["context_length"], 2, c["transformer_width"], generator=gen).half().float()
            dt = torch.float32 if mode == "fp32" else torch.float16
            out[f"vis_block0_{mode}"] = model.visual.transformer.resblocks[0](x.to(dt)).float()
            out[f"txt_block0_{mode}"] = model.transformer.resblocks[0](xt.to(dt)).float()
            out["vis_block0_in"] = x.half()
            out["txt_block0_in"] = xt.half()
    return out


def adapters_fixture():
    out = {}
    for D in (64, 512, 768, 1024):
        gen = torch.Generator().manual_seed(100 + D)
        x = torch.randn(6, D, generator=gen)
        x = (x / x.norm(dim=-1, keepdim=True)).half()
        out[f"x_{D}"] = x
        for kind in ("fc", "conv-2x", "conv-3x"):
            sd = synthetic.make_adapter_state_dict(kind, D, seed=4)
            for mode, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
                if kind == "fc":
                    m = ref.model.Adapter_FC(D, dtype=dt)              # model.py:81
                else:
                    m = ref.model.Adapter(D, c_type=kind, dtype=dt)    # model.py:12
                m.load_state_dict({k: v.to(dt) for k, v in sd.items()}, strict=False)
                out[f"{kind}_{D}_{mode}"] = m(x.to(dt)).float()
    return out


def proto_statements(V, T, K, dt):
    """main.py:399-405 verbatim semantics (tensor statements, not a function in the reference)."""
    ndim = V.shape[-1]
    zs_imgs = V.to(dt).view(-1, K, ndim)
    zs_imgs = zs_imgs / zs_imgs.norm(dim=-1, keepdim=True)
    z_img_proto = zs_imgs.mean(dim=1)
    z_img_proto = z_img_proto / z_img_proto.norm(dim=-1, keepdim=True)
    zs_text = T.to(dt)
    z_text_proto = zs_text / zs_text.norm(dim=-1, keepdim=True)
    return z_img_proto, z_text_proto


def head_fixture():
    gen = torch.Generator().manual_seed(5)
    N, K, D, Q = 12, 4, 64, 9
    centers = torch.randn(N, D, generator=gen)
    V = (centers[:, None, :] + 0.3 * torch.randn(N, K, D, generator=gen)).reshape(N * K, D).half()
    T = (centers + 0.3 * torch.randn(N, D, generator=gen)).half()
    q = centers[torch.arange(Q) % N] + 0.3 * torch.randn(Q, D, generator=gen)
    q = (q / q.norm(dim=-1, keepdim=True)).half()
    out = {"V": V, "T": T, "q": q, "N": N, "K": K}
    for mode, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
        zi, zt = proto_statements(V, T, K, dt)
        out[f"z_img_{mode}"], out[f"z_txt_{mode}"] = zi.float(), zt.float()
        for (a, b) in ((0.5, 12.0), (0.2, 5.5), (1.0, 1.0), (0.0, 20.0)):
            p = ref.utils.P(q.to(dt), zi, zt, a, b)                    # utils.py:225
            out[f"p_{mode}_{a}_{b}"] = p.float()
            out[f"pred_{mode}_{a}_{b}"] = p.max(1)[1]                  # main.py:438
    # zero-shot variant: mean without the per-shot renorm (main.py:173-176)
    zs = V.float().view(-1, K, D).mean(dim=1)
    out["z_img_zeroshot_fp32"] = zs / zs.norm(dim=-1, keepdim=True)
    return out


def ckpt_fixture(name: str, n_classes: int, kind: str, alpha: float, beta: float):
    """Class-subset of a shipped Proto-CLIP-F checkpoint, classified by the reference head
    (main.py:399-409,436-438) with the memory bank itself as queries (SURVEY.md §4 KAT)."""
    d = os.path.join(reference_shims.REFERENCE_ROOT, "pretrained_ckpt", name)
    V = torch.load(os.path.join(d, "memory_bank_v.pt"), map_location="cpu", weights_only=False).data
    T = torch.load(os.path.join(d, "memory_bank_t.pt"), map_location="cpu", weights_only=False).data
    A = torch.load(os.path.join(d, "query_adapter.pt"), map_location="cpu", weights_only=False)
    K, D = 16, V.shape[1]
    V, T = V[: n_classes * K].clone().half(), T[:n_classes].clone().half()
    out = {"V": V, "T": T, "adapter": {k: v.clone().half() for k, v in A.items()}, "K": K, "kind": kind,
           "alpha": alpha, "beta": beta}
    for mode, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
        m = ref.model.Adapter_FC(D, dtype=dt) if kind == "fc" else ref.model.Adapter(D, c_type=kind, dtype=dt)
        m.load_state_dict({k: v.to(dt) for k, v in A.items()})
        zi, zt = proto_statements(V, T, K, dt)
        q = m(V.to(dt))
        q = q / q.norm(dim=-1, keepdim=True)                           # main.py:407-409
        p = ref.utils.P(q, zi, zt, alpha, beta)
        out[f"q_{mode}"] = q[:32].float()  # first 32 rows keep the fixture small
        out[f"p_{mode}"] = p.float()
        out[f"pred_{mode}"] = p.max(1)[1]
    return out


def main(only=None):
    torch.manual_seed(0)
    torch.save(tower_fixture("tiny", B=4, P=3), os.path.join(OUT, "tower_tiny.pt"))
    print("tower_tiny done")
    torch.save(tower_fixture("small", B=3, P=3), os.path.join(OUT, "tower_small.pt"))
    print("tower_small done")
    torch.save(adapters_fixture(), os.path.join(OUT, "adapters.pt"))
    print("adapters done")
    torch.save(head_fixture(), os.path.join(OUT, "head.pt"))
    print("head done")
    torch.save(ckpt_fixture("imagenet-F", 16, "conv-2x", 0.5, 12.0), os.path.join(OUT, "ckpt_imagenet_F_16.pt"))
    torch.save(ckpt_fixture("fewsol-198-F", 24, "fc", 0.2, 12.0), os.path.join(OUT, "ckpt_fewsol_198_F_24.pt"))
    print("ckpt subsets done")
    for arch, B, P in (("ViT-B/32", 2, 2), ("ViT-B/16", 4, 3), ("ViT-L/14", 2, 2)):
        fx = tower_fixture(arch, B=B, P=P, with_fp16=True, with_blocks=False)
        torch.save(fx, os.path.join(OUT, f"tower_{arch.replace('/', '_').replace('-', '_')}.pt"))
        print(arch, "done")


def rn_fixture(arch: str, B: int, with_fp16: bool):
    """ModifiedResNet towers (clip/model.py:95-152, config C5): reference image features on seeded images."""
    out = {"arch": arch, "seed": 0, "B": B, "image_seed": 7}
    images = images_for(arch, B, 7)
    for mode in (["fp32", "fp16"] if with_fp16 else ["fp32"]):
        model, _ = build_ref_clip(arch, 0, fp32=(mode == "fp32"))
        out[f"image_features_{mode}"] = model.encode_image(images).float().clone()   # clip/model.py:338 -> :137-152
    return out


def main_rn():
    for arch, B, h in (("rn_tiny", 4, True), ("rn_small", 3, True), ("RN50", 2, False), ("RN50x16", 1, False)):
        torch.save(rn_fixture(arch, B, h), os.path.join(OUT, f"tower_{arch}.pt"))
        print(arch, "done")


def preprocess_fixture():
    """The reference's `_transform(n_px)` (clip/clip.py:77-84: PIL bicubic resize, centre crop, ToTensor, Normalize) on
    seeded RGB images of assorted sizes: inputs (uint8) and the reference's fp32 outputs."""
    import numpy as np
    from PIL import Image
    tf = {n: ref.clip.clip._transform(n) for n in (32, 64, 96)}  # clip/clip.py:77-84, the real reference function
    rng = np.random.default_rng(5)
    cases = []
    for (h, w, n) in ((37, 53, 32), (53, 37, 32), (64, 64, 64), (100, 80, 64), (45, 200, 32), (240, 320, 96), (300, 224, 96)):
        base = rng.random((h // 4 + 2, w // 4 + 2, 3))
        img = np.asarray(Image.fromarray((base * 255).astype(np.uint8)).resize((w, h), Image.BILINEAR))  # smooth content
        img = np.clip(img.astype(np.int32) + rng.integers(-20, 21, img.shape), 0, 255).astype(np.uint8)  # + texture
        cases.append({"image": torch.from_numpy(img.copy()), "n_px": n, "out": tf[n](Image.fromarray(img)).clone()})
    return {"cases": cases, "pillow": __import__("PIL").__version__, "torchvision": __import__("torchvision").__version__}


def main_preprocess():
    torch.save(preprocess_fixture(), os.path.join(OUT, "preprocess.pt"))
    print("preprocess done")


def preprocess_train_fixture():
    """The reference's `get_random_train_tfm()` (datasets/imagenet.py:8-23, the module loaded by file path) on seeded RGB
    images under `torch.manual_seed(seed)`: inputs (uint8), the seed, and of the reference's fp32 output [3, 224, 224]
    its bytes (the tensor mapped back through Normalize / ToTensor: exact) + the sha256 of the fp32 tensor itself."""
    import hashlib
    import importlib.util
    import numpy as np
    from PIL import Image
    spec = importlib.util.spec_from_file_location("_ref_imagenet", os.path.join(reference_shims.REFERENCE_ROOT, "datasets", "imagenet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tf = mod.get_random_train_tfm()  # the real reference function
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(3, 1, 1)
    rng = np.random.default_rng(9)
    cases = []
    for seed, (h, w) in enumerate(((120, 160), (224, 224), (150, 100), (64, 250), (97, 131))):
        base = rng.random((h // 4 + 2, w // 4 + 2, 3))
        img = np.asarray(Image.fromarray((base * 255).astype(np.uint8)).resize((w, h), Image.BILINEAR))
        img = np.clip(img.astype(np.int32) + rng.integers(-20, 21, img.shape), 0, 255).astype(np.uint8)
        torch.manual_seed(100 + seed)
        out = tf(Image.fromarray(img)).clone()
        u8 = torch.round((out * std + mean) * 255).to(torch.uint8)
        assert torch.equal(((u8.float() / 255) - mean) / std, out)
        cases.append({"image": torch.from_numpy(img.copy()), "seed": 100 + seed, "out_u8": u8,
                      "out_sha256": hashlib.sha256(out.contiguous().numpy().tobytes()).hexdigest()})
    return {"cases": cases, "pillow": __import__("PIL").__version__, "torchvision": __import__("torchvision").__version__}


def main_preprocess_train():
    torch.save(preprocess_train_fixture(), os.path.join(OUT, "preprocess_train.pt"))
    print("preprocess_train done")


def main_336():
    """ViT-L/14@336px (config C4: L = 577 tokens -> the attention kernel's multi-block path). fp32 reference only."""
    fx = tower_fixture("ViT-L/14@336px", B=2, P=1, with_fp16=False, with_blocks=False)
    torch.save(fx, os.path.join(OUT, "tower_ViT_L_14_336px.pt"))
    print("ViT-L/14@336px done")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "336":
        main_336()
    elif len(sys.argv) > 1 and sys.argv[1] == "rn":
        main_rn()
    elif len(sys.argv) > 1 and sys.argv[1] == "preprocess":
        main_preprocess()
    elif len(sys.argv) > 1 and sys.argv[1] == "preprocess_train":
        main_preprocess_train()
    else:
        main()
