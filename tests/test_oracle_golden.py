"""Pins oracle/protoclip_oracle.py (the CPU restatement) to the REAL reference: committed golden fixtures made
by tests/golden/make_golden.py, plus live comparisons when /root/reference is mounted (authoring container)."""
import hashlib

import pytest
import torch

from conftest import golden_images, load_golden, rel_err
from oracle import protoclip_oracle as O
from oracle import reference_shims
from proto_clip_b200 import synthetic

torch.set_grad_enabled(False)
FP32_TOL = 5e-6   # oracle(fp32) vs reference fp32: same math, different op grouping
FP16_TOL = 4e-3   # oracle(fp16 emulation) vs reference CPU-half kernels: rounding-order noise of fp16


@pytest.mark.parametrize("name", ["tiny", "small", "ViT_B_32", "ViT_B_16"])
def test_towers_match_reference_goldens(name):
    fx = load_golden(f"tower_{name}.pt")
    sd = synthetic.make_state_dict(fx["arch"], fx["seed"])
    images = golden_images(fx)
    assert rel_err(O.encode_image(sd, images, "fp32"), fx["image_features_fp32"]) < FP32_TOL
    assert rel_err(O.encode_text(sd, fx["tokens"], "fp32"), fx["text_features_fp32"]) < FP32_TOL
    if name in ("tiny", "small"):
        assert rel_err(O.encode_image(sd, images, "fp16"), fx["image_features_fp16"]) < FP16_TOL
        assert rel_err(O.encode_text(sd, fx["tokens"], "fp16"), fx["text_features_fp16"]) < FP16_TOL


@pytest.mark.parametrize("name", ["rn_tiny", "rn_small", "RN50"])
def test_resnet_towers_match_reference_goldens(name):
    """ModifiedResNet / Bottleneck / AttentionPool2d (clip/model.py:10-152) restatement vs the reference's outputs."""
    fx = load_golden(f"tower_{name}.pt")
    sd = synthetic.make_state_dict(fx["arch"], fx["seed"])
    images = golden_images(fx)
    assert rel_err(O.encode_image(sd, images, "fp32"), fx["image_features_fp32"]) < FP32_TOL
    if "image_features_fp16" in fx:
        assert rel_err(O.encode_image(sd, images, "fp16"), fx["image_features_fp16"]) < FP16_TOL


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_resblock_matches_reference_goldens(name):
    fx = load_golden(f"tower_{name}.pt")
    c = synthetic.arch_config(fx["arch"])
    sd = synthetic.make_state_dict(fx["arch"], fx["seed"])
    for tower, prefix, heads, causal in (("vis", "visual.transformer.resblocks.0.", c["vision_width"] // 64, False),
                                         ("txt", "transformer.resblocks.0.", c["transformer_heads"], True)):
        x = fx[f"{tower}_block0_in"].float().permute(1, 0, 2)  # reference [L,B,d] -> batch-first
        y32 = O.resblock(x, sd, prefix, heads, causal, "fp32").permute(1, 0, 2)
        y16 = O.resblock(x, sd, prefix, heads, causal, "fp16").permute(1, 0, 2)
        assert rel_err(y32, fx[f"{tower}_block0_fp32"]) < FP32_TOL
        assert rel_err(y16, fx[f"{tower}_block0_fp16"]) < FP16_TOL


@pytest.mark.parametrize("D", [64, 512, 768, 1024])
@pytest.mark.parametrize("kind", ["fc", "conv-2x", "conv-3x"])
def test_adapters_match_reference_goldens(kind, D):
    fx = load_golden("adapters.pt")
    sd = synthetic.make_adapter_state_dict(kind, D, seed=4)
    x = fx[f"x_{D}"]
    f = (lambda m: O.adapter_fc(sd, x, m)) if kind == "fc" else (lambda m: O.adapter_conv(sd, x, kind, m))
    assert rel_err(f("fp32"), fx[f"{kind}_{D}_fp32"]) < 2e-5
    assert rel_err(f("fp16"), fx[f"{kind}_{D}_fp16"]) < FP16_TOL


def test_head_matches_reference_goldens():
    fx = load_golden("head.pt")
    N, K = fx["N"], fx["K"]
    zi = O.build_prototypes(fx["V"], N, K, True, "fp32")
    zt = O.text_prototypes(fx["T"], "fp32")
    assert rel_err(zi, fx["z_img_fp32"]) < FP32_TOL and rel_err(zt, fx["z_txt_fp32"]) < FP32_TOL
    assert rel_err(O.build_prototypes(fx["V"], N, K, True, "fp16"), fx["z_img_fp16"]) < 2e-3
    assert rel_err(O.build_prototypes(fx["V"], N, K, False, "fp32"), fx["z_img_zeroshot_fp32"]) < FP32_TOL
    for (a, b) in ((0.5, 12.0), (0.2, 5.5), (1.0, 1.0), (0.0, 20.0)):
        p = O.P(fx["q"], fx["z_img_fp32"], fx["z_txt_fp32"], a, b)
        assert rel_err(p, fx[f"p_fp32_{a}_{b}"]) < 1e-5
        assert torch.equal(O.predict(p), fx[f"pred_fp32_{a}_{b}"])
        assert torch.allclose(p.sum(1), torch.ones(p.shape[0]), atol=1e-5)


@pytest.mark.parametrize("name", ["ckpt_imagenet_F_16.pt", "ckpt_fewsol_198_F_24.pt"])
def test_shipped_checkpoint_subsets(name):
    """Class subsets of pretrained_ckpt/{imagenet-F,fewsol-198-F}: reference head on its own memory bank."""
    fx = load_golden(name)
    K, kind = fx["K"], fx["kind"]
    N = fx["T"].shape[0]
    for mode in ("fp32", "fp16"):
        zi = O.build_prototypes(fx["V"], N, K, True, mode)
        zt = O.text_prototypes(fx["T"], mode)
        q = O.adapter_fc(fx["adapter"], fx["V"], mode) if kind == "fc" else O.adapter_conv(fx["adapter"], fx["V"], kind, mode)
        q = O.l2_normalize(q, mode)
        p = O.P(q, zi, zt, fx["alpha"], fx["beta"])
        assert rel_err(q[:32], fx[f"q_{mode}"]) < (2e-5 if mode == "fp32" else FP16_TOL)
        assert rel_err(p, fx[f"p_{mode}"]) < (1e-4 if mode == "fp32" else 2e-2)
        assert torch.equal(O.predict(p), fx[f"pred_{mode}"])
        labels = torch.arange(N).repeat_interleave(K)
        assert torch.equal(fx[f"pred_{mode}"], labels)  # self-accuracy 1.0 (SURVEY.md §4 KAT)


# ----------------------------------------------------------------------------- live reference (container only)
needs_ref = pytest.mark.skipif(not reference_shims.available(), reason="/root/reference not mounted")


@needs_ref
def test_live_reference_tower_tiny():
    ref = reference_shims.reference()
    sd = synthetic.make_state_dict("tiny", 3)
    model = ref.clip_model.build_model({k: v.clone() for k, v in sd.items()}).float()
    images = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(9))
    assert rel_err(O.encode_image(sd, images, "fp32"), model.encode_image(images)) < FP32_TOL


@needs_ref
@pytest.mark.parametrize("name,kind,alpha,beta,sha", [("imagenet-F", "conv-2x", 0.5, 12.0, "5f560bc903babdf4"),
                                                      ("fewsol-198-F", "fc", 0.2, 12.0, "47eedc14c9b203fe")])
def test_live_full_checkpoint_kat(name, kind, alpha, beta, sha):
    """full shipped checkpoints, self-accuracy 1.0 and the sha256 prefix of the int64 predictions."""
    import os
    d = os.path.join(reference_shims.REFERENCE_ROOT, "pretrained_ckpt", name)
    V = torch.load(os.path.join(d, "memory_bank_v.pt"), map_location="cpu", weights_only=False).data.half()
    T = torch.load(os.path.join(d, "memory_bank_t.pt"), map_location="cpu", weights_only=False).data.half()
    A = {k: v.half() for k, v in torch.load(os.path.join(d, "query_adapter.pt"), map_location="cpu", weights_only=False).items()}
    N, K = T.shape[0], 16
    zi, zt = O.build_prototypes(V, N, K, True, "fp32"), O.text_prototypes(T, "fp32")
    q = O.adapter_fc(A, V, "fp32") if kind == "fc" else O.adapter_conv(A, V, kind, "fp32")
    pred = O.predict(O.P(O.l2_normalize(q, "fp32"), zi, zt, alpha, beta))
    assert torch.equal(pred, torch.arange(N).repeat_interleave(K))
    assert hashlib.sha256(pred.numpy().tobytes()).hexdigest().startswith(sha)


@needs_ref
def test_live_tokenizer_kat():
    ref = reference_shims.reference()
    assert ref.clip.tokenize("a photo of a dog.")[0, :8].tolist() == [49406, 320, 1125, 539, 320, 1929, 269, 49407]
