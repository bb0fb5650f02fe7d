"""Image preprocessing (reference clip/clip.py:77-84 `_transform`): the numpy oracle against the reference's own
transform (committed fixture + live PIL / torchvision when importable), and the CUDA path against the oracle —
bit-exact, this is byte / integer work followed by three correctly rounded fp32 operations."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import preprocess_oracle as PO

SIZES = [(37, 53, 16), (53, 37, 16), (224, 224, 224), (224, 300, 224), (100, 80, 224), (81, 100, 96), (500, 333, 336),
         (480, 640, 224), (17, 400, 32), (225, 224, 224), (64, 64, 32), (31, 31, 64), (1, 9, 8), (640, 481, 224)]


def random_image(h, w, seed):
    rng = np.random.default_rng(seed)
    coarse = rng.random((h // 5 + 2, w // 5 + 2, 3))
    img = np.kron(coarse, np.ones((5, 5, 1)))[:h, :w] * 255  # blocky content with edges + noise
    return np.clip(img + rng.integers(-30, 31, (h, w, 3)), 0, 255).astype(np.uint8)


def test_oracle_matches_reference_fixture():
    fx = load_golden("preprocess.pt")
    assert len(fx["cases"]) >= 7
    for case in fx["cases"]:
        got = PO.clip_preprocess(case["image"].numpy(), case["n_px"])
        assert np.array_equal(got, case["out"].numpy()), (tuple(case["image"].shape), case["n_px"])


@pytest.mark.parametrize("h,w,n", SIZES)
def test_oracle_matches_live_pil_and_torchvision(h, w, n):
    Image = pytest.importorskip("PIL.Image")
    T = pytest.importorskip("torchvision.transforms")
    img = random_image(h, w, h * 1000 + w)
    tf = T.Compose([T.Resize(n, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(n), lambda im: im.convert("RGB"),
                    T.ToTensor(), T.Normalize(PO.CLIP_MEAN, PO.CLIP_STD)])
    ref = tf(Image.fromarray(img)).numpy()
    assert np.array_equal(PO.clip_preprocess(img, n), ref)
    nh, nw = PO.resized_size(h, w, n)
    pil = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BICUBIC))
    assert np.array_equal(PO.resize_bicubic_u8(img, nw, nh), pil)


def test_gpu_transform_refuses_modes_it_cannot_reproduce():
    """Palette / alpha images are resized differently by Pillow (NEAREST, premultiplied alpha) before convert("RGB"):
    the device transform says so instead of silently diverging from the reference."""
    Image = pytest.importorskip("PIL.Image")
    from proto_clip_b200.clip.clip import GPUTransform
    tf = GPUTransform(32)
    for mode in ("P", "RGBA", "1"):
        with pytest.raises(ValueError, match="host `preprocess`"):
            tf(Image.new(mode, (40, 50)))


def test_grayscale_resize_commutes_with_rgb_conversion():
    """The reference converts to RGB after the resize; for mode L the GPU path converts first — same bytes."""
    Image = pytest.importorskip("PIL.Image")
    g = random_image(70, 45, 1)[:, :, 0].copy()
    after = np.asarray(Image.fromarray(g).resize((32, 49), Image.BICUBIC).convert("RGB"))
    before = PO.resize_bicubic_u8(np.repeat(g[:, :, None], 3, axis=2), 32, 49)
    assert np.array_equal(after, before)


def test_size_arithmetic():
    assert PO.resized_size(480, 640, 224) == (224, 298) and PO.resized_size(640, 480, 224) == (298, 224)
    assert PO.crop_offsets(224, 298, 224) == (0, 37) and PO.crop_offsets(225, 224, 224) == (0, 0)  # round half to even
    assert PO.crop_offsets(227, 224, 224) == (2, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,n", SIZES)
def test_cuda_preprocess_bit_exact(h, w, n):
    from proto_clip_b200 import _native as nat
    img = random_image(h, w, h * 1000 + w + 7)
    ref = PO.clip_preprocess(img, n)
    dev_img = torch.from_numpy(img).cuda()
    got = nat.preprocess_image(dev_img, n)
    assert got.shape == (3, n, n) and got.dtype == torch.float32
    assert np.array_equal(got.cpu().numpy(), ref)
    got16 = nat.preprocess_image(dev_img, n, dtype=torch.float16)
    assert torch.equal(got16.cpu(), torch.from_numpy(ref).half())


@pytest.mark.gpu
def test_cuda_preprocess_fixture_and_errors():
    from proto_clip_b200 import _native as nat
    fx = load_golden("preprocess.pt")
    batch = torch.empty(len(fx["cases"]), 3, 96, 96, device="cuda")
    for case in fx["cases"]:
        got = nat.preprocess_image(case["image"].cuda(), case["n_px"])
        assert torch.equal(got.cpu(), case["out"])
    # writing straight into a slot of an encoder batch
    big = [c for c in fx["cases"] if c["n_px"] == 96]
    for i, case in enumerate(big):
        nat.preprocess_image(case["image"].cuda(), 96, out=batch[i])
        assert torch.equal(batch[i].cpu(), case["out"])
    # a batch of same-size images in two launches = the images one by one
    imgs = torch.stack([torch.from_numpy(random_image(120, 90, s)) for s in range(9)]).cuda()
    one_by_one = torch.stack([nat.preprocess_image(im, 64) for im in imgs])
    assert torch.equal(nat.preprocess_image(imgs, 64), one_by_one)
    assert np.array_equal(one_by_one[4].cpu().numpy(), PO.clip_preprocess(imgs[4].cpu().numpy(), 64))
    with pytest.raises(nat.NativeError):
        nat.preprocess_image(torch.zeros(8, 8, 3, dtype=torch.uint8), 8)        # CPU tensor: no fallback
    with pytest.raises(ValueError):
        nat.preprocess_image(torch.zeros(8, 8, 4, dtype=torch.uint8, device="cuda"), 8)
