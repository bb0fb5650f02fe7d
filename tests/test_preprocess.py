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


# ---------------------------------------------------------------- training augmentation (datasets/imagenet.py:8-23)
TRAIN_SIZES = [(480, 640), (224, 224), (300, 224), (100, 80), (333, 500), (64, 700), (700, 64), (17, 19), (224, 500)]


def replay_draw(h, w, seed):
    """The box / flip torchvision draws for an h x w image under `seed` (oracle restatement of get_params)."""
    torch.manual_seed(seed)
    box = PO.random_resized_crop_params(h, w)
    return box, PO.random_flip()


def test_train_oracle_matches_reference_fixture():
    import hashlib
    fx = load_golden("preprocess_train.pt")
    assert len(fx["cases"]) >= 5
    flips = 0
    for case in fx["cases"]:
        img = case["image"].numpy()
        box, flip = replay_draw(img.shape[0], img.shape[1], case["seed"])
        flips += flip
        assert np.array_equal(PO.train_transform_u8(img, *box, flip, 224).transpose(2, 0, 1), case["out_u8"].numpy())
        got = PO.train_transform(img, *box, flip, 224)
        assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == case["out_sha256"]
    assert 0 < flips < len(fx["cases"])  # both branches of the flip are pinned


@pytest.mark.parametrize("h,w", TRAIN_SIZES)
def test_train_oracle_matches_live_torchvision(h, w):
    Image = pytest.importorskip("PIL.Image")
    pytest.importorskip("torchvision.transforms")
    from proto_clip_b200 import datasets
    tf = datasets.get_random_train_tfm()  # the reference's Compose, datasets/imagenet.py:14-23
    img = random_image(h, w, h * 1000 + w + 3)
    for seed in range(3):
        torch.manual_seed(seed * 31 + h)
        ref = tf(Image.fromarray(img)).numpy()
        after_ref = torch.rand(1)
        box, flip = replay_draw(h, w, seed * 31 + h)
        assert torch.equal(torch.rand(1), after_ref), "the oracle consumed the generator differently"
        assert np.array_equal(PO.train_transform(img, *box, flip, 224), ref), (box, flip)


def test_gpu_train_transform_draws_like_torchvision():
    """The shell's host side (no GPU needed): same box as RandomResizedCrop.get_params for the same seed, including the
    central fallback for extreme aspect ratios, and the same generator state afterwards."""
    T = pytest.importorskip("torchvision.transforms")
    from proto_clip_b200 import datasets
    tf = datasets.GPUTrainTransform(224)
    for h, w in TRAIN_SIZES + [(20, 900), (900, 20)]:
        for seed in range(4):
            torch.manual_seed(seed)
            want = T.RandomResizedCrop.get_params(torch.empty(3, h, w), list(datasets.TRAIN_SCALE), list(datasets.TRAIN_RATIO))
            state = torch.get_rng_state()
            torch.manual_seed(seed)
            assert tf.get_params(h, w) == tuple(want)
            assert torch.equal(torch.get_rng_state(), state)
    with pytest.raises(ValueError, match="RGB"):
        tf(__import__("PIL.Image").Image.new("L", (30, 30)))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", TRAIN_SIZES)
def test_cuda_train_preprocess_bit_exact(h, w):
    from proto_clip_b200 import _native as nat
    img = random_image(h, w, h * 1000 + w + 11)
    dev_img = torch.from_numpy(img).cuda()
    for seed in range(4):
        box, flip = replay_draw(h, w, seed)
        ref = PO.train_transform(img, *box, flip, 224)
        got = nat.preprocess_train_image(dev_img, box, flip, 224)
        assert np.array_equal(got.cpu().numpy(), ref), (box, flip)
    # the whole image as the box, both flips, another output size, fp16
    for flip in (False, True):
        ref = PO.train_transform(img, 0, 0, h, w, flip, 96)
        got = nat.preprocess_train_image(dev_img, (0, 0, h, w), flip, 96)
        assert np.array_equal(got.cpu().numpy(), ref)
        got16 = nat.preprocess_train_image(dev_img, (0, 0, h, w), flip, 96, dtype=torch.float16)
        assert torch.equal(got16.cpu(), torch.from_numpy(ref).half())


@pytest.mark.gpu
def test_cuda_train_transform_fixture_seed_parity_and_errors():
    """GPUTrainTransform under the fixture's seed = the reference's get_random_train_tfm() output; a box that leaves
    the image is refused."""
    import hashlib
    from proto_clip_b200 import _native as nat
    from proto_clip_b200 import datasets
    fx = load_golden("preprocess_train.pt")
    tf = datasets.get_random_train_tfm(device="cuda")
    batch = torch.empty(len(fx["cases"]), 3, 224, 224, device="cuda")
    for i, case in enumerate(fx["cases"]):
        torch.manual_seed(case["seed"])
        got = tf(case["image"], out=batch[i])  # straight into a slot of an encoder batch
        assert got.data_ptr() == batch[i].data_ptr()
        assert hashlib.sha256(got.cpu().contiguous().numpy().tobytes()).hexdigest() == case["out_sha256"]
    img = torch.zeros(40, 50, 3, dtype=torch.uint8, device="cuda")
    for box in ((0, 0, 41, 50), (1, 0, 40, 50), (0, 10, 40, 41), (0, 0, 0, 5), (-1, 0, 5, 5)):
        with pytest.raises(nat.NativeError):
            nat.preprocess_train_image(img, box, False, 32)
    with pytest.raises(nat.NativeError):
        nat.preprocess_train_image(img.cpu(), (0, 0, 40, 50), False, 32)


@pytest.mark.gpu
def test_gpu_augmented_loader_reproduces_the_host_loader():
    """Two passes (= two augment epochs of build_cache_model) over decoded support images: every batch equals the
    reference's Compose applied image by image under the same seed (its loader at num_workers=0, shuffle=False)."""
    Image = pytest.importorskip("PIL.Image")
    from proto_clip_b200 import datasets
    sizes = [(90, 120), (224, 224), (150, 100), (64, 250), (97, 131), (300, 200), (240, 320)]
    source = [(random_image(h, w, 50 + i) if i % 2 else Image.fromarray(random_image(h, w, 50 + i)), i % 3)
              for i, (h, w) in enumerate(sizes)]
    host_tf = datasets.get_random_train_tfm()
    loader = datasets.GPUAugmentedLoader(source, batch_size=3)
    assert len(loader) == 3
    torch.manual_seed(1234)
    want = [[host_tf(im if hasattr(im, "convert") else Image.fromarray(im)) for im, _ in source] for _ in range(2)]
    torch.manual_seed(1234)
    for epoch in range(2):
        seen = 0
        for images, target in loader:
            assert images.is_cuda and images.shape[1:] == (3, 224, 224) and target.dtype == torch.int64
            for k in range(images.shape[0]):
                assert torch.equal(images[k].cpu(), want[epoch][seen + k]), (epoch, seen + k)
                assert int(target[k]) == source[seen + k][1]
            seen += images.shape[0]
        assert seen == len(source)
    # a rank's batch range draws for its own images only
    torch.manual_seed(7)
    part = list(loader.iter_range(1, 2))
    assert len(part) == 1 and part[0][0].shape[0] == 3
    torch.manual_seed(7)
    assert torch.equal(part[0][0][0].cpu(), host_tf(Image.fromarray(source[3][0])))


@pytest.mark.gpu
@pytest.mark.parametrize("shift", [1, 2, 3])
def test_cuda_preprocess_unaligned_buffers(shift):
    """The horizontal pass reads aligned 32-bit words around every tap run: image buffers that do not start on a word
    boundary (a view into a byte buffer) and runs that end with the buffer must give the same bytes."""
    from proto_clip_b200 import _native as nat
    for (h, w, n) in ((37, 53, 16), (100, 80, 224), (64, 64, 32), (9, 7, 8), (40, 2, 8)):
        img = random_image(h, w, 77 + h)
        flat = torch.zeros(h * w * 3 + shift, dtype=torch.uint8, device="cuda")
        view = flat[shift:].view(h, w, 3)
        view.copy_(torch.from_numpy(img))
        assert view.data_ptr() % 4 == (flat.data_ptr() + shift) % 4
        assert np.array_equal(nat.preprocess_image(view, n).cpu().numpy(), PO.clip_preprocess(img, n))
        for box, flip in (((0, 0, h, w), True), ((0, 1, h - 1, w - 1), False), ((1, 0, h - 1, w), False)):
            got = nat.preprocess_train_image(view, box, flip, 32)
            assert np.array_equal(got.cpu().numpy(), PO.train_transform(img, *box, flip, 32)), (h, w, box)
    # a batch: only the last image's last row ends with the buffer
    imgs = np.stack([random_image(40, 30, s) for s in range(3)])
    flat = torch.zeros(imgs.size + shift, dtype=torch.uint8, device="cuda")
    view = flat[shift:].view(3, 40, 30, 3)
    view.copy_(torch.from_numpy(imgs))
    got = nat.preprocess_image(view, 24)
    for i in range(3):
        assert np.array_equal(got[i].cpu().numpy(), PO.clip_preprocess(imgs[i], 24))


def test_augmented_loader_host_logic(tmp_path):
    """GPUAugmentedLoader without a GPU (a stand-in transform on the CPU): Datum-like items are decoded once from their
    `impath` like the reference's read_image (RGB), (image, label) pairs are taken as they are, batches keep the source
    order, `iter_range` serves a rank's batch range, and every pass calls the transform once per image in order."""
    Image = pytest.importorskip("PIL.Image")
    from types import SimpleNamespace
    from proto_clip_b200 import datasets

    class Recorder:
        size, dtype = 4, torch.float32

        def __init__(self):
            self.seen = []

        def __call__(self, rgb, out=None):
            assert rgb.dtype == torch.uint8 and rgb.dim() == 3 and rgb.shape[-1] == 3
            self.seen.append(rgb.clone())
            out.fill_(float(rgb[0, 0, 0]))
            return out

    arrays = [random_image(20 + i, 30 - i, i) for i in range(5)]
    items = []
    for i, a in enumerate(arrays):
        path = tmp_path / f"img{i}.png"
        Image.fromarray(a if i != 2 else a[:, :, 0]).save(path)  # image 2 is greyscale on disk -> convert("RGB")
        items.append(SimpleNamespace(impath=str(path), label=i % 2))
    rec = Recorder()
    loader = datasets.GPUAugmentedLoader(items, batch_size=2, tfm=rec, device="cpu")
    assert len(loader) == 3
    for epoch in range(2):
        batches = list(loader)
        assert [b[0].shape[0] for b in batches] == [2, 2, 1]
        assert torch.cat([b[1] for b in batches]).tolist() == [0, 1, 0, 1, 0]
    assert len(rec.seen) == 10  # decoded once, transformed once per image per pass, in source order
    for i, a in enumerate(arrays):
        want = a if i != 2 else np.repeat(a[:, :, :1], 3, axis=2)
        assert np.array_equal(rec.seen[i].numpy(), want) and np.array_equal(rec.seen[5 + i].numpy(), want)
    rec.seen.clear()
    part = list(loader.iter_range(1, 3))
    assert [b[0].shape[0] for b in part] == [2, 1] and len(rec.seen) == 3
    assert np.array_equal(rec.seen[0].numpy(), arrays[2][:, :, :1].repeat(3, axis=2))
    # (image, label) pairs with a tensor / array / PIL image; anything that is not RGB uint8 is refused
    mixed = datasets.GPUAugmentedLoader([(torch.from_numpy(arrays[0]), 3), (arrays[1], 4), (Image.fromarray(arrays[3]), 5)],
                                        batch_size=8, tfm=Recorder(), device="cpu")
    (images, target), = list(mixed)
    assert images.shape == (3, 3, 4, 4) and target.tolist() == [3, 4, 5]
    with pytest.raises(ValueError, match="RGB uint8"):
        list(datasets.GPUAugmentedLoader([(torch.zeros(4, 4, 3), 0)], tfm=Recorder(), device="cpu"))
