"""Host-side logic that needs no GPU: CLI / config overlay, adapter alias dispatch, state-dict surface,
tokenizer, synthetic generators, head-state packing, sharding and the gloo world_size-2 path."""
import os
import subprocess
import sys

import pytest
import torch
import yaml

from conftest import ROOT
from oracle import reference_shims
from proto_clip_b200 import dist as pdist
from proto_clip_b200 import pipeline, synthetic

PKG = os.path.join(ROOT, "proto-clip_b200")


def load_main():
    import importlib
    return importlib.import_module("proto_clip_b200.main")


def test_cli_overlay_and_quirks():
    main = load_main()
    cfg = yaml.load(open(os.path.join(PKG, "configs", "imagenet.yml")), Loader=yaml.Loader)
    assert cfg["alpha"] == 0.5 and cfg["beta"] == 12 and cfg["adapter"] == "conv-2x" and cfg["shots"] == 16
    a = main.get_arguments(["--config", "x.yml", "--dataset", "imagenet", "--alpha", "0", "--beta", "3.5",
                            "--adapter", "fc", "--backbone", "ViT-B/16", "--only_test", "--shots", "4"])
    cfg = main.populate_cfg_using_args(cfg, a)
    assert cfg["alpha"] == 0.5          # reference quirk: `if args.alpha:` ignores 0 (main.py:56)
    assert cfg["beta"] == 3.5 and cfg["adapter"] == "fc" and cfg["backbone"] == "ViT-B/16" and cfg["shots"] == 4
    assert cfg["only_test"] is True     # honoured here (superset of the reference)
    with pytest.raises(SystemExit):
        main.get_arguments([])          # --config is required


def test_every_reference_config_exists_with_same_keys():
    names = ["caltech101", "dtd", "eurosat", "fewsol", "fewsol_198", "fgvc", "food101", "imagenet", "master",
             "oxford_flowers", "oxford_pets", "stanford_cars", "sun397", "ucf101"]
    for n in names:
        cfg = yaml.load(open(os.path.join(PKG, "configs", f"{n}.yml")), Loader=yaml.Loader)
        assert {"root_path", "shots", "backbone", "lr", "augment_epoch", "train_epoch", "losses"} <= set(cfg)
        if n != "master":
            assert {"alpha", "beta", "adapter", "dataset", "only_test", "train_vis_mem_only"} <= set(cfg)
        if reference_shims.available():
            ref = yaml.load(open(os.path.join(reference_shims.REFERENCE_ROOT, "configs", f"{n}.yml")), Loader=yaml.Loader)
            assert ref == cfg, f"configs/{n}.yml differs from the reference"


def test_search_scale_step_table():
    """main.py:74-102: cfg gets the per-dataset (search_scale, search_step) pair, None for unknown datasets; the table
    equals the reference's (read out of its source with ast when the tree is mounted)."""
    import ast
    main = load_main()
    assert main.search_scale_step({"dataset": "imagenet"}) == {"dataset": "imagenet", "search_scale": [7, 3], "search_step": [200, 20]}
    assert main.search_scale_step({"dataset": "synthetic:8"})["search_scale"] is None
    ref_main = os.path.join(reference_shims.REFERENCE_ROOT, "main.py")
    if reference_shims.available() and os.path.isfile(ref_main):
        fn = next(n for n in ast.parse(open(ref_main).read()).body if isinstance(n, ast.FunctionDef) and n.name == "search_scale_step")
        table = next(ast.literal_eval(n.value) for n in ast.walk(fn) if isinstance(n, ast.Assign) and isinstance(n.value, ast.Dict))
        assert {k: (list(v[0]), list(v[1])) for k, v in table.items()} == {k: (v[0], v[1]) for k, v in main._SEARCH.items()}


def test_support_loader_choice(monkeypatch):
    """main.make_support_loader: the reference's DataLoader + get_random_train_tfm() by default, the HBM-resident GPU
    loader with `gpu_augment` (never for the synthetic alias, whose images are generated on the device already)."""
    from types import SimpleNamespace
    from proto_clip_b200 import datasets
    main = load_main()
    a = main.get_arguments(["--config", "x.yml", "--dataset", "dtd", "--gpu_augment"])
    cfg = main.populate_cfg_using_args({"dataset": "x"}, a)
    assert cfg["gpu_augment"] is True and cfg["dataset"] == "dtd"
    calls = {}
    monkeypatch.setattr(datasets, "build_data_loader", lambda **kw: calls.update(kw) or "host-loader")
    ds = SimpleNamespace(train_x=[SimpleNamespace(impath="a.jpg", label=0)])
    assert main.make_support_loader({"dataset": "dtd"}, ds, 64, None) == "host-loader"
    assert calls["shuffle"] is False and calls["is_train"] is True and calls["tfm"] is not None and calls["batch_size"] == 64
    loader = main.make_support_loader(cfg, ds, 64, None)
    assert isinstance(loader, datasets.GPUAugmentedLoader) and loader.batch_size == 64 and len(loader) == 1
    assert main.make_support_loader({"dataset": "synthetic:4", "gpu_augment": True}, ds, 64, None) == "host-loader"


def test_alpha_beta_grid_is_11_by_29():
    main = load_main()
    a, b = main.alpha_beta_lists()
    assert len(a) == 11 and len(b) == 29 and a[0] == 0 and a[-1] == 1.0 and abs(b[0] - 0.1) < 1e-9 and b[-1] == 20


def test_adapter_state_dict_keys_match_shipped_checkpoints():
    from proto_clip_b200.model import Adapter, Adapter_FC
    conv = Adapter(1024, "conv-2x", dtype=torch.half)
    fc = Adapter_FC(768, dtype=torch.half)
    assert list(conv.state_dict()) == ["conv1.weight", "bn1.weight", "bn1.bias", "conv2.weight", "bn2.weight",
                                       "bn2.bias", "conv3.weight", "bn3.weight", "bn3.bias"]
    assert list(fc.state_dict()) == ["fc.0.weight", "fc.1.weight", "fc.1.bias", "fc.2.weight", "fc.3.weight",
                                     "fc.3.bias"]
    assert conv.bn1.weight.shape == (16, 32, 32) and Adapter(512, "conv-3x").bn3.weight.shape == (1, 23, 23)
    if reference_shims.available():
        d = os.path.join(reference_shims.REFERENCE_ROOT, "pretrained_ckpt")
        conv.load_state_dict(torch.load(os.path.join(d, "imagenet-F", "query_adapter.pt"), map_location="cpu", weights_only=False))
        fc.load_state_dict(torch.load(os.path.join(d, "fewsol-198-F", "query_adapter.pt"), map_location="cpu", weights_only=False))
    main = load_main()
    with pytest.raises(NameError):
        main.make_adapter({"adapter": "mlp"}, 512)


def test_clip_surface_on_cpu():
    from proto_clip_b200 import _native as nat
    from proto_clip_b200 import clip
    assert clip.available_models() == ["RN50", "RN101", "RN50x4", "RN50x16", "ViT-B/32", "ViT-B/16", "ViT-L/14"]
    with pytest.raises(RuntimeError):
        clip.load("no-such-model")
    model, preprocess = clip.load("synthetic:tiny", device="cpu")
    assert model.dtype == torch.float16 and model.visual.input_resolution == 32
    sd = model.state_dict()
    ref = synthetic.make_state_dict("tiny", 0)
    assert list(sd) == [k for k in ref] and all(torch.equal(sd[k].float(), ref[k].float()) for k in ref)
    assert sd["visual.conv1.weight"].dtype == torch.float16 and sd["visual.ln_pre.weight"].dtype == torch.float32
    with pytest.raises(nat.NativeError):  # no CPU execution path
        model.encode_image(torch.zeros(1, 3, 32, 32))
    assert callable(preprocess)


def test_tokenizer_matches_reference():
    from proto_clip_b200.clip import bpe_tokenizer, tokenize
    try:
        bpe_tokenizer.find_vocab()
    except RuntimeError:
        pytest.skip("CLIP BPE vocabulary not available on this box")
    t = tokenize("a photo of a dog.")
    assert t.shape == (1, 77) and t[0, :8].tolist() == [49406, 320, 1125, 539, 320, 1929, 269, 49407]
    with pytest.raises(RuntimeError):
        tokenize("word " * 100)
    assert tokenize("word " * 100, truncate=True)[0, -1].item() == 49407
    if reference_shims.available():
        ref = reference_shims.reference()
        texts = ["itap of a great white shark.", "A bad photo of the Tench, Tinca tinca!!", "don't we're 123 hello-world",
                 "a crème brûlée &amp; naïve café", "  many   spaces\tand\nnewlines ", "a centered satellite photo of Annual Crop Land."]
        assert torch.equal(tokenize(texts), ref.clip.tokenize(texts))


def test_synthetic_is_deterministic_and_typed():
    a, b = synthetic.make_state_dict("tiny", 0), synthetic.make_state_dict("tiny", 0)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["visual.proj"], synthetic.make_state_dict("tiny", 1)["visual.proj"])
    assert a["visual.transformer.resblocks.0.attn.in_proj_weight"].dtype == torch.float16
    assert a["visual.transformer.resblocks.0.ln_1.weight"].dtype == torch.float32
    assert abs(synthetic.vit_flops_per_image("ViT-B/16") / 35.13e9 - 1) < 0.01
    assert abs(synthetic.vit_flops_per_image("ViT-L/14") / 162.03e9 - 1) < 0.01


def test_head_state_pack_roundtrip():
    N, K, D = 10, 4, 64
    g = torch.Generator().manual_seed(0)
    for kind in ("fc", "conv-3x"):
        head = pipeline.HeadState(torch.randn(N, D, generator=g).half(), torch.randn(N, D, generator=g).half(),
                                  torch.rand(N, generator=g), torch.rand(N, generator=g), kind,
                                  synthetic.make_adapter_state_dict(kind, D), 0.5, 12.0)
        flat = head.pack()
        assert flat.dtype == torch.float16 and flat.numel() == pipeline.HeadState.packed_numel(N, D, kind)
        back = pipeline.HeadState.unpack(flat, N, D, kind, 0.5, 12.0)
        assert torch.equal(back.z_img, head.z_img) and torch.equal(back.zt_n2, head.zt_n2)
        assert all(torch.equal(back.adapter[k], head.adapter[k]) for k in back.adapter)


def test_shard_bounds_cover_queries_in_order():
    for Q in (0, 1, 7, 1000, 50000):
        for R in (1, 2, 3, 8):
            spans = [pdist.shard_bounds(Q, r, R) for r in range(R)]
            assert spans[0][0] == 0 and spans[-1][1] == Q
            assert all(spans[i][1] == spans[i + 1][0] for i in range(R - 1))
            assert all(0 <= hi - lo <= (Q + R - 1) // R for lo, hi in spans)


WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["REPO"])
from proto_clip_b200 import dist as pdist, pipeline, synthetic
rank, local, world = pdist.init("gloo")
N, K, D, Q = 6, 2, 64, 11
numel = pipeline.HeadState.packed_numel(N, D, "fc")
flat = None
if rank == 0:
    g = torch.Generator().manual_seed(0)
    head = pipeline.HeadState(torch.randn(N, D, generator=g).half(), torch.randn(N, D, generator=g).half(),
                              torch.rand(N, generator=g), torch.rand(N, generator=g), "fc",
                              synthetic.make_adapter_state_dict("fc", D), 0.5, 12.0)
    flat = head.pack()
flat = pdist.broadcast_flat(flat, numel, torch.float16, torch.device("cpu"))      # the single collective
head = pipeline.HeadState.unpack(flat, N, D, "fc", 0.5, 12.0)
lo, hi = pdist.shard_bounds(Q, rank, world)
q = torch.randn(Q, D, generator=torch.Generator().manual_seed(1))
local_pred = (q[lo:hi] @ head.z_img.float().t()).argmax(1)                        # stand-in for the GPU classify
allp = pdist.gather_predictions(local_pred, Q)
# sharded memory-bank build (SURVEY f2): every rank produces its slice of the rows, one all-gather restores the order
rows = torch.arange(37 * 5, dtype=torch.float32).view(37, 5)
a, b = pdist.shard_bounds(37, rank, world)
full = pdist.all_gather_rows(rows[a:b].clone(), 37)
assert torch.equal(full, rows), full
t = pdist.max_over_ranks(float(rank + 1), torch.device("cpu"))
# the CLI's collectives (main.py / utils.py under torchrun): contiguous BATCH ranges of a loader gathered in order
# (ranks may hold no batch at all), integer hit counts summed, the prototypes in one broadcast
loader = [(torch.full((n, 3), float(i)), torch.full((n,), i, dtype=torch.int64)) for i, n in enumerate([4, 4, 4, 2])]
mine = list(pdist.sharded_batches(loader))
feats = torch.cat([b[0] for b in mine]) if mine else torch.empty((0, 3))
labs = torch.cat([b[1] for b in mine]) if mine else torch.empty(0, dtype=torch.int64)
assert torch.equal(pdist.all_gather_varlen(feats), torch.cat([b[0] for b in loader]))
assert torch.equal(pdist.all_gather_varlen(labs), torch.cat([b[1] for b in loader]))
one = [(torch.ones(5, 2), torch.zeros(5, dtype=torch.int64))]              # fewer batches than ranks
got = list(pdist.sharded_batches(one))
assert len(got) == (1 if rank == 0 else 0)
assert pdist.all_gather_varlen(got[0][0] if got else torch.empty((0, 2))).shape == (5, 2)
counts = pdist.all_reduce_sum(torch.full((11, 29), rank + 1, dtype=torch.int32))
assert int(counts[0, 0]) == world * (world + 1) // 2
za, zb = torch.full((4, 8), float(rank)).half(), torch.full((4, 8), float(10 + rank)).half()
za, zb = pdist.broadcast_tensors([za, zb])
assert float(za[0, 0]) == 0.0 and float(zb[3, 7]) == 10.0 and za.shape == (4, 8)
assert pdist.active() and pdist.is_main() == (rank == 0)
pdist.barrier()
if rank == 0:
    ref = (q @ head.z_img.float().t()).argmax(1)
    assert torch.equal(allp, ref), (allp, ref)
    assert t == float(world)
    print("OK", world)
'''


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_broadcast_shard_gather(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=ROOT, MASTER_ADDR="127.0.0.1")
    port = 29600 + world + (os.getpid() % 200)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert f"OK {world}" in r.stdout


def test_datasets_front_door_and_qt_variant(monkeypatch):
    """datasets alias dispatch (synthetic built in, reference aliases delegated or refused loudly) and the main.qt.py
    variant switches (un-rounded alpha grid, best-alpha-beta checkpoint directory)."""
    from proto_clip_b200 import datasets
    ds = datasets.build_dataset("synthetic:6:40", "DATA", 3)
    assert len(ds.classnames) == 6 and len(ds.train_x) == 18 and len(ds.test) == 40 and len(ds.val) == 40
    assert ds.train_x.labels.tolist() == [c for c in range(6) for _ in range(3)]
    assert len(datasets.build_data_loader(data_source=ds.test, batch_size=16, is_train=False)) == 3
    monkeypatch.delenv("PROTOCLIP_REFERENCE_ROOT", raising=False)
    monkeypatch.setattr(datasets, "_ref_pkg", None)
    with pytest.raises(RuntimeError, match="PROTOCLIP_REFERENCE_ROOT"):
        datasets.build_dataset("dtd", "DATA", 1)
    main = load_main()
    import numpy as np
    a_main, _ = main.alpha_beta_lists()
    monkeypatch.setattr(main, "VARIANT", "qt")
    a_qt, b_qt = main.alpha_beta_lists()
    assert len(a_qt) == 11 and len(b_qt) == 29 and np.allclose(a_qt, a_main) and a_qt[3] != a_main[3]  # 0.30000000000000004
    assert os.path.isfile(os.path.join(PKG, "main.qt.py"))


def test_resnet_state_dict_surface():
    """build_model on a ModifiedResNet state dict (clip/model.py:406-414): architecture inferred from tensor shapes,
    convert_weights dtypes (conv / attention-pool Linear fp16, BatchNorm + pos-emb fp32, counters int64), and the
    CPU copy refuses to run (no fallback)."""
    from proto_clip_b200 import _native as nat
    from proto_clip_b200.clip import model as M
    sd = synthetic.make_state_dict("rn_small", 0)
    m = M.build_model({k: v.float() if v.is_floating_point() else v for k, v in sd.items()})  # an fp32 checkpoint
    assert m.visual.input_resolution == 96 and m.visual.layers == (2, 1, 2, 1) and m.visual.output_dim == 128
    out = m.state_dict()
    assert list(out.keys()) == list(sd.keys())
    for k, v in out.items():
        if k.endswith("num_batches_tracked"):
            assert v.dtype == torch.int64
        elif k.startswith("visual.") and (".conv" in k or "downsample.0" in k or "_proj." in k):
            assert v.dtype == torch.float16, k
        elif k.startswith("visual."):
            assert v.dtype == torch.float32, k
    with pytest.raises(nat.NativeError):
        m.encode_image(torch.zeros(1, 3, 96, 96))
    # the flop model used by tools/gpu_probe.py matches SURVEY.md's table (RN50x16: ~149.4 GFLOP / image)
    assert abs(synthetic.rn_flops_per_image("RN50x16") / 1e9 - 149.4) < 0.1


def test_abi_argument_errors_without_gpu():
    """Entry points validate their arguments before touching the device: bad calls return a negative PC_ERR_* code and
    set pc_last_error(), with or without a GPU."""
    import ctypes
    from proto_clip_b200 import _native as nat
    lib = nat.load_library()
    rc = lib.pc_linear_shift_relu_forward(None, 8, None, 8, None, None, 0, None, 8, 4, 8, 8, nat.EPI_BIAS_QUICKGELU, 1, None)
    assert rc == -1 and b"epilogue" in lib.pc_last_error()
    assert lib.pc_rn_bind_weights(None, None) < 0 and len(lib.pc_last_error()) > 0
    assert lib.pc_encode_image_workspace_bytes(None, 0) == 0
    w = nat.AdapterConvWeights()
    assert lib.pc_adapter_conv_forward(ctypes.byref(w), 5, None, None, 1, 512, None) < 0
