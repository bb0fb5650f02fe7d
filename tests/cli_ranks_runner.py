"""Worker of tests/test_gpu_parity.py::test_cli_two_ranks_equal_one_rank: the drop-in main.py CLI (training-free grid,
then --only_test on a trained-model file set) on the synthetic dataset, as ONE process or as a torchrun job; rank 0
saves what the run produced so that the test can compare world sizes.   python cli_ranks_runner.py <workdir>"""
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proto_clip_b200 import dist as pdist  # noqa: E402
from proto_clip_b200 import main as M  # noqa: E402
from proto_clip_b200 import synthetic, utils  # noqa: E402

work = sys.argv[1]
os.makedirs(work, exist_ok=True)
os.chdir(work)
backbone, adapter, N, K, Q = "synthetic:small", "fc", 6, 2, 44       # 44 queries, batch 16: a ragged last batch
c = synthetic.arch_config("small")
cfg = {"root_path": "DATA", "shots": K, "backbone": backbone, "dataset": f"synthetic:{N}:{Q}", "only_test": False,
       "lr": 0.0001, "augment_epoch": 2, "train_epoch": 1, "alpha": 0.5, "beta": 12, "adapter": adapter,
       "train_vis_mem_only": False, "losses": ["L1"]}
rank, _, world = pdist.env_rank()
if rank == 0:
    with open("cfg.yml", "w") as f:
        yaml.safe_dump(cfg, f)
tok = torch.zeros(N, c["context_length"], dtype=torch.int64)          # the small tower's vocabulary is synthetic
gen = torch.Generator().manual_seed(5)
for i in range(N):
    n = int(torch.randint(3, 9, (1,), generator=gen))
    tok[i, 0], tok[i, 1 + n] = c["vocab_size"] - 2, c["vocab_size"] - 1
    tok[i, 1:1 + n] = torch.randint(1, c["vocab_size"] - 2, (n,), generator=gen)
utils.clip.tokenize = lambda prompts: tok
# small loader batches so that every rank gets several (main.py hard-codes 1024 like the reference)
_bdl = M.__dict__.get("datasets")
from proto_clip_b200 import datasets  # noqa: E402
_orig = datasets.build_data_loader
datasets.build_data_loader = lambda **kw: _orig(**{**kw, "batch_size": 16 if kw.get("is_train") is False else 5})
pdist.init()
pdist.barrier()
argv = ["--config", "cfg.yml", "--dataset", cfg["dataset"]]
out = M.main(argv)
root = utils.get_model_dir_root({**cfg, "cache_dir": os.path.join("./caches", cfg["dataset"])})
if rank == 0:
    keys = torch.load(f"{root}/aug/visual_mb_keys_aug_2_{K}_shots.pt")
    D = c["embed_dim"]
    asd = synthetic.make_adapter_state_dict(adapter, D, seed=4, out_gain=synthetic.trained_like_gain(D))
    T = synthetic.aligned_text_memory(keys.t().contiguous(), N, K, seed=6)
    mdir = f"{root}/alpha-beta/0.5-12"
    os.makedirs(mdir, exist_ok=True)
    prefix = f"{mdir}/best_lr_0.0001_aug_2_epochs_1"
    torch.save(torch.nn.Parameter(keys.t().contiguous().clone()), prefix + "_v.pt")
    torch.save(torch.nn.Parameter(T.cuda()), prefix + "_t.pt")
    torch.save({k: v.cuda() for k, v in asd.items()}, prefix + "_a.pt")
pdist.barrier()
res = M.main(argv + ["--only_test"])
# the tensor-in / tensor-out sharded builder bench.py's prelude uses (pipeline.build_memory_sharded), both input forms
from proto_clip_b200 import _native as nat  # noqa: E402
from proto_clip_b200 import pipeline  # noqa: E402
dev = torch.device("cuda", torch.cuda.current_device())
ctx = nat.Context(dev)
sd_small = synthetic.make_state_dict("small", 0)
ctx.bind_visual(sd_small)
ctx.bind_text(sd_small)
bases = synthetic.class_bases(N, c["image_resolution"], seed=1, device=dev)
sup_labels = torch.arange(N * 3, device=dev) // 3
sup = synthetic.class_structured_images(bases, sup_labels, seed=2)
V_t, T_t = pipeline.build_memory_sharded(ctx, sup, tok, num_support=None, chunk=5)
V_c, _ = pipeline.build_memory_sharded(ctx, lambda lo, hi: sup[lo:hi], None, num_support=N * 3, chunk=4)
assert torch.equal(V_t, V_c)
if rank == 0:
    text_mb = utils.load(f"{root}/text_mb_{utils.beautify(backbone)}_K_{K}.pkl", "text memory")
    torch.save({"world": world, "zero_val_grid": out["val_grid"], "zero_test_grid": out["test_grid"],
                "keys": torch.load(f"{root}/aug/visual_mb_keys_aug_2_{K}_shots.pt").cpu(),
                "values": torch.load(f"{root}/aug/visual_mb_values_aug_2_{K}_shots.pt").cpu(),
                "text_mb": text_mb.cpu(),
                "test_features": torch.load(f"{root}/test_features.pt").cpu(),
                "test_labels": torch.load(f"{root}/test_labels.pt").cpu(),
                "val_grid": res["val_grid"], "test_acc_grid": res["test_acc_grid"], "test_pred": res["test_pred"],
                "hp_pred": res["hp_pred"], "test_acc": res["test_acc"], "sharded_V": V_t.cpu(), "sharded_T": T_t.cpu()},
               "result.pt")
    print("RUNNER OK", world)
pdist.barrier()
