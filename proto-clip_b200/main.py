"""Drop-in for the reference's main.py CLI on the inference path.

    python proto-clip_b200/main.py --config proto-clip_b200/configs/dtd.yml --dataset dtd --only_test ...

Same argparse flags (main.py:24-49), YAML keys and CLI->cfg overlay (main.py:52-71, including the reference's
truthiness quirk: `--alpha 0` / `--beta 0` are ignored), cache directory layout (./caches/<dataset>/...),
memory-bank / feature / hyper-parameter-search file formats, and printed accuracies. Differences:
  * `--only_test` and `--train_vis_memory_only` on the command line ARE honoured here (the reference parses
    them but never copies them into cfg) — a superset of the reference behaviour;
  * episodic training (main.py:216-381) is outside the inference hot path: with `only_test: False` the run stops
    after the training-free evaluation and says so. Testing a trained Proto-CLIP-F needs its `_v/_t/_a.pt` files
    exactly as in the reference (main.py:385-398);
  * every encoder / adapter / P call runs on libprotoclip_b200 (sm_100a);
  * multi-GPU: `torchrun --nproc-per-node N proto-clip_b200/main.py ...` (the reference is single-GPU, README.md:44).
    One process per GPU (dist.py): the memory-bank builders and pre_load_features shard their loader batches / prompts
    over the ranks and meet in one all-gather each (utils.py), rank 0 alone writes the cache files, the prototype
    memory is built on rank 0 and shipped in ONE broadcast, every rank scores its contiguous slice of the query
    features (grid-search hit counts are summed, predictions gathered in query order). Results are identical to the
    single-process run (tests/test_gpu_parity.py::test_cli_two_ranks_equal_one_rank).
"""
from __future__ import annotations

import argparse
import os
import random
import sys

import numpy as np
import torch
import yaml
from tqdm import tqdm

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))

from proto_clip_b200 import clip  # noqa: E402
from proto_clip_b200 import dist as pdist  # noqa: E402
from proto_clip_b200.model import Adapter, Adapter_FC  # noqa: E402
from proto_clip_b200.utils import (P, beautify, build_cache_model, build_prototypes, get_model_dir_root,  # noqa: E402
                                   get_seed, get_textual_memory_bank, load, pre_load_features, predict, save)


def get_arguments(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--logs", dest="logs_dir_path", help="log directory path", required=False)
    parser.add_argument("--config", dest="config", help="settings of Proto-CLIP in yaml format", required=True)
    parser.add_argument("--alpha", dest="alpha", help="alpha", type=float, required=False)
    parser.add_argument("--beta", dest="beta", help="beta", type=float, required=False)
    parser.add_argument("--adapter", dest="adapter", help="adapter to use: ['conv-3x', 'conv-2x', 'fc']", type=str,
                        required=False)
    parser.add_argument("--train_vis_memory_only", dest="train_vis_mem_only", help="train visual memory only",
                        action="store_true")
    parser.add_argument("--only_test", dest="only_test", help="flag to perform only testing", action="store_true")
    parser.add_argument("--shots", dest="shots", help="shots in few-shot setups", type=int, required=False)
    parser.add_argument("--losses", nargs="+", dest="losses", help="List of loss aliases: {'L1', 'L2', 'L3'}",
                        required=False)
    parser.add_argument("--backbone", dest="backbone",
                        help="backbones: [ViT-B/16, ViT-B/32, ViT-L/14], a state-dict path, or synthetic:<arch>",
                        type=str, required=False)
    parser.add_argument("--dataset", dest="dataset",
                        help="dataset alias: [ caltech101, dtd, eurosat, fgvc, food101, imagenet, oxford_flowers, "
                             "oxford_pets, stanford_cars, sun397, ucf101 ]", required=False)
    parser.add_argument("--gpu_augment", dest="gpu_augment", action="store_true",
                        help="(not in the reference) keep the decoded support images in HBM and run get_random_train_tfm "
                             "on the GPU while the visual memory bank is built")
    return parser.parse_args(argv)


def populate_cfg_using_args(cfg, args):
    """main.py:52-71 (truthiness tests kept: a 0 value never overrides the YAML), plus the two store_true flags."""
    if args.logs_dir_path:
        cfg["logs_dir_path"] = args.logs_dir_path
    if args.alpha:
        cfg["alpha"] = args.alpha
    if args.beta:
        cfg["beta"] = args.beta
    if args.adapter:
        cfg["adapter"] = args.adapter
    if args.shots:
        cfg["shots"] = args.shots
    if args.losses:
        cfg["losses"] = args.losses
    if args.backbone:
        cfg["backbone"] = args.backbone
    if args.dataset:
        cfg["dataset"] = args.dataset
    if args.only_test:
        cfg["only_test"] = True
    if args.train_vis_mem_only:
        cfg["train_vis_mem_only"] = True
    if getattr(args, "gpu_augment", False):
        cfg["gpu_augment"] = True
    return cfg


# main.py:74-102: per-dataset (search_scale, search_step) inherited from Tip-Adapter; the reference stores them in cfg
# and never reads them again on this path -- kept so that a cfg dumped by either program has the same keys
_SEARCH = {"caltech101": ([12, 5], [200, 20]), "dtd": ([13, 13], [200, 20]), "eurosat": ([12, 10], [200, 20]),
           "fgvc": ([30, 30], [200, 20]), "food101": ([10, 10], [200, 20]), "imagenet": ([7, 3], [200, 20]),
           "oxford_flowers": ([50, 50], [200, 20]), "oxford_pets": ([7, 3], [200, 20]),
           "stanford_cars": ([20, 10], [200, 20]), "sun397": ([12, 10], [200, 20]), "ucf101": ([7, 3], [200, 20]),
           "fewsol": ([13, 13], [200, 20])}


def search_scale_step(cfg):
    cfg["search_scale"], cfg["search_step"] = _SEARCH.get(cfg["dataset"], (None, None))
    return cfg


def make_adapter(cfg, ndim):
    """Adapter alias dispatch of main.py:118-121: any alias containing 'conv' -> Adapter, exactly 'fc' ->
    Adapter_FC; anything else is an error (the reference dies with an unbound `adapter`)."""
    if "conv" in cfg["adapter"]:
        return Adapter(ndim, c_type=cfg["adapter"], dtype=torch.half).cuda()
    if cfg["adapter"] == "fc":
        return Adapter_FC(ndim, dtype=torch.half).cuda()
    raise NameError(f"unknown adapter alias {cfg['adapter']!r}: expected 'conv-3x', 'conv-2x' or 'fc'")


# "main" = the reference's main.py, "qt" = its main.qt.py (the Q^T training variant; on the inference path it differs
# only in the un-rounded alpha grid, main.qt.py:109-111, and in the checkpoint directory name, main.qt.py:327).
VARIANT = "main"


def alpha_beta_lists():
    """main.py:142-146 / main.qt.py:109-113: 11 alphas x 29 betas."""
    alpha_list = np.arange(0, 1.1, 0.1)
    if VARIANT == "main":
        alpha_list = alpha_list.round(1)
    beta_list = np.concatenate((np.arange(0.1, 1, 0.1), np.arange(1, 21, 1.0)))
    return alpha_list, beta_list


def my_slice(n):
    """This rank's contiguous share [lo, hi) of n query rows (everything for a single process)."""
    rank, _, world = pdist.env_rank()
    return pdist.shard_bounds(n, rank, world) if pdist.active() else (0, n)


def grid_accuracy(features, labels, z_img_proto, z_text_proto):
    """[319, 3] float64 rows (alpha, beta, accuracy) — the format of zero_shot_hp_search_*.pkl (main.py:187-207)."""
    from proto_clip_b200 import _native as nat
    alpha_list, beta_list = alpha_beta_lists()
    # one fused pass instead of 319 P() calls on identical matmuls: counts[a, b] = #correct at (alpha_a, beta_b);
    # every rank scores its slice of the queries, the integer hit counts are summed over the ranks
    total = features.shape[0]
    lo, hi = my_slice(total)
    q = features[lo:hi].half().contiguous()
    zi, zt = z_img_proto.half().contiguous(), z_text_proto.half().contiguous()
    if hi > lo:
        counts = nat.proto_grid_search(q, zi, zt, zi.float().pow(2).sum(-1), zt.float().pow(2).sum(-1),
                                       labels[lo:hi].to(q.device).long().contiguous(), alpha_list, beta_list)
    else:
        counts = torch.zeros((len(alpha_list), len(beta_list)), dtype=torch.int32, device=features.device)
    counts = pdist.all_reduce_sum(counts)
    # `(pred == labels).float().mean()` of the reference: an fp32 division of an exactly representable count
    acc = (counts.float() / float(total)).cpu().numpy().astype(np.float64)
    rows = [[alpha, beta, acc[i, j]] for i, alpha in enumerate(alpha_list) for j, beta in enumerate(beta_list)]
    return np.array(rows)


def best_alpha_beta(val_acc):
    """Selection rule of plot_zero_shot_alpha_beta (utils.py:197-203): the first grid point with the best
    validation accuracy."""
    i = int(np.argmax(val_acc[:, 2]))
    return float(val_acc[i, 0]), float(val_acc[i, 1]), float(val_acc[i, 2])


def run_proto_clip(cfg, visual_memory_keys, visual_memory_values, val_features, val_labels, test_features,
                   test_labels, textual_memory_bank, clip_model, text_prompts):
    cfg = search_scale_step(cfg)  # main.py:111
    ndim, NxK = visual_memory_keys.shape
    K = cfg["shots"]
    N = NxK // K
    model_dir_root = get_model_dir_root(cfg)
    os.makedirs(model_dir_root, exist_ok=True)
    tag = f"{beautify(cfg['backbone'])}_K_{cfg['shots']}"
    paths = {s: os.path.join(model_dir_root, f"zero_shot_hp_search_{s}_{tag}.pkl") for s in ("val", "test", "train")}
    train_labels = torch.argmax(visual_memory_values, dim=1)
    adapter = make_adapter(cfg, ndim)

    # ---- training-free evaluation: (alpha, beta) grid with zero-shot prototypes (main.py:166-207)
    pdist.barrier()  # every rank sees the same cache state
    cached = all(os.path.exists(p) for p in paths.values())
    pdist.barrier()
    if cached:
        val_acc = load(paths["val"], "hp based on val set")
        test_acc = load(paths["test"], "hp based on test set")
        train_acc = load(paths["train"], "hp based on test set")
    else:
        with torch.no_grad():
            keys_t = visual_memory_keys.t().contiguous()
            from proto_clip_b200 import _native as nat
            z_img_proto, _ = nat.build_prototypes(keys_t.half(), N, K, per_shot_norm=False)  # main.py:173-176
            z_text_proto = nat.l2_normalize(textual_memory_bank.t().contiguous().half())
            z_img_proto, z_text_proto = pdist.broadcast_tensors([z_img_proto, z_text_proto])  # rank 0's prototypes
            train_features = nat.l2_normalize(keys_t.half())
            val_acc = grid_accuracy(nat.l2_normalize(val_features), val_labels, z_img_proto, z_text_proto)
            test_acc = grid_accuracy(nat.l2_normalize(test_features), test_labels, z_img_proto, z_text_proto)
            train_acc = grid_accuracy(train_features, train_labels, z_img_proto, z_text_proto)
        save(val_acc, paths["val"], "hp based on val set")
        save(test_acc, paths["test"], "hp based on test set")
        save(train_acc, paths["train"], "hp based on test set")
    za, zb, zacc = best_alpha_beta(val_acc)
    i = int(np.argmax(val_acc[:, 2]))
    log(f"**** Zero-shot Proto-CLIP: best val accuracy {zacc * 100:.2f}% at alpha={za}, beta={zb}; "
          f"test accuracy there {test_acc[i, 2] * 100:.2f}% ****")

    best_alpha, best_beta = cfg["alpha"], cfg["beta"]  # main.py:213-214
    if not cfg["only_test"]:
        log("Episodic training of the memory banks / adapter (reference main.py:216-381) is outside this "
              "inference build; stopping after the training-free evaluation. Re-run with only_test once "
              "trained `_v/_t/_a.pt` files exist.")
        return {"zero_shot_val_acc": zacc, "zero_shot_alpha": za, "zero_shot_beta": zb, "val_grid": val_acc,
                "test_grid": test_acc}

    # ---- testing a trained Proto-CLIP-F (main.py:383-455)
    with torch.no_grad():
        log("Testing...")
        ab_dir = "alpha-beta" if VARIANT == "main" else "best-alpha-beta"  # main.py:385 / main.qt.py:327
        model_dir = f"{model_dir_root}/{ab_dir}/{best_alpha}-{best_beta}"
        model_prefix = f"best_lr_{cfg['lr']}_aug_{cfg['augment_epoch']}_epochs_{cfg['train_epoch']}"
        pv, pt, pa = (os.path.join(model_dir, f"{model_prefix}_{s}.pt") for s in ("v", "t", "a"))
        try:
            embeddings_v = torch.load(pv, weights_only=False)
            embeddings_t = torch.load(pt, weights_only=False)
            adapter.load_state_dict(torch.load(pa, weights_only=False))
        except Exception:
            raise FileNotFoundError(f"File does not exist: {pv} and {pt}")
        z_img_proto, z_text_proto = build_prototypes(embeddings_v.data.cuda(), embeddings_t.data.cuda(), K)
        z_img_proto, z_text_proto = pdist.broadcast_tensors([z_img_proto, z_text_proto])  # rank 0's prototypes
        from proto_clip_b200 import _native as nat
        test_f = nat.l2_normalize(adapter(test_features))                      # main.py:407-409
        train_f = nat.l2_normalize(adapter(visual_memory_keys.t().contiguous()))
        val_f_adapt = adapter(val_features)                                      # un-renormalised, main.py:415 (quirk 6)
        val_acc = grid_accuracy(val_f_adapt, val_labels, z_img_proto, z_text_proto)
        test_acc_grid = grid_accuracy(test_f, test_labels, z_img_proto, z_text_proto)
        grid_accuracy(train_f, train_labels, z_img_proto, z_text_proto)
        lo, hi = my_slice(test_f.shape[0])

        def predictions(alpha, beta):
            """P(...).max(1)[1] (main.py:436-438) on this rank's query slice, gathered in query order."""
            mine = P(test_f[lo:hi], z_img_proto, z_text_proto, alpha, beta).max(1)[1] if hi > lo else \
                torch.empty(0, dtype=torch.int64, device=test_f.device)
            return pdist.all_gather_rows(mine, test_f.shape[0])

        test_pred = predictions(best_alpha, best_beta)
        test_acc = (test_pred == test_labels).float().mean()
        log("**** Fixed-alp-beta: Proto-CLIP's test accuracy: {:.2f}% ****\n".format(test_acc * 100))
        log("fixed_best_alpha", best_alpha, "fixed_best_beta", best_beta)
        ha, hb, _ = best_alpha_beta(val_acc)
        hp_pred = predictions(ha, hb)
        hp_acc = (hp_pred == test_labels).float().mean()
        log("**** HP-search: Proto-CLIP's test accuracy: {:.2f}% ****\n".format(hp_acc * 100))
        log("hp_search_best_alpha", ha, "hp_search_best_beta", hb)
    return {"test_acc": float(test_acc), "hp_test_acc": float(hp_acc), "hp_alpha": ha, "hp_beta": hb,
            "test_acc_grid": test_acc_grid, "val_grid": val_acc, "test_pred": test_pred.cpu(), "hp_pred": hp_pred.cpu()}


def log(*a, **k):
    """print on rank 0 only (every rank computes the same numbers)."""
    if pdist.is_main():
        print(*a, **k)


def seed_worker(worker_id):
    worker_seed = get_seed()
    np.random.seed(worker_seed)
    random.seed(worker_seed)


def make_support_loader(cfg, dataset, batch_size, generator):
    """The loader build_cache_model walks `augment_epoch` times (main.py:529-533: `get_random_train_tfm()`, shuffle=False).
    With `gpu_augment: True` in the YAML (or --gpu_augment) the few-shot split is decoded once, kept in HBM as uint8 and
    re-augmented on the GPU every pass (`datasets.GPUAugmentedLoader`: same transform, bit-identical to the reference's
    loader at num_workers=0 under the same seed) instead of going through 8 PIL worker processes."""
    from proto_clip_b200 import datasets
    if cfg.get("gpu_augment") and not str(cfg["dataset"]).startswith("synthetic"):
        return datasets.GPUAugmentedLoader(dataset.train_x, batch_size=batch_size)
    return datasets.build_data_loader(data_source=dataset.train_x, batch_size=batch_size,
                                      tfm=datasets.get_random_train_tfm(), is_train=True, shuffle=False,
                                      worker_init_fn=seed_worker, generator=generator)


def main(argv=None):
    args = get_arguments(argv)
    assert os.path.exists(args.config)
    cfg = yaml.load(open(args.config, "r"), Loader=yaml.Loader)
    if args.dataset is None:
        raise SystemExit("Please provide alias of dataset")
    cfg = populate_cfg_using_args(cfg, args)
    rank, local_rank, world = pdist.init()       # torchrun: one process per GPU; a plain launch is rank 0 of 1
    if torch.cuda.is_available():
        torch.cuda.set_device(pdist.device_for(local_rank))
    cache_dir = os.path.join("./caches", cfg["dataset"])
    os.makedirs(cache_dir, exist_ok=True)
    cfg["cache_dir"] = cache_dir
    log("\nRunning configs.")
    log(cfg, "\n")
    if world > 1:
        log(f"{world} ranks: loader batches / prompts / query features are sharded, one all-gather per bank, "
            f"one broadcast of the prototypes")

    clip_model, preprocess = clip.load(cfg["backbone"])
    clip_model.eval()

    seed = get_seed()
    random.seed(seed)
    np.random.seed(seed)
    g = torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    n_workers, train_bs, val_bs, test_bs = 8, 1024, 1024, 1024

    log("Preparing dataset.")
    from proto_clip_b200 import datasets
    datasets.configure(clip_model.visual.input_resolution)
    if cfg["dataset"] == "imagenet":
        dataset = datasets.ImageNet(cfg["root_path"], cfg["shots"], preprocess)
        train_loader_cache = torch.utils.data.DataLoader(dataset.train, batch_size=train_bs, num_workers=n_workers,
                                                         shuffle=False, worker_init_fn=seed_worker, generator=g)
        val_loader = torch.utils.data.DataLoader(dataset.test, batch_size=val_bs, num_workers=n_workers, shuffle=False)
        test_loader = torch.utils.data.DataLoader(dataset.test, batch_size=test_bs, num_workers=n_workers,
                                                  shuffle=False)
    else:
        dataset = datasets.build_dataset(cfg["dataset"], cfg["root_path"], cfg["shots"])
        train_loader_cache = make_support_loader(cfg, dataset, train_bs, g)
        val_loader = datasets.build_data_loader(data_source=dataset.val, batch_size=val_bs, is_train=False,
                                                tfm=preprocess, shuffle=False)
        test_loader = datasets.build_data_loader(data_source=dataset.test, batch_size=test_bs, is_train=False,
                                                 tfm=preprocess, shuffle=False)

    log("Constructing memory bank by few-shot visual and textual features.")
    visual_memory_keys, visual_memory_values = build_cache_model(cfg, clip_model, train_loader_cache)
    text_prompts, textual_memory_bank = get_textual_memory_bank(cfg, dataset.classnames, dataset.template, clip_model)
    log("Loading visual features and labels from val set.")
    val_features, val_labels = pre_load_features(cfg, "val", clip_model, val_loader)
    log("Loading visual features and labels from test set.")
    test_features, test_labels = pre_load_features(cfg, "test", clip_model, test_loader)
    out = run_proto_clip(cfg, visual_memory_keys, visual_memory_values, val_features, val_labels, test_features,
                         test_labels, textual_memory_bank, clip_model, text_prompts)
    pdist.barrier()
    return out


if __name__ == "__main__":
    main()
