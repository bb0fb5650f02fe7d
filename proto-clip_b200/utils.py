"""Drop-in for the hot-path half of the reference's utils.py: same function names, arguments, return layouts
and cache-file formats; the arithmetic runs on libprotoclip_b200.

  P                         utils.py:225-244     build_cache_model        utils.py:284-332
  clip_classifier           utils.py:256-273     get_textual_memory_bank  utils.py:54-69
  pre_load_features         utils.py:335-361     cls_acc / save / load / beautify / get_model_dir_root

Plotting, TensorBoard and the training losses (utils.py:72-222) are outside the inference hot path and are
not provided here.
"""
from __future__ import annotations

import os
import pickle

import torch
import torch.nn.functional as F
from tqdm import tqdm

try:
    from . import _native as nat
    from . import clip
    from . import dist as pdist
except ImportError:  # pragma: no cover - top-level import when main.py runs as a script
    from proto_clip_b200 import _native as nat
    from proto_clip_b200 import clip
    from proto_clip_b200 import dist as pdist

# Multi-GPU (torchrun --nproc-per-node N main.py ...; dist.py): the three builders below shard their loader batches /
# prompts over the ranks, meet in ONE all-gather each (per-row arithmetic does not depend on the shard, so the result
# is bit-identical to the single-process run), and only rank 0 writes the cache files, between two barriers.


def get_seed():
    return 1


def _here():
    """map_location of the cache files: this rank's device (they were written from rank 0's)."""
    return f"cuda:{torch.cuda.current_device()}"


def dir_exists(path):
    return os.path.exists(path)


def save(obj, filepath, msg):
    if not pdist.is_main():  # one writer per cache file
        return
    print(f"Saving {msg} to {filepath}")
    with open(filepath, "wb") as handle:
        pickle.dump(obj, handle, protocol=pickle.HIGHEST_PROTOCOL)


def load(filepath, msg):
    print(f"Loading {msg} from {filepath}")
    with open(filepath, "rb") as handle:
        return pickle.load(handle)


def beautify(string):
    return string.strip().replace("/", "_").replace("-", "_")


def get_model_dir_root(cfg):
    return f"{cfg['cache_dir']}/models/{beautify(cfg['backbone'])}/K-{cfg['shots']}"


_n2_cache = {}


def _norm2(z):
    """||z_n||^2 [N] fp32 of a prototype matrix, remembered per (storage, shape, version): main.py calls P() / predict()
    hundreds of times on the same prototypes (main.py:419-448)."""
    key = (z.data_ptr(), tuple(z.shape), z._version, str(z.device))
    hit = _n2_cache.get(key)
    if hit is None:
        if len(_n2_cache) > 16:
            _n2_cache.clear()
        hit = z.float().pow(2).sum(-1)
        _n2_cache[key] = hit
    return hit


def P(zq_imgs_flat, z_img_proto, z_text_proto, alpha, beta):
    """p = alpha * softmax(-beta * ||q - c_img||^2) + (1 - alpha) * softmax(-beta * ||q - c_txt||^2), fp32 [Q, N].

    Inputs of any float dtype on a CUDA device (the reference casts to fp32 itself, utils.py:231-233); they are
    taken as fp16 values, which is what every caller passes (features / prototypes are fp16 tensors)."""
    q = zq_imgs_flat.half().contiguous()
    zi = z_img_proto.half().contiguous()
    zt = z_text_proto.half().contiguous()
    p, _, _ = nat.proto_classify(q, zi, zt, _norm2(zi), _norm2(zt), float(alpha), float(beta), want_p=True,
                                 want_argmax=False)
    return p


def predict(zq_imgs_flat, z_img_proto, z_text_proto, alpha, beta):
    """P(...).max(1)[1] (main.py:438) without materialising the [Q, N] probability matrix."""
    q = zq_imgs_flat.half().contiguous()
    zi = z_img_proto.half().contiguous()
    zt = z_text_proto.half().contiguous()
    _, am, _ = nat.proto_classify(q, zi, zt, _norm2(zi), _norm2(zt), float(alpha), float(beta), want_p=False,
                                  want_argmax=True)
    return am


def cls_acc(output, target, topk=1):
    pred = output.topk(topk, 1, True, True)[1].t()
    correct = pred.eq(target.view(1, -1).expand_as(pred))
    acc = float(correct[:topk].reshape(-1).float().sum(0, keepdim=True).cpu().numpy())
    return 100 * acc / target.shape[0]


def clip_classifier(classnames, template, clip_model):
    """Textual memory bank [D, N] fp16: per class, mean over templates of the L2-normalised text features,
    renormalised (utils.py:256-273). All N*T prompts go through ONE encode_text call instead of the reference's
    per-class loop — the per-prompt arithmetic is identical."""
    with torch.no_grad():
        prompts = []
        for classname in classnames:
            classname = classname.replace("_", " ")
            prompts += [t.format(classname) for t in template]
        texts = clip.tokenize(prompts)
        rank, _, world = pdist.env_rank()
        lo, hi = pdist.shard_bounds(texts.shape[0], rank, world) if pdist.active() else (0, texts.shape[0])
        emb = nat.l2_normalize(clip_model.encode_text(texts[lo:hi].cuda())) if hi > lo else \
            torch.empty((0, clip_model.text_projection.shape[1]), dtype=torch.float16, device="cuda")  # utils.py:266-267
        emb = pdist.all_gather_rows(emb, texts.shape[0])
        emb = emb.view(len(classnames), len(template), -1)
        mean = emb.float().mean(dim=1).half()                                    # fp16 mean, fp32 accumulation
        clip_weights = nat.l2_normalize(mean).t().contiguous()                   # utils.py:268-271
    return classnames, clip_weights


def get_textual_memory_bank(cfg, classnames, template, clip_model):
    msg = "text_memory_bank"
    model_dir_root = get_model_dir_root(cfg)
    os.makedirs(model_dir_root, exist_ok=True)
    path = os.path.join(model_dir_root, f"text_mb_{beautify(cfg['backbone'])}_K_{cfg['shots']}.pkl")
    pdist.barrier()  # every rank sees the same cache state
    if dir_exists(path):
        return classnames, load(path, msg)
    pdist.barrier()
    text_prompts, textual_memory_bank = clip_classifier(classnames, template, clip_model)
    save(textual_memory_bank, path, msg)
    pdist.barrier()
    return text_prompts, textual_memory_bank


def build_cache_model(cfg, clip_model, train_loader_cache):
    """Visual memory bank: keys [D, N*K] fp16 (columns sorted by label), values one-hot int64 [N*K, N];
    cached under <cache_dir>/models/<backbone>/K-<shots>/aug/ (utils.py:284-332)."""
    model_dir_root = get_model_dir_root(cfg) + "/aug"
    os.makedirs(model_dir_root, exist_ok=True)

    def get_filename(kind):
        return f"{model_dir_root}/visual_mb_{kind}_aug_{cfg['augment_epoch']}_{cfg['shots']}_shots.pt"

    key_path, value_path = get_filename("keys"), get_filename("values")
    pdist.barrier()  # every rank sees the same cache state
    if dir_exists(key_path) and dir_exists(value_path):
        return torch.load(key_path, map_location=_here()), torch.load(value_path, map_location=_here())
    pdist.barrier()
    cache_keys, cache_values = [], []
    D = clip_model.visual.output_dim
    with torch.no_grad():
        for augment_idx in range(cfg["augment_epoch"]):
            train_features = []
            if pdist.is_main():
                print("Augment Epoch: {:} / {:}".format(augment_idx, cfg["augment_epoch"]))
            for images, target in tqdm(pdist.sharded_batches(train_loader_cache), disable=not pdist.is_main()):
                train_features.append(clip_model.encode_image(images.cuda()))
                if augment_idx == 0:
                    cache_values.append(target.cuda())
            train_features = torch.cat(train_features, dim=0) if train_features else \
                torch.empty((0, D), dtype=torch.float16, device="cuda")
            cache_keys.append(train_features.unsqueeze(0))
    cache_keys = torch.cat(cache_keys, dim=0).float().mean(dim=0).half()       # fp16 mean over augment epochs
    cache_values = torch.cat(cache_values, dim=0) if cache_values else torch.empty(0, dtype=torch.int64, device="cuda")
    cache_keys = pdist.all_gather_varlen(cache_keys)                           # the ranks' batch ranges, in order
    cache_values = pdist.all_gather_varlen(cache_values)
    cache_keys = nat.l2_normalize(cache_keys).permute(1, 0)
    index = torch.argsort(cache_values)
    cache_values = F.one_hot(cache_values[index])
    cache_keys = cache_keys[:, index]
    if pdist.is_main():
        torch.save(cache_keys, key_path)
        torch.save(cache_values, value_path)
    pdist.barrier()
    return cache_keys, cache_values


def pre_load_features(cfg, split, clip_model, loader):
    """L2-normalised image features fp16 [Q, D] + labels int64 [Q], cached as <split>_{features,labels}.pt
    (utils.py:335-361). The normalisation is fused into the encoder call."""
    root_dir_prefix = f"{get_model_dir_root(cfg)}/{split}"
    feature_path, label_path = f"{root_dir_prefix}_features.pt", f"{root_dir_prefix}_labels.pt"
    pdist.barrier()  # every rank sees the same cache state
    if dir_exists(feature_path) and dir_exists(label_path):
        if pdist.is_main():
            print(f"Loading cached features and labels from {root_dir_prefix}")
        return torch.load(feature_path, map_location=_here()), torch.load(label_path, map_location=_here())
    pdist.barrier()
    if pdist.is_main():
        print(f"Creating cached (features, labels) and saving to {root_dir_prefix}")
    features, labels = [], []
    with torch.no_grad():
        for images, target in tqdm(pdist.sharded_batches(loader), disable=not pdist.is_main()):
            features.append(nat.l2_normalize(clip_model.encode_image(images.cuda())))
            labels.append(target.cuda())
    features = torch.cat(features) if features else \
        torch.empty((0, clip_model.visual.output_dim), dtype=torch.float16, device="cuda")
    labels = torch.cat(labels) if labels else torch.empty(0, dtype=torch.int64, device="cuda")
    features, labels = pdist.all_gather_varlen(features), pdist.all_gather_varlen(labels)
    if pdist.is_main():
        os.makedirs(os.path.dirname(feature_path), exist_ok=True)
        torch.save(features, feature_path)
        torch.save(labels, label_path)
    pdist.barrier()
    return features, labels


def build_prototypes(embeddings_v, embeddings_t, K):
    """main.py:399-405 as a function: (z_img_proto [N, D], z_text_proto [N, D]) fp16."""
    V = embeddings_v.detach().half().contiguous()
    T = embeddings_t.detach().half().contiguous()
    N = T.shape[0]
    z_img, _ = nat.build_prototypes(V.view(N * K, -1), N, K, per_shot_norm=True)
    z_txt, _ = nat.build_prototypes(T, N, 1, per_shot_norm=False)
    return z_img, z_txt
