"""Drop-in for the hot-path callers of the reference's toolkit:

  load_pretrained_mb_and_adapters   toolkit/proto_clip_toolkit/utils/model_utils.py:12-70
  pre_load_features_without_cache   toolkit/proto_clip_toolkit/utils/model_utils.py:72-83
  ProtoClipClassifier               toolkit/proto_clip_toolkit/ros/utils/proto_clip_classifier.py:25-158
                                    (construction, prototypes, classify_objects top-k; no ROS / drawing code)
  test_ood_performance              toolkit/proto_clip_toolkit/utils/ood_utils.py:58-111

Same names, argument meaning and file formats; encoders, adapter, prototypes and P() run on libprotoclip_b200.
ROS nodes, speech recognition, part-of-speech tagging, t-SNE and the PIL result canvas are outside the hot path.
"""
from __future__ import annotations

import json
import os
import random
from typing import List, Optional, Sequence

import numpy as np
import torch
import yaml

try:
    from . import _native as nat
    from . import clip
    from .model import Adapter, Adapter_FC
    from .utils import P, build_prototypes, get_model_dir_root, get_seed, pre_load_features
except ImportError:  # pragma: no cover
    from proto_clip_b200 import _native as nat
    from proto_clip_b200 import clip
    from proto_clip_b200.model import Adapter, Adapter_FC
    from proto_clip_b200.utils import P, build_prototypes, get_model_dir_root, get_seed, pre_load_features


def _make_adapter(adapter_type: Optional[str], ndim: int):
    if adapter_type is None:
        raise Exception("Please mention the adapter type in the args or in the config file.")
    if "conv" in adapter_type:
        return Adapter(ndim, c_type=adapter_type, dtype=torch.half).cuda()
    if adapter_type == "fc":
        return Adapter_FC(ndim, dtype=torch.half).cuda()
    raise NameError(f"unknown adapter alias {adapter_type!r}: expected 'conv-3x', 'conv-2x' or 'fc'")


def load_pretrained_mb_and_adapters(config=None, memory_bank_v_path=None, memory_bank_t_path=None, adapter_type=None,
                                    adapter_weights_path=None):
    """Trained visual / textual memory banks (`*_v.pt`, `*_t.pt`: pickled fp16 nn.Parameters) and the query adapter,
    located either through a run config (cache layout of main.py:383-390) or through explicit paths. The reference
    passes `config['adapter']` to the conv Adapter even on the explicit-path branch where config is None
    (model_utils.py:55); here that branch uses `adapter_type`, which is what the caller asked for."""
    if config:
        model_dir = f"{get_model_dir_root(config)}/alpha-beta/{config['alpha']}-{config['beta']}"
        prefix = f"best_lr_{config['lr']}_aug_{config['augment_epoch']}_epochs_{config['train_epoch']}"
        memory_bank_v_path = os.path.join(model_dir, f"{prefix}_v.pt")
        memory_bank_t_path = os.path.join(model_dir, f"{prefix}_t.pt")
        adapter_weights_path = os.path.join(model_dir, f"{prefix}_a.pt")
        adapter_type = config["adapter"]
    with torch.no_grad():
        try:
            embeddings_v = torch.load(memory_bank_v_path, weights_only=False)
            embeddings_t = torch.load(memory_bank_t_path, weights_only=False)
        except Exception:
            raise FileNotFoundError(f"File does not exist: {memory_bank_v_path} and {memory_bank_t_path}")
        adapter = _make_adapter(adapter_type, embeddings_v.shape[1])
        try:
            adapter.load_state_dict(torch.load(adapter_weights_path, weights_only=False))
        except FileNotFoundError:
            raise FileNotFoundError(f"File does not exist: {adapter_weights_path}")
    return embeddings_v, embeddings_t, adapter


def pre_load_features_without_cache(clip_model, loader):
    """L2-normalised CLIP features of an image-only loader, not cached (model_utils.py:72-83)."""
    features = []
    with torch.no_grad():
        for images in loader:
            features.append(nat.l2_normalize(clip_model.encode_image(images.cuda())))
    return torch.cat(features)


class ProtoClipClassifier:
    """Real-world object classifier of the robot demo: prototypes from a trained memory bank, top-k Proto-CLIP
    predictions with class names for a list of cropped RGB images."""

    def __init__(self, args):
        assert os.path.exists(args.config)
        self.cfg = yaml.load(open(args.config, "r"), Loader=yaml.Loader)
        print("\nRunning configs.")
        print(self.cfg, "\n")
        self.clip_model, self.preprocess = clip.load(self.cfg["backbone"])
        self.clip_model.eval()
        seed = get_seed()
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
        self.test_bs, self.n_workers = 1, 1
        self.class_id_mapping = {}
        self.parse_splits_file(args.splits_path)
        self._load_trained_models_and_embeddings(args)

    def _load_trained_models_and_embeddings(self, args):
        with torch.no_grad():
            embeddings_v, embeddings_t, self.adapter = load_pretrained_mb_and_adapters(
                adapter_type=args.adapter if getattr(args, "adapter", None) else self.cfg["adapter"],
                memory_bank_v_path=args.memory_bank_v_path, memory_bank_t_path=args.memory_bank_t_path,
                adapter_weights_path=args.adapter_weights_path)
            # per-shot normalise, class mean, renormalise / normalised text memory (proto_clip_classifier.py:62-71)
            self.z_img_proto, self.z_text_proto = build_prototypes(embeddings_v.data.cuda(), embeddings_t.data.cuda(),
                                                                   self.cfg["shots"])

    def parse_splits_file(self, config_path):
        """class id -> class name from the `train` entries of a split file (proto_clip_classifier.py:73-80)."""
        with open(config_path) as f:
            data = json.load(f)
        for entry in data["train"]:
            self.class_id_mapping[entry[1]] = entry[2]

    def _features(self, cropped_images) -> torch.Tensor:
        if isinstance(cropped_images, torch.Tensor):  # already preprocessed [B, 3, R, R]
            batches = [cropped_images[i:i + 64] for i in range(0, cropped_images.shape[0], 64)]
        else:  # HxWx3 uint8 arrays, as the segmentation node hands them over (image_utils.py:8-25): the reference
            # runs Image.fromarray + the host transform per crop; the same arithmetic runs on the GPU here
            n = self.clip_model.visual.input_resolution
            batch = torch.empty(len(cropped_images), 3, n, n, device="cuda")
            for i, im in enumerate(cropped_images):
                self.clip_model.preprocess_gpu(im, out=batch[i])
            batches = [batch[i:i + 64] for i in range(0, batch.shape[0], 64)]
        return pre_load_features_without_cache(self.clip_model, batches)

    def classify_objects(self, cropped_images, log=False, rgb_image=None):
        """(top-k class names, top-k probabilities [B, k]) — proto_clip_classifier.py:132-158."""
        test_features = self._features(cropped_images)
        with torch.no_grad():
            test_features = nat.l2_normalize(self.adapter(test_features))
            p = P(test_features, self.z_img_proto, self.z_text_proto, self.cfg["alpha"], self.cfg["beta"])
            top_k_class_probs, top_k_class_idxs = p.topk(k=self.cfg["top_k"], dim=1)
            top_k_class_names = [[self.class_id_mapping[x.item()].replace("_", " ") for x in row]
                                 for row in top_k_class_idxs]
            if log:
                import time
                os.makedirs("./ros-demo-logs", exist_ok=True)
                np.save(f"./ros-demo-logs/experiment_pred_{int(time.time())}.npy",
                        {"rgb_image": rgb_image, "cropped_images": cropped_images, "top_k_classes": top_k_class_names,
                         "top_k_probs": top_k_class_probs.cpu().numpy()})
            return top_k_class_names, top_k_class_probs


def test_ood_performance(cfg, test_dataset_name, n_workers, test_bs, memory_bank_v_path=None, memory_bank_t_path=None,
                         adapter_type=None, adapter_weights_path=None, test_loader=None):
    """Accuracy (%) of a trained Proto-CLIP on an out-of-distribution test set (ood_utils.py:58-111). `test_loader`
    (batches of (images, labels)) may be passed directly; otherwise the reference's two aliases are built with
    torchvision's ImageFolder / imagenetv2_pytorch, which must be installed."""
    clip_model, preprocess = clip.load(cfg["backbone"])
    clip_model.eval()
    seed = get_seed()
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    if test_loader is None:
        print("Preparing dataset.")
        if test_dataset_name == "imagenet_v2":
            from imagenetv2_pytorch import ImageNetV2Dataset
            test_dataset = ImageNetV2Dataset("matched-frequency", transform=preprocess)
        elif test_dataset_name == "imagenet_sketch":
            from torchvision.datasets import ImageFolder
            test_dataset = ImageFolder("./DATA/sketch", transform=preprocess)
        else:
            raise ValueError(f"unknown OOD dataset alias {test_dataset_name!r}")
        test_loader = torch.utils.data.DataLoader(test_dataset, batch_size=test_bs, num_workers=n_workers, shuffle=False)
    test_features, test_labels = pre_load_features(cfg, "test", clip_model, test_loader)
    with torch.no_grad():
        print("Testing...")
        embeddings_v, embeddings_t, adapter = load_pretrained_mb_and_adapters(
            memory_bank_v_path=memory_bank_v_path, memory_bank_t_path=memory_bank_t_path, adapter_type=adapter_type,
            adapter_weights_path=adapter_weights_path)
        z_img_proto, z_text_proto = build_prototypes(embeddings_v.data.cuda(), embeddings_t.data.cuda(), cfg["shots"])
        test_features = nat.l2_normalize(adapter(test_features))
        p = P(test_features, z_img_proto, z_text_proto, cfg["alpha"], cfg["beta"])
        test_acc = (p.max(1)[1] == test_labels).float().mean() * 100.0
    print("**** Proto-CLIP's OOD test accuracy: {:.2f}% ****".format(test_acc))
    return float(test_acc)


test_ood_performance.__test__ = False  # a reference function name, not a pytest case
