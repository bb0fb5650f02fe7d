"""proto-clip_b200: B200-native (sm_100a) few-shot inference hot path of Proto-CLIP.

Layout
  csrc/        hand-written CUDA kernels + the C ABI (libprotoclip_b200.so, include/protoclip_b200.h)
  _native.py   ctypes binding of the C ABI (fails loudly when the library or an sm_100 GPU is missing)
  clip/        drop-in for the reference ``clip`` package: load / tokenize / available_models, CLIP object
  model.py     drop-in ``Adapter`` / ``Adapter_FC`` nn.Modules (same state-dict keys)
  utils.py     drop-in ``P`` / ``build_cache_model`` / ``clip_classifier`` / ``pre_load_features`` ...
  main.py      drop-in CLI (same argparse / YAML surface), configs/ next to it
  dist.py      query-batch sharding + the single NCCL broadcast of the prototype memory
"""
__version__ = "0.1.0"
