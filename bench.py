#!/usr/bin/env python
"""Benchmark of the Proto-CLIP few-shot query path (BASELINE.json metric: query images/sec, ImageNet
1000-way 16-shot ViT-B/16).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

A step = one batch of --batch synthetic 224x224 query images per rank through
encode_image -> /norm -> Adapter_FC -> /norm -> P(alpha=0.5, beta=12) -> argmax over 1000 prototypes.
`value` is timed with inputs resident in HBM; `e2e` goes through the public API with pinned-host inputs
(H2D of every batch and D2H of the predictions inside the timed region). One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH = "ViT-B/16"
N_CLASSES, K_SHOTS, N_TEMPLATES = 1000, 16, 7
ALPHA, BETA = 0.5, 12.0
METRIC = "query images/sec, ImageNet 1000-way 16-shot ViT-B/16"

# BASELINE.json `configs`: c2 is the headline (the metric is quoted on it); c3 / c4 / c5 are reported as `secondary`
# entries of the same JSON line (query path only, random prototypes: the memory-bank build is c2's prelude).
# alpha / beta: the reference's configs/{imagenet,fewsol_198,sun397}.yml.
WORKLOADS = {
    "c2": dict(arch="ViT-B/16", n_classes=1000, adapter="fc", alpha=0.5, beta=12.0, batch=1024, micro_batch=512,
               name="imagenet 16-shot ViT-B/16 fc adapter, 1000-way eval (configs[1])"),
    "c3": dict(arch="ViT-L/14", n_classes=198, adapter="conv-3x", alpha=0.2, beta=12.0, batch=512, micro_batch=0,
               name="fewsol_198 16-shot ViT-L/14 conv-3x adapter, Proto-CLIP-F (configs[2])"),
    "c4": dict(arch="ViT-L/14@336px", n_classes=1000, adapter="fc", alpha=0.5, beta=12.0, batch=256, micro_batch=0,
               name="imagenet 16-shot ViT-L/14@336px fc adapter, query batch sharded over the ranks (configs[3])"),
    "c5": dict(arch="RN50x16", n_classes=397, adapter="conv-2x", alpha=1.0, beta=11.0, batch=256, micro_batch=0,
               name="sun397 16-shot RN50x16, main.qt.py path, conv-2x adapter (configs[4])"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1353.6), d.get("bf16_tflops", 1633.5), "measured"
    return 1400.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period: float = 0.2):
        self.index, self.period, self.rows, self._stop, self._t = index, period, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def synthetic_tokens(n_prompts: int, ctx: int, vocab: int, seed: int) -> torch.Tensor:
    """Prompt-shaped token rows (SOT, 4..12 word ids, EOT = max id, zero padding), as clip.tokenize emits."""
    gen = torch.Generator().manual_seed(seed)
    t = torch.zeros(n_prompts, ctx, dtype=torch.int64)
    n = torch.randint(4, 13, (n_prompts,), generator=gen)
    body = torch.randint(1, vocab - 2, (n_prompts, 12), generator=gen)
    for i in range(n_prompts):
        k = int(n[i])
        t[i, 0] = vocab - 2
        t[i, 1:1 + k] = body[i, :k]
        t[i, 1 + k] = vocab - 1
    return t


# --------------------------------------------------------------------------------------- this repo's arm
def run_ours(args):
    from proto_clip_b200 import _native as nat
    from proto_clip_b200 import dist as pdist
    from proto_clip_b200 import pipeline, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    rank, local_rank, world = pdist.init("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    c = synthetic.arch_config(ARCH)
    R, D = c["image_resolution"], c["embed_dim"]
    B, mb = args.batch, args.micro_batch

    sd = synthetic.make_state_dict(ARCH, 0)
    ctx = nat.Context(dev)
    ctx.bind_visual(sd)
    ctx.bind_text(sd)

    # ---- head state: built on rank 0 (support + text memory), one NCCL broadcast to everyone
    bases = synthetic.class_bases(N_CLASSES, R, seed=1, device=dev)           # same on every rank (same seed)
    numel = pipeline.HeadState.packed_numel(N_CLASSES, D, "fc")
    flat = None
    if args.lite:
        if rank == 0:
            # profiling mode (ncu launch lists): random prototypes, no support / text encoding, same query path
            g = torch.Generator(device=dev).manual_seed(9)
            V = nat.l2_normalize(torch.randn(N_CLASSES * K_SHOTS, D, generator=g, device=dev).half())
            T = nat.l2_normalize(torch.randn(N_CLASSES, D, generator=g, device=dev).half())
            adapter = synthetic.make_adapter_state_dict("fc", D, seed=4, out_gain=synthetic.trained_like_gain(D))
            flat = pipeline.build_head_state(V, T, N_CLASSES, K_SHOTS, "fc", adapter, ALPHA, BETA).pack()
    else:
        # one-off memory-bank build (utils.py:284-332, 256-273), sharded over the ranks with one all-gather per bank
        # (SURVEY f2): 16 000 support images (augment_epoch 1) and N x 7 prompts. A random-init text tower is not
        # class-aligned, so its output is timed / exercised only: the classifier's textual memory is the aligned
        # synthetic bank below.
        def support_slice(lo, hi):
            labels = torch.arange(lo, hi, device=dev) // K_SHOTS
            return synthetic.class_structured_images(bases, labels, seed=2 + lo)

        tokens = synthetic_tokens(N_CLASSES * N_TEMPLATES, c["context_length"], c["vocab_size"], 5)
        V, te = pipeline.build_memory_sharded(ctx, support_slice, tokens, micro_batch=mb,
                                              num_support=N_CLASSES * K_SHOTS)        # utils.py:310,319 / 266-267
        if rank == 0:
            _ = nat.l2_normalize(te.view(N_CLASSES, N_TEMPLATES, D).float().mean(dim=1).half())   # utils.py:268-269
            T = synthetic.aligned_text_memory(V, N_CLASSES, K_SHOTS, seed=6)
            adapter = synthetic.make_adapter_state_dict("fc", D, seed=4, out_gain=synthetic.trained_like_gain(D))
            head0 = pipeline.build_head_state(V, T, N_CLASSES, K_SHOTS, "fc", adapter, ALPHA, BETA)
            flat = head0.pack()
            assert flat.numel() == numel
    flat = pdist.broadcast_flat(flat, numel, torch.float16, dev, src=0)
    head = pipeline.HeadState.unpack(flat, N_CLASSES, D, "fc", ALPHA, BETA)
    clf = pipeline.FewShotClassifier(ctx, head, micro_batch=mb)

    # ---- query pool: 2 distinct batches per rank, class-structured, pinned host + device copies
    pool_dev, pool_host, pool_labels = [], [], []
    for j in range(2):
        labels = (torch.arange(B, device=dev) * 7 + 13 * j + 101 * rank) % N_CLASSES
        imgs = synthetic.class_structured_images(bases, labels, seed=1000 + 10 * rank + j)
        pool_dev.append(imgs)
        h = torch.empty(imgs.shape, dtype=torch.float32, pin_memory=True)
        h.copy_(imgs)
        pool_host.append(h)
        pool_labels.append(labels)
    del bases
    torch.cuda.synchronize()

    def step_resident(i):
        return clf.classify(pool_dev[i % 2])[1]

    # ---- roofline legs first, on an idle GPU (rank 0; the others wait at the barrier below): the attention kernel
    #      alone (burst clocks) and the four Linear launches of a block with the towers' own epilogues (burst loop +
    #      a 1.5 s loop that reaches the power-capped clocks of the step), CUDA events on the launching stream
    roof = attn = None
    g = c["image_resolution"] // c["vision_patch_size"]
    eff_mb = effective_micro_batch(B, g * g + 1, mb, dev)
    if rank == 0 and not args.lite:
        attn = attention_roofline(nat, dev, local_rank, eff_mb, g * g + 1, c["vision_width"] // 64)
        roof = gemm_roofline(ctx, dev, local_rank, eff_mb, g * g + 1, c["vision_width"])
    pdist.barrier()

    # ---- value: inputs resident in HBM, CUDA events on the launching stream, barrier + sync both sides
    for i in range(args.warmup):
        step_resident(i)
    torch.cuda.synchronize()
    pdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        e0.record()
        for i in range(args.steps):
            pred = step_resident(i)
        e1.record()
        torch.cuda.synchronize()
    pdist.barrier()
    ms_total = pdist.max_over_ranks(e0.elapsed_time(e1), dev)
    value = world * B * args.steps / (ms_total / 1e3)
    acc = (pred == pool_labels[(args.steps - 1) % 2]).float().mean().item()

    if args.lite:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": round(value, 1), "unit": "images/s", "n_gpus": world,
                              "steps": args.steps, "lite": True, "ms_per_step": round(ms_total / args.steps, 3)}), flush=True)
        return

    # ---- e2e: pinned host -> device copy of every batch (copy stream, double-buffered) + D2H of predictions
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty_like(pool_dev[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    out_host = torch.empty(B, dtype=torch.int64, pin_memory=True)

    def e2e_run(nsteps):
        cur = torch.cuda.current_stream(dev)
        for k in range(2):
            consumed[k].record(cur)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[0])
            stage[0].copy_(pool_host[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(nsteps):
            s = i % 2
            if i + 1 < nsteps:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[(i + 1) % 2])
                    stage[(i + 1) % 2].copy_(pool_host[(i + 1) % 2], non_blocking=True)
                    ready[(i + 1) % 2].record(copy_stream)
            cur.wait_event(ready[s])
            _, am, _ = clf.classify(stage[s])
            consumed[s].record(cur)
            out_host.copy_(am, non_blocking=True)
        cur.synchronize()

    # the bare pinned-host -> device copy of one batch (context for e2e: when it is longer than the compute of a step,
    # the end-to-end rate is the PCIe rate, whatever the kernels do)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_stream):
        stage[0].copy_(pool_host[0], non_blocking=True)
        c0.record(copy_stream)
        for k in range(4):
            stage[k % 2].copy_(pool_host[k % 2], non_blocking=True)
        c1.record(copy_stream)
    torch.cuda.synchronize()
    h2d_ms = c0.elapsed_time(c1) / 4
    e2e_run(max(2, args.warmup))
    torch.cuda.synchronize()
    pdist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_run(args.steps)
    t1.record()
    torch.cuda.synchronize()
    pdist.barrier()
    ms_e2e = pdist.max_over_ranks(t0.elapsed_time(t1), dev)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    # ---- the same step with the last visual block computed for every token, as the reference does before it keeps
    #      x[:, 0, :] (clip/model.py:232-233); reported beside `value`, which runs that block on the CLS rows alone
    full_last = None
    if True:
        n_ab = max(4, args.steps // 2)

        def timed_steps(n):
            for i in range(2):
                step_resident(i)
            torch.cuda.synchronize()
            pdist.barrier()
            e0.record()
            for i in range(n):
                p_ = step_resident(i)
            e1.record()
            torch.cuda.synchronize()
            pdist.barrier()
            return pdist.max_over_ranks(e0.elapsed_time(e1), dev) / n, p_

        # the clocks drift under the power cap during a run: interleave (full, CLS-only, full) after the two timed
        # loops (value, e2e) and compare within this A / B group
        ctx.set_full_last_block(True)
        ms_f1, pred_full = timed_steps(n_ab)
        ctx.set_full_last_block(None)
        ms_c, pred_cls = timed_steps(n_ab)
        ctx.set_full_last_block(True)
        ms_f2, _ = timed_steps(n_ab)
        ctx.set_full_last_block(None)
        ms_full = 0.5 * (ms_f1 + ms_f2)
        same = (pred_full == pred_cls).float().mean().item()
        full_last = {"value": round(world * B / (ms_full / 1e3), 1), "unit": "images/s", "steps": 2 * n_ab,
                     "ms_per_step": round(ms_full, 3),
                     "cls_only_in_the_same_group": {"value": round(world * B / (ms_c / 1e3), 1), "ms_per_step": round(ms_c, 3)},
                     "predictions_equal_to_cls_only": round(same, 6)}


    # ---- parity leg at every world size (rank 0, first cpu-sample queries of its batch 0) + the CPU baseline (N = 1)
    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu, parity = cpu_baseline_and_parity(sd, head, pool_host[0], clf, pool_dev[0], args.cpu_sample,
                                              timed=(world == 1))
    # ---- the other BASELINE.json configs (query path, resident images), same launch / timing rules
    del clf, pool_dev, stage
    torch.cuda.empty_cache()
    secondary = []
    if not args.no_secondary:
        for key in args.secondary.split(","):
            if key and key != "c2":
                secondary.append(run_workload_lite(key, args.secondary_steps, 2, rank, local_rank, world, dev))

    if rank == 0:
        layers = c["vision_layers"]
        n_mb = math.ceil(B / eff_mb)
        # per micro-batch: patchify, conv1 GEMM, embed+ln_pre (+ row statistics); per block 4 GEMMs (LayerNorm folded into two
        # of them) + attention; ln_post, proj GEMM, l2norm. Per step: adapter (2 GEMMs, 2 LN kernels), l2norm,
        # 1 similarity GEMM over the adjacent [z_img; z_txt] banks of the packed head state + softmax/argmax.
        launches = args.steps * (n_mb * (3 + 5 * layers + 3) + 7)
        sustained, burst, src = measured_peaks()
        head_flops = 4.0 * N_CLASSES * D + 2.0 * D * D / 4 * 2
        flops_ref = synthetic.vit_flops_per_image(ARCH) + head_flops          # what the reference computes per image
        flops_img = synthetic.vit_executed_flops_per_image(ARCH) + head_flops   # what this library executes per image
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOADS["c2"]["name"],
                       "backbone": ARCH, "n_classes": N_CLASSES, "shots": K_SHOTS, "adapter": "fc", "alpha": ALPHA,
                       "beta": BETA, "batch_per_gpu": B, "micro_batch": eff_mb, "image": "3x224x224 fp32",
                       "l2": "inputs larger than L2 (616 MB per batch, 2 alternating batches)",
                       "parallelism": f"dp{world} (query shards; one-off memory-bank build sharded with 2 all-gathers, then 1 NCCL broadcast of the head state; no collective in the timed step)"},
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": int(pool_host[0].numel() * 4),
                    "d2h_bytes_per_step": int(B * 8), "ms_per_step": round(ms_e2e / args.steps, 3),
                    "h2d_copy_alone_ms": round(h2d_ms, 3),
                    "h2d_copy_alone_gbps": round(pool_host[0].numel() * 4 / h2d_ms / 1e6, 1)},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "model_tflops": round(value * flops_img / 1e12, 1),
            "model_frac_of_sustained_peak": round(value / world * flops_img / 1e12 / sustained, 4),
            "flops_per_image": {"executed": round(flops_img / 1e9, 3), "reference": round(flops_ref / 1e9, 3), "unit": "GFLOP",
                                "note": "executed = reference minus the last block's non-CLS rows; model_tflops and model_frac use executed"},
            "full_last_block": full_last,
            "accuracy_on_synthetic_queries": round(acc, 4),
            "roofline": roof, "roofline_attention": attn, "cpu_baseline": cpu, "parity": parity, "peaks": src,
            "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def effective_micro_batch(B: int, L: int, mb: int, dev) -> int:
    """Images per encoder pass: the caller's --micro-batch, else the library's policy (csrc/api.cu pick_micro_batch: at
    most 12 row-block waves of the CTA-pair GEMM per pass, the batch cut into equal passes)."""
    if mb > 0:
        return mb
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    cap = 12 * max(1, (sms // 2) * 256 // L)
    passes = max(1, math.ceil(B / cap))
    return math.ceil(B / passes)


def flops_per_image(arch: str) -> float:
    from proto_clip_b200 import synthetic
    c = synthetic.arch_config(arch)
    return (synthetic.rn_flops_per_image(arch) if isinstance(c["vision_layers"], tuple)
            else synthetic.vit_executed_flops_per_image(arch))


def run_workload_lite(key: str, steps: int, warmup: int, rank: int, local_rank: int, world: int, dev) -> dict:
    """One of BASELINE.json's other configs through the same public path (FewShotClassifier.classify: encode_image ->
    /norm -> adapter -> /norm -> P -> argmax), images resident in HBM, random prototypes. Weak scaling: every rank
    runs `batch` query images per step; the value is the whole-job aggregate over the max-over-ranks device time."""
    from proto_clip_b200 import _native as nat
    from proto_clip_b200 import dist as pdist
    from proto_clip_b200 import pipeline, synthetic
    w = WORKLOADS[key]
    c = synthetic.arch_config(w["arch"])
    R, D, N = c["image_resolution"], c["embed_dim"], w["n_classes"]
    sd = synthetic.make_state_dict(w["arch"], 0)
    ctx = nat.Context(dev)
    ctx.bind_visual(sd)
    del sd
    g = torch.Generator(device=dev).manual_seed(9)
    V = nat.l2_normalize(torch.randn(N * K_SHOTS, D, generator=g, device=dev).half())
    T = nat.l2_normalize(torch.randn(N, D, generator=g, device=dev).half())
    kind = w["adapter"]
    adapter = synthetic.make_adapter_state_dict("fc" if kind == "fc" else "conv", D, seed=4,
                                                out_gain=synthetic.trained_like_gain(D))
    head = pipeline.build_head_state(V, T, N, K_SHOTS, kind, adapter, w["alpha"], w["beta"])
    clf = pipeline.FewShotClassifier(ctx, head, micro_batch=w["micro_batch"])
    B = w["batch"]
    pool = [torch.randn(B, 3, R, R, device=dev, generator=g) for _ in range(2)]  # 2 x >= 115 MB per rank
    for i in range(warmup):
        clf.classify(pool[i % 2])
    torch.cuda.synchronize()
    pdist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            clf.classify(pool[i % 2])
        e1.record()
        torch.cuda.synchronize()
    pdist.barrier()
    ms = pdist.max_over_ranks(e0.elapsed_time(e1), dev)
    value = world * B * steps / (ms / 1e3)
    sustained, _, src = measured_peaks()
    fl = flops_per_image(w["arch"])
    out = {"workload": w["name"], "key": key, "backbone": w["arch"], "n_classes": N, "adapter": kind,
           "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "batch_per_gpu": B, "micro_batch": w["micro_batch"], "ms_per_step": round(ms / steps, 3),
           "model_tflops": round(value * fl / 1e12, 1),
           "model_frac_of_sustained_peak": round(value / world * fl / 1e12 / sustained, 4),
           "clocks": clocks.summary(), "data": "synthetic, resident in HBM, random prototypes"}
    del clf, head, ctx, pool
    torch.cuda.empty_cache()
    return out


def timed_loop(fn, min_ms: float, index: int):
    """fn back to back for at least min_ms of device time (CUDA events on the launching stream); returns
    (ms per call, median SM clock sampled by a background nvidia-smi thread while the loop ran, or None when the
    loop is shorter than the sampling period)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    iters = max(20, int(min_ms / max(e0.elapsed_time(e1) / 10, 1e-3)))
    with ClockSampler(index, period=0.05) as clocks:
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, clocks.summary()["sm_mhz"]


# (M, d) -> dram__bytes_read.sum + dram__bytes_write.sum of the block's four Linear launches (ncu --set full, cold caches)
GEMM_TRAFFIC = {(18912, 768): 388.4e6,      # profiles/r01_ncu_gemm_summary.txt (plain epilogues)
                (100864, 768): 2689.3e6,   # profiles/r02_ncu_block_summary.txt (the towers' epilogues, 512 images)
                (201728, 768): 5513.0e6}   # profiles/r02_ncu_block1024_summary.txt (1024 images in one pass)


def gemm_roofline(ctx, dev, index, B, L, d):
    """The four Linear launches of one ResidualAttentionBlock exactly as the towers run them
    (pc_resblock_forward_parts, parts = GEMMs: EPI_LN_BIAS QKV, EPI_BIAS_RES + row statistics out_proj, EPI_LN_QGELU
    c_fc, EPI_BIAS_RES + statistics c_proj; clip/model.py:187-190) at the micro-batch's M = B*L. Timed twice: a short
    loop on an idle GPU (burst clocks -> burst peak) and a >= 1.5 s loop (power-capped clocks, as inside the step ->
    sustained peak). `frac` is the sustained one."""
    from proto_clip_b200 import _native as nat
    sustained, burst, src = measured_peaks()
    torch.manual_seed(0)
    x = (torch.randn(B * L, d, device=dev) * 0.5).half()
    x0 = x.clone()
    ctx.resblock_forward_parts(nat.PC_TOWER_VISUAL, 0, x, B, L, False, parts=31, chained=False)  # leaves valid statistics

    def block():
        ctx.resblock_forward_parts(nat.PC_TOWER_VISUAL, 0, x, B, L, False, parts=29, chained=True)

    time.sleep(0.5)
    ms_b, mhz_b = timed_loop(block, 10.0, index)
    x.copy_(x0)
    ms_s, mhz_s = timed_loop(block, 1500.0, index)
    M = B * L
    flops = 24.0 * M * d * d
    tf_b, tf_s = flops / (ms_b / 1e3) / 1e12, flops / (ms_s / 1e3) / 1e12
    # context, not a target: the vendor library (torch.matmul -> cuBLAS, fp16, NO bias / LayerNorm / GELU / residual
    # epilogue) on the same four problem shapes, same loop. MEASURED_PEAKS' denominators are 8192^3 problems; these
    # shapes pay a launch, a pipeline fill and an un-overlapped last epilogue every ~100-350 us.
    a1 = torch.randn(M, d, device=dev).half()
    a4 = torch.randn(M, 4 * d, device=dev).half()
    ws_ = [torch.randn(n, k, device=dev).half() * 0.03 for n, k in ((3 * d, d), (d, d), (4 * d, d), (d, 4 * d))]
    outs = [torch.empty(M, w.shape[0], device=dev, dtype=torch.float16) for w in ws_]

    def lib_block():
        torch.matmul(a1, ws_[0].t(), out=outs[0])
        torch.matmul(a1, ws_[1].t(), out=outs[1])
        torch.matmul(a1, ws_[2].t(), out=outs[2])
        torch.matmul(a4, ws_[3].t(), out=outs[3])

    time.sleep(0.5)
    ms_lb, _ = timed_loop(lib_block, 10.0, index)
    ms_ls, mhz_l = timed_loop(lib_block, 1500.0, index)
    del a1, a4, ws_, outs
    # DRAM read + write bytes of the same four launches from one `ncu --set full` capture with cold caches
    # (profiles/README.md names the file); only valid for the shape it was captured at.
    traffic = GEMM_TRAFFIC.get((M, d))
    return {"bound": "tensor", "achieved": round(tf_s, 1), "peak": sustained, "unit": "TFLOP/s",
            "frac": round(tf_s / sustained, 4), "traffic": traffic,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the four launches, ncu --set full, cold "
                              "caches (profiles/README.md names the capture); null when no capture exists for this M",
            "algorithmic_bytes_per_4_launches": 2.0 * (M * d * (1 + 3 + 1 + 1 + 1 + 1 + 4 + 4 + 1 + 1) + 12 * d * d),
            "kernel": f"gemm_tn_kernel<pair, EPI_LN_BIAS | EPI_BIAS_RES | EPI_LN_QGELU | EPI_BIAS_RES>: the 4 Linear "
                      f"launches of one ResidualAttentionBlock as the tower runs them, M={M}, d={d}",
            "flops_per_4_launches": flops, "ms_per_4_launches": round(ms_s, 4), "sm_mhz": mhz_s,
            "peak_kind": f"bf16 sustained ({src}); loop of >= 1.5 s, power-capped clocks like the step",
            "cublas_same_shapes_no_epilogue": {"achieved_sustained": round(flops / (ms_ls / 1e3) / 1e12, 1),
                                               "achieved_burst": round(flops / (ms_lb / 1e3) / 1e12, 1),
                                               "ms_per_4_launches": round(ms_ls, 4), "sm_mhz": mhz_l},
            "burst": {"achieved": round(tf_b, 1), "peak": burst, "frac": round(tf_b / burst, 4),
                      "ms_per_4_launches": round(ms_b, 4), "sm_mhz": mhz_b,
                      "peak_kind": f"bf16 burst ({src}); 10 ms loop on an idle GPU"}}


def attention_roofline(nat, dev, index, B, L, heads):
    """The attention kernel alone at the micro-batch's shape: algorithmic 4*B*heads*L^2*64 flop per launch (QK^T + PV).
    Timed on an idle GPU before the step loops. `frac` follows MEASURED_PEAKS.json's burst protocol (a kernel timed
    alone: best of 10 short groups of launches, CUDA events) against the burst peak; the >= 300 ms back-to-back loop is
    reported beside it with the SM clock sampled while it ran (right after the GEMM-heavy step the power-capped
    1.4 GHz clock would inflate either by a third, which is why this leg runs first)."""
    sustained, burst, src = measured_peaks()
    torch.manual_seed(1)
    qkv = torch.randn(B * L, 3 * heads * 64, device=dev).half()
    time.sleep(0.5)
    for _ in range(5):
        nat.attention(qkv, B, L, heads, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    groups = []
    for _ in range(10):
        e0.record()
        for _ in range(10):
            nat.attention(qkv, B, L, heads, False)
        e1.record()
        torch.cuda.synchronize()
        groups.append(e0.elapsed_time(e1) / 10)
    ms = min(groups)
    ms_loop, mhz = timed_loop(lambda: nat.attention(qkv, B, L, heads, False), 300.0, index)
    flops = 4.0 * B * heads * L * L * 64
    achieved = flops / (ms / 1e3) / 1e12
    exps = float(B) * heads * L * L
    name = "attention6_kernel (whole-row S in TMEM, L <= 208)" if L <= 208 else "attention_kernel (64-key blocks)"
    return {"bound": "tensor", "achieved": round(achieved, 1), "peak": burst, "unit": "TFLOP/s",
            "frac": round(achieved / burst, 4), "traffic": None,
            "kernel": f"{name}<causal=false>: B={B}, L={L}, heads={heads}, head_dim=64",
            "flops_per_launch": flops, "us_per_launch": round(ms * 1e3, 2),
            "peak_kind": f"bf16 burst ({src}); best of 10 groups of 10 launches on an idle GPU",
            "loop_300ms": {"us_per_launch": round(ms_loop * 1e3, 2), "achieved": round(flops / (ms_loop / 1e3) / 1e12, 1),
                           "frac_of_burst_peak": round(flops / (ms_loop / 1e3) / 1e12 / burst, 4), "sm_mhz": mhz},
            "hbm_bytes_per_launch": float(B * L * heads * 64 * 2 * 4),
            "note": "head_dim 64: one exp2 per 256 tensor flops; the MUFU pipe (16 exp2/clk/SM) alone bounds this shape "
                    "at 0.69 of the tensor peak, the M = 128 / N <= 208 UMMA shapes at 0.55; at this batch qkv "
                    "(B*L*3d fp16) no longer fits the L2 and the kernel also moves hbm_bytes_per_launch",
            "gexp_per_s": round(exps / (ms / 1e3) / 1e9, 1)}


def reference_step_fn(sd, asd, D, zi, zt, alpha, beta):
    """(step(images) -> (p, argmax), kind): the reference's OWN modules on the CPU when they can be imported
    (/root/reference, or the byte-for-byte snapshot oracle/make_ref.py left in oracle/_ref/): clip/model.py build_model
    + encode_image (clip/model.py:338-339, fp32 like clip.load(device='cpu'), clip/clip.py:137-138), utils.py:351-352
    normalisation, model.py:81-95 Adapter_FC, utils.py:225-244 P, main.py:438 argmax. Else the oracle port."""
    from oracle import protoclip_oracle as O
    from oracle import reference_shims
    if reference_shims.available():
        ref = reference_shims.reference()
        model = ref.clip_model.build_model({k: v.clone() for k, v in sd.items()}).float()
        adapter = ref.model.Adapter_FC(D, dtype=torch.float32)
        adapter.load_state_dict({k: v.float() for k, v in asd.items()})

        def step(images):
            f = model.encode_image(images)
            f = f / f.norm(dim=-1, keepdim=True)
            q = adapter(f)
            q = q / q.norm(dim=-1, keepdim=True)
            p = ref.utils.P(q, zi, zt, alpha, beta)
            return p, p.max(1)[1]
        where = "oracle/_ref snapshot of the reference" if reference_shims.is_snapshot() else reference_shims.REFERENCE_ROOT
        return step, "reference", f"clip/model.py + model.py + utils.py of {where}"

    def step(images):
        p, pr, _ = O.classify_queries(sd, asd, "fc", images, zi, zt, alpha, beta, "fp32")
        return p, pr
    return step, "port", "oracle/protoclip_oracle.py"


def cpu_baseline_and_parity(sd, head, host_images, clf, dev_images, sample, timed=True):
    """The reference path (fp32 like clip.load(device='cpu')) on this box's host cores, on the first `sample` queries
    of batch 0; also the argmax parity of the CUDA path on exactly those queries."""
    torch.set_num_threads(os.cpu_count() or 1)
    zi, zt = head.z_img.float().cpu(), head.z_txt.float().cpu()
    asd = {k: v.cpu() for k, v in head.adapter.items()}
    step, kind, what = reference_step_fn(sd, asd, zi.shape[1], zi, zt, ALPHA, BETA)
    imgs = host_images[:sample].clone()
    bs = 32
    with torch.no_grad():
        if timed:
            step(imgs[:bs])  # warm-up
        best, preds, ps = float("inf"), [], []
        for rep in range(2 if timed else 1):
            t = time.perf_counter()
            preds, ps = [], []
            for i in range(0, sample, bs):
                p, pr = step(imgs[i:i + bs])
                preds.append(pr)
                ps.append(p)
            best = min(best, time.perf_counter() - t)
    pred_cpu, p_cpu = torch.cat(preds), torch.cat(ps)
    p_gpu, pred_gpu, _ = clf.classify(dev_images[:sample], want_p=True)
    top2 = p_cpu.topk(2, dim=1).values
    mism = int((pred_gpu.cpu() != pred_cpu).sum())
    cpu = None
    if timed:
        cpu = {"value": round(sample / best, 2), "unit": "images/s", "cores": os.cpu_count(), "kind": kind,
               "sample": f"first {sample} queries of batch 0, batch 32, fp32, best of 2 after 1 warm-up ({what})"}
    parity = {"checked": sample, "against": kind, "argmax_mismatches": mism,
              "max_abs_dp": round((p_gpu.cpu() - p_cpu).abs().max().item(), 6),
              "min_top1_top2_margin_ref": round((top2[:, 0] - top2[:, 1]).min().item(), 6)}
    return cpu, parity


# --------------------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores. Uses the real reference modules
    when /root/reference is mounted (authoring container), else the oracle port (GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import protoclip_oracle as O
    from oracle import reference_shims
    from proto_clip_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_grad_enabled(False)
    c = synthetic.arch_config(ARCH)
    R, D = c["image_resolution"], c["embed_dim"]
    sd = synthetic.make_state_dict(ARCH, 0)
    asd = synthetic.make_adapter_state_dict("fc", D, seed=4, out_gain=synthetic.trained_like_gain(D))
    gen = torch.Generator().manual_seed(3)
    zi = torch.nn.functional.normalize(torch.randn(N_CLASSES, D, generator=gen), dim=-1).half().float()
    zt = torch.nn.functional.normalize(torch.randn(N_CLASSES, D, generator=gen), dim=-1).half().float()
    bs = args.ref_batch
    bases = synthetic.class_bases(8, R, seed=1)
    imgs = synthetic.class_structured_images(bases, torch.arange(bs) % 8, seed=3)
    step_fn, kind, what = reference_step_fn(sd, asd, D, zi, zt, ALPHA, BETA)

    def step():
        return step_fn(imgs)[1]

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = min(args.steps, args.ref_max_steps)
    t = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t
    v = round(bs * steps / dt, 2)
    sample = f"{steps} steps of {bs} synthetic queries, fp32, {os.cpu_count()} threads ({what})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": round(dt / steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "imagenet 16-shot ViT-B/16 fc adapter, 1000-way eval (configs[1])", "backbone": ARCH,
                   "batch": bs, "device": "cpu"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": os.cpu_count(), "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="query images per rank per step (main.py:505 uses 1024)")
    ap.add_argument("--micro-batch", type=int, default=0, help="images per encoder pass (0 = library default, 96)")
    ap.add_argument("--cpu-sample", type=int, default=128, help="queries timed on the host CPU for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lite", action="store_true", help="profiling mode: random head state, value arm only")
    ap.add_argument("--secondary", default="c3,c4,c5", help="other BASELINE.json configs reported in `secondary`")
    ap.add_argument("--secondary-steps", type=int, default=3)
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=32)
    ap.add_argument("--ref-max-steps", type=int, default=6)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
