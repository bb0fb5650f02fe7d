"""Bring-up probe: runs each primitive of libprotoclip_b200 in its own process (a trapped kernel kills only
its case) and prints error statistics against torch fp32 on the same GPU. Not a test; tests/ holds those.

    python tools/gpu_probe.py            # all cases
    python tools/gpu_probe.py --case gemm_768
"""
import argparse
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def stats(name, got, ref):
    import torch
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    bad = (~torch.isfinite(got)).sum().item()
    print(f"[{name}] max_abs_err={err.max().item():.4e} mean_abs_err={err.mean().item():.4e} "
          f"ref_absmax={denom:.4e} rel={err.max().item() / denom:.4e} nonfinite={bad}", flush=True)
    if err.max().item() / denom > 2e-2 or bad:
        # where are the errors? (row / column histogram helps spot swizzle / descriptor mistakes)
        e2 = err.reshape(-1, err.shape[-1])
        rows = (e2.max(dim=1).values > 1e-2 * denom).nonzero().flatten()[:16].tolist()
        cols = (e2.max(dim=0).values > 1e-2 * denom).nonzero().flatten()[:32].tolist()
        print(f"[{name}]   bad rows (first 16): {rows}\n[{name}]   bad cols (first 32): {cols}", flush=True)
        print(f"[{name}]   got[0,:8]={got.reshape(-1, got.shape[-1])[0, :8].tolist()}")
        print(f"[{name}]   ref[0,:8]={ref.reshape(-1, ref.shape[-1])[0, :8].tolist()}")


class Clocks:
    """Median SM clock (MHz) sampled with nvidia-smi while a timed loop runs."""

    def __enter__(self):
        import threading
        self.vals, self._stop = [], threading.Event()

        def run():
            while not self._stop.is_set():
                try:
                    o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                        "-i", "0"], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.vals.append((float(o[0]), float(o[1])))
                except Exception:
                    pass
                self._stop.wait(0.05)
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def mhz(self):
        v = sorted(x[0] for x in self.vals)
        return v[len(v) // 2] if v else float("nan")

    def watts(self):
        v = sorted(x[1] for x in self.vals)
        return v[len(v) // 2] if v else float("nan")


def timed_loop(fn, flops, label, seconds=0.6):
    """Sustained timing: run fn back to back for ~`seconds`, CUDA events around the loop, clocks sampled meanwhile."""
    import torch
    for _ in range(5):
        fn()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(20):
        fn()
    t1.record(); torch.cuda.synchronize()
    burst_ms = t0.elapsed_time(t1) / 20
    iters = max(20, int(seconds * 1e3 / burst_ms))
    with Clocks() as c:
        t0.record()
        for _ in range(iters):
            fn()
        t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / iters
    tf = flops / ms / 1e9
    mhz = c.mhz()
    peak = 8192.0 * 148 * mhz * 1e6 / 1e12
    print(f"[{label}] burst {burst_ms * 1e3:.1f} us ({flops / burst_ms / 1e9:.0f} TF)  sustained {ms * 1e3:.1f} us "
          f"{tf:.0f} TFLOP/s @ {mhz:.0f} MHz {c.watts():.0f} W -> {100 * tf / peak:.1f}% of clock peak", flush=True)


def case_gemm(M, N, K, epi="bias"):
    import torch
    from proto_clip_b200 import _native as nat
    torch.manual_seed(0)
    x = (torch.randn(M, K, device="cuda") * 0.5).half()
    w = (torch.randn(N, K, device="cuda") * 0.05).half()
    b = torch.randn(N, device="cuda").half()
    r = torch.randn(M, N, device="cuda").half()
    acc = x.float() @ w.float().t() + b.float()
    if epi == "bias":
        got, ref = nat.linear(x, w, b, nat.EPI_BIAS), acc
    elif epi == "gelu":
        h = acc.half().float()
        got, ref = nat.linear(x, w, b, nat.EPI_BIAS_QUICKGELU), h * torch.sigmoid(1.702 * h)
    elif epi == "res":
        got, ref = nat.linear(x, w, b, nat.EPI_BIAS_RESIDUAL, residual=r), acc.half().float() + r.float()
    elif epi == "f32":
        got, ref = nat.linear(x, w, None, nat.EPI_F32), x.float() @ w.float().t()
    torch.cuda.synchronize()
    stats(f"gemm {M}x{N}x{K} {epi}", got, ref)
    code = {"bias": 0, "gelu": 1, "res": 2, "f32": 3}[epi]
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if epi == "f32" else torch.float16)
    timed_loop(lambda: nat.linear(x, w, b if epi != "f32" else None, code, residual=r if epi == "res" else None, out=out),
               2.0 * M * N * K, f"gemm {M}x{N}x{K} {epi}")


def case_attn(B, L, heads, causal):
    import torch
    from proto_clip_b200 import _native as nat
    torch.manual_seed(1)
    d = heads * 64
    qkv = (torch.randn(B * L, 3 * d, device="cuda")).half()
    got = nat.attention(qkv, B, L, heads, causal)
    q, k, v = [t.reshape(B, L, heads, 64).permute(0, 2, 1, 3).float() for t in qkv.split(d, dim=1)]
    s = (q @ k.transpose(-1, -2)) * 0.125
    if causal:
        s = s + torch.full((L, L), float("-inf"), device="cuda").triu(1)
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * L, d)
    torch.cuda.synchronize()
    stats(f"attn B{B} L{L} h{heads} causal={causal}", got, ref)
    timed_loop(lambda: nat.attention(qkv, B, L, heads, causal), 4.0 * B * heads * L * L * 64, f"attn B{B} L{L}", 0.3)


def case_block_parts(B=96, arch="ViT-B/16"):
    """Each launch of one ResidualAttentionBlock on its own, with the epilogue the tower uses
    (pc_resblock_forward_parts), next to the plain-epilogue GEMM of the same shape."""
    import torch
    from proto_clip_b200 import _native as nat
    from proto_clip_b200 import synthetic
    c = synthetic.arch_config(arch)
    d, L = c["vision_width"], (c["image_resolution"] // c["vision_patch_size"]) ** 2 + 1
    sd = synthetic.make_state_dict(arch, 0)
    ctx = nat.Context(torch.device("cuda:0"))
    ctx.bind_visual(sd)
    x = (torch.randn(B * L, d, device="cuda") * 0.5).half()
    ctx.resblock_forward_parts(nat.PC_TOWER_VISUAL, 0, x, B, L, False, 31, False)
    M = B * L
    for name, mask, fl in (("qkv   LN_BIAS ", 1, 2.0 * M * 3 * d * d), ("out   RES+stat", 4, 2.0 * M * d * d),
                           ("c_fc  LN_QGELU", 8, 2.0 * M * 4 * d * d), ("c_proj RES+stat", 16, 2.0 * M * 4 * d * d),
                           ("4 gemms       ", 29, 24.0 * M * d * d), ("attention     ", 2, 4.0 * B * (d // 64) * L * L * 64)):
        timed_loop(lambda: ctx.resblock_forward_parts(nat.PC_TOWER_VISUAL, 0, x, B, L, False, mask, True), fl,
                   f"block {arch} B{B} {name}", 0.5)


def case_rows():
    import torch
    from proto_clip_b200 import _native as nat
    torch.manual_seed(2)
    for d in (512, 768, 1024):
        x = (torch.randn(1000, d, device="cuda") * 2 + 0.3).half()
        g = torch.randn(d, device="cuda"); b = torch.randn(d, device="cuda")
        stats(f"layernorm d{d}", nat.layernorm(x, g, b), torch.nn.functional.layer_norm(x.float(), (d,), g, b))
        stats(f"l2norm d{d}", nat.l2_normalize(x), x.float() / x.float().norm(dim=-1, keepdim=True))


def case_cublas():
    """torch.matmul (cuBLAS) fp16 on the encoder's GEMM shapes: the library baseline for the same problems."""
    import torch
    for (M, N, K) in [(18912, 2304, 768), (18912, 768, 768), (18912, 3072, 768), (18912, 768, 3072), (8192, 8192, 8192)]:
        x = (torch.randn(M, K, device="cuda") * 0.5).half()
        w = (torch.randn(N, K, device="cuda") * 0.05).half()
        out = torch.empty(M, N, device="cuda", dtype=torch.float16)
        timed_loop(lambda: torch.matmul(x, w.t(), out=out), 2.0 * M * N * K, f"cublas {M}x{N}x{K}")


def case_tower(arch, B, micro_batch=0):
    """encode_image throughput of a whole visual tower (resident images), against the tower's 2*MAC count."""
    import torch
    from proto_clip_b200 import _native as nat
    from proto_clip_b200 import synthetic
    c = synthetic.arch_config(arch)
    sd = synthetic.make_state_dict(arch, 0)
    ctx = nat.Context(torch.device("cuda:0"))
    ctx.bind_visual(sd)
    R = c["image_resolution"]
    images = torch.randn(B, 3, R, R, device="cuda")
    out = torch.empty(B, c["embed_dim"], device="cuda", dtype=torch.float16)
    flops = (synthetic.rn_flops_per_image(arch) if isinstance(c["vision_layers"], tuple)
             else synthetic.vit_flops_per_image(arch)) * B
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        ctx.encode_image(images, l2norm=True, micro_batch=micro_batch, out=out)
    t0.record()
    n = 3
    for _ in range(n):
        ctx.encode_image(images, l2norm=True, micro_batch=micro_batch, out=out)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / n
    print(f"[tower {arch} B{B} mb{micro_batch}] {ms:.2f} ms  {B / ms * 1e3:.0f} img/s  {flops / ms / 1e9:.0f} TFLOP/s", flush=True)


def case_preprocess(h=480, w=640, n_px=224, count=256):
    """pc_preprocess_image (the reference's `_transform`) per image, images resident on the device, next to the reference's
    host pipeline (PIL + torchvision) on one core."""
    import numpy as np
    import torch
    from proto_clip_b200 import _native as nat
    from proto_clip_b200 import clip
    rng = np.random.default_rng(0)
    imgs = [torch.from_numpy((rng.random((h, w, 3)) * 255).astype(np.uint8)).cuda() for _ in range(8)]
    batch = torch.empty(count, 3, n_px, n_px, device="cuda")
    for i in range(8):
        nat.preprocess_image(imgs[i % 8], n_px, out=batch[i])
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(count):
        nat.preprocess_image(imgs[i % 8], n_px, out=batch[i])
    t1.record(); torch.cuda.synchronize()
    us = t0.elapsed_time(t1) * 1e3 / count
    bytes_alg = h * w * 3 + 3 * n_px * n_px * 4
    stack = torch.stack([imgs[i % 8] for i in range(count)])
    for _ in range(3):
        nat.preprocess_image(stack, n_px, out=batch)
    t0.record()
    for _ in range(10):
        nat.preprocess_image(stack, n_px, out=batch)
    t1.record(); torch.cuda.synchronize()
    us_b = t0.elapsed_time(t1) * 1e3 / 10 / count
    print(f"[preprocess {h}x{w} -> {n_px}] batched ({count} images / call): {us_b:.2f} us / image ({1e6 / us_b:.0f} img/s, "
          f"{bytes_alg / us_b / 1e3:.0f} GB/s algorithmic)", flush=True)
    from PIL import Image
    tf = clip.clip._transform(n_px)
    pil = [Image.fromarray(im.cpu().numpy()) for im in imgs]
    t = time.time()
    for i in range(64):
        tf(pil[i % 8])
    cpu_us = (time.time() - t) * 1e6 / 64
    print(f"[preprocess {h}x{w} -> {n_px}] GPU {us:.1f} us / image ({1e6 / us:.0f} img/s, {bytes_alg / us / 1e3:.1f} GB/s algorithmic)"
          f"  |  host PIL + torchvision {cpu_us:.0f} us / image on one core ({1e6 / cpu_us:.0f} img/s)", flush=True)


CASES = {
    "preprocess_vga": lambda: case_preprocess(480, 640, 224),
    "preprocess_small": lambda: case_preprocess(96, 120, 224),
    "preprocess_hd": lambda: case_preprocess(1080, 1920, 336),
    "tower_rn50x16": lambda: case_tower("RN50x16", 128),
    "tower_rn50x16_mb16": lambda: case_tower("RN50x16", 128, 16),
    "tower_rn50x16_mb64": lambda: case_tower("RN50x16", 128, 64),
    "tower_rn50": lambda: case_tower("RN50", 512),
    "tower_vitb16": lambda: case_tower("ViT-B/16", 960),
    "block_parts": case_block_parts,
    "block_parts_512": lambda: case_block_parts(512),
    "gemm_small": lambda: case_gemm(128, 256, 64),
    "gemm_k": lambda: case_gemm(128, 256, 768),
    "gemm_n128": lambda: case_gemm(300, 128, 512),
    "gemm_ragged": lambda: case_gemm(1000, 2000, 512, "f32"),
    "gemm_pair_small": lambda: case_gemm(256, 256, 64),
    "gemm_pair_k": lambda: case_gemm(512, 512, 768),
    "gemm_pair_ragged": lambda: case_gemm(1000, 1000, 512, "f32"),
    "gemm_pair_tail": lambda: case_gemm(257, 768, 592, "res"),
    "gemm_qkv": lambda: case_gemm(96 * 197, 2304, 768),
    "gemm_out": lambda: case_gemm(96 * 197, 768, 768, "res"),
    "gemm_gelu": lambda: case_gemm(96 * 197, 3072, 768, "gelu"),
    "gemm_res": lambda: case_gemm(96 * 197, 768, 3072, "res"),
    "attn_197": lambda: case_attn(4, 197, 12, False),
    "attn_77c": lambda: case_attn(5, 77, 8, True),
    "attn_50": lambda: case_attn(3, 50, 12, False),
    "attn_257": lambda: case_attn(2, 257, 16, False),
    "attn_big": lambda: case_attn(96, 197, 12, False),
    "attn_512": lambda: case_attn(512, 197, 12, False),
    "attn_577": lambda: case_attn(2, 577, 16, False),
    "attn_256_big": lambda: case_attn(192, 256, 16, False),
    "attn_257_big": lambda: case_attn(192, 257, 16, False),
    "attn_577_big": lambda: case_attn(64, 577, 16, False),
    "attn_text_big": lambda: case_attn(192, 77, 8, True),
    "attn_L14_big": lambda: case_attn(48, 257, 16, False),
    "attn_336_big": lambda: case_attn(16, 577, 16, False),
    "rows": case_rows,
    "cublas": case_cublas,
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--timeout", type=int, default=120)
    ap.add_argument("--only", default=None, help="substring filter on case names")
    a = ap.parse_args()
    if a.case:
        CASES[a.case]()
    else:
        for name in CASES:
            if a.only and not any(o in name for o in a.only.split(",")):
                continue
            t = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, "--case", name], timeout=a.timeout,
                                   capture_output=True, text=True)
                out = (r.stdout + r.stderr[-1500:]) if r.returncode else r.stdout
                print(f"=== {name}: rc={r.returncode} ({time.time() - t:.1f}s)\n{out}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"=== {name}: TIMEOUT after {a.timeout}s", flush=True)
