"""Multi-GPU plumbing: one process per GPU, query images sharded, ONE broadcast of the head state.

The reference is single-GPU (README.md:44 `CUDA_VISIBLE_DEVICES`); query images are independent, so the
path shards with no data-path collective: every rank loads the same CLIP weights, rank 0 builds
the support/text memory and broadcasts one packed buffer (prototypes + squared norms + adapter weights,
~2.3 MB for ImageNet ViT-B/16) over NCCL / NVLink, then each rank classifies its contiguous query slice.
Backend "nccl" on GPUs; the same code runs on "gloo" for the CPU tests of the host logic.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when not launched by it."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Join the torchrun job (no-op for a single process). Backend: `backend`, else $PROTOCLIP_DIST_BACKEND, else
    "nccl" with CUDA and "gloo" without. "gloo" also serves GPU runs whose ranks share one device (the N-rank ==
    1-rank test on a single-GPU box): collectives then go through host copies (see _host_staged)."""
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = os.environ.get("PROTOCLIP_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def active() -> bool:
    return dist.is_initialized() and dist.get_world_size() > 1


def is_main() -> bool:
    return not dist.is_initialized() or dist.get_rank() == 0


def device_for(local_rank: int) -> torch.device:
    """cuda:<local_rank>, folded onto the visible devices when several ranks share a GPU (gloo test runs)."""
    n = torch.cuda.device_count()
    return torch.device("cuda", local_rank % n) if n else torch.device("cpu")


def _host_staged(t: torch.Tensor) -> bool:
    """gloo has no CUDA all-gather: stage CUDA tensors through the host when the job runs on it."""
    return t.is_cuda and dist.get_backend() == "gloo"


def shard_bounds(num_queries: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of rank `rank`: ceil(Q / R) queries per rank, so concatenating the ranks'
    results in rank order preserves the reference's query order."""
    per = (num_queries + world - 1) // world
    lo = min(rank * per, num_queries)
    return lo, min(lo + per, num_queries)


def broadcast_flat(flat: Optional[torch.Tensor], numel: int, dtype: torch.dtype, device: torch.device,
                   src: int = 0) -> torch.Tensor:
    """The single collective of the path: rank `src` passes the packed head buffer, everyone gets it."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert flat is not None
        return flat
    if dist.get_rank() == src:
        assert flat is not None and flat.numel() == numel and flat.dtype == dtype
        buf = flat.to(device).contiguous()
    else:
        buf = torch.empty(numel, dtype=dtype, device=device)
    if _host_staged(buf):
        host = buf.cpu()
        dist.broadcast(host, src=src)
        return host.to(device)
    dist.broadcast(buf, src=src)
    return buf


def broadcast_tensors(tensors, src: int = 0):
    """ONE broadcast for a list of same-dtype tensors (shapes known on every rank): packed flat on `src`, unpacked
    as views of the received buffer everywhere. Used by main.py for the prototype memory."""
    if not active():
        return list(tensors)
    dtype, device = tensors[0].dtype, tensors[0].device
    numel = sum(t.numel() for t in tensors)
    flat = torch.cat([t.reshape(-1) for t in tensors]) if dist.get_rank() == src else None
    buf = broadcast_flat(flat, numel, dtype, device, src)
    out, off = [], 0
    for t in tensors:
        out.append(buf[off:off + t.numel()].view(t.shape))
        off += t.numel()
    return out


def gather_predictions(local: torch.Tensor, num_queries: int) -> Optional[torch.Tensor]:
    """Collect per-rank argmax slices on rank 0 in query order (outside the timed hot path)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (num_queries + world - 1) // world
    device = local.device
    if _host_staged(local):
        local = local.cpu()
    padded = torch.full((per,), -1, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, out, dst=0)
    if rank != 0:
        return None
    out = [o.to(device) for o in out]
    return torch.cat(out)[:num_queries] if per * world == num_queries else torch.cat(
        [o[: shard_bounds(num_queries, r, world)[1] - shard_bounds(num_queries, r, world)[0]] for r, o in enumerate(out)])


def all_gather_rows(local: torch.Tensor, total_rows: int) -> torch.Tensor:
    """Every rank holds the rows [shard_bounds(total_rows, rank, world)) of a [total_rows, ...] tensor; returns the
    full tensor on every rank in row order. ONE all-gather (rows padded to ceil(total / world) per rank). This is
    the collective of the sharded memory-bank build: support images / prompts are independent."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert local.shape[0] == total_rows
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (total_rows + world - 1) // world
    lo, hi = shard_bounds(total_rows, rank, world)
    assert local.shape[0] == hi - lo, f"rank {rank} holds {local.shape[0]} rows, expected {hi - lo}"
    device = local.device
    if _host_staged(local):
        local = local.cpu()
    padded = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous())
    out = out.to(device)
    if per * world == total_rows:
        return out
    keep = []
    for r in range(world):
        a, b = shard_bounds(total_rows, r, world)
        keep.append(out[r * per: r * per + (b - a)])
    return torch.cat(keep)


def all_gather_varlen(local: torch.Tensor) -> torch.Tensor:
    """Concatenation, in rank order, of every rank's rows (any number of rows per rank, including none): the ranks
    hold contiguous BATCH ranges of a loader, whose row counts are not known in advance. Two collectives (the row
    counts, then the padded rows)."""
    if not active():
        return local
    world = dist.get_world_size()
    device = local.device
    staged = _host_staged(local)
    work = local.cpu() if staged else local
    shape = torch.tensor([work.shape[0]] + [int(x) for x in work.shape[1:]], dtype=torch.int64, device=work.device)
    shapes = [torch.empty_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape)
    rows = [int(sh[0]) for sh in shapes]
    tail = next((tuple(int(x) for x in sh[1:]) for sh in shapes if int(sh[0]) > 0), tuple(work.shape[1:]))
    per = max(rows) if rows else 0
    if per == 0:
        return local
    padded = torch.zeros((per,) + tail, dtype=work.dtype, device=work.device)
    if work.shape[0]:
        padded[: work.shape[0]] = work.reshape((work.shape[0],) + tail)
    out = torch.empty((world * per,) + tail, dtype=work.dtype, device=work.device)
    dist.all_gather_into_tensor(out, padded.contiguous())
    full = torch.cat([out[r * per: r * per + rows[r]] for r in range(world)])
    return full.to(device)


def all_reduce_sum(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over the ranks (integer hit counts of the sharded (alpha, beta) grid search)."""
    if not active():
        return t
    if _host_staged(t):
        host = t.cpu()
        dist.all_reduce(host, op=dist.ReduceOp.SUM)
        t.copy_(host)
        return t
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def sharded_batches(loader):
    """This rank's contiguous range of the loader's batches (all of them for a single process). Loaders that can
    start anywhere expose `iter_range(lo, hi)`; any other iterable is walked and the foreign batches skipped."""
    if not active():
        yield from loader
        return
    lo, hi = shard_bounds(len(loader), dist.get_rank(), dist.get_world_size())
    if hasattr(loader, "iter_range"):
        yield from loader.iter_range(lo, hi)
        return
    for i, batch in enumerate(loader):
        if i >= hi:
            break
        if i >= lo:
            yield batch


def max_over_ranks(value: float, device: torch.device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device="cpu" if dist.get_backend() == "gloo" else device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
