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
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


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
    dist.broadcast(buf, src=src)
    return buf


def gather_predictions(local: torch.Tensor, num_queries: int) -> Optional[torch.Tensor]:
    """Collect per-rank argmax slices on rank 0 in query order (outside the timed hot path)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (num_queries + world - 1) // world
    padded = torch.full((per,), -1, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, out, dst=0)
    if rank != 0:
        return None
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
    padded = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous())
    if per * world == total_rows:
        return out
    keep = []
    for r in range(world):
        a, b = shard_bounds(total_rows, r, world)
        keep.append(out[r * per: r * per + (b - a)])
    return torch.cat(keep)


def max_over_ranks(value: float, device: torch.device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
