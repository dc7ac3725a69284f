"""Sharding of a batch of independent images across ranks / devices.

Every image is independent (each operation reads only the original raster,
internal/usecase/processor/image_processor.go:64-65), so the path shards by image with
no data-path collective: rank r of W processes images r, r+W, r+2W, ... of the batch
(the reference's analogue is Kafka partitioning across worker containers,
internal/worker/worker.go:88-96).  torch.distributed is used for the run's plumbing only:
a barrier around the timed region and the max / sum over ranks of the measurements.
Works on any backend (nccl on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Indices of the batch this rank owns: round-robin, so mixed-size streams balance."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_items, world))


def least_loaded(outstanding_bytes: Sequence[int]) -> int:
    """Device choice for a mixed-size stream inside one process (the engine applies the same
    rule to its devices, engine.cpp submit_impl): fewest outstanding bytes, lowest index wins ties."""
    best = 0
    for i, v in enumerate(outstanding_bytes):
        if v < outstanding_bytes[best]:
            best = i
    return best


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def reduce_max(v: float) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(v)
    t = torch.tensor([v], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(v: float) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(v)
    t = torch.tensor([v], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_throughput(local_units: float, local_seconds: float) -> float:
    """Units all ranks processed / the slowest rank's time (the bench contract's `value`)."""
    return reduce_sum(local_units) / max(reduce_max(local_seconds), 1e-12)


def gather_checksums(local: Sequence[int], n_items: int, rank: int, world: int) -> List[int]:
    """Per-image checksums of every rank, put back in batch order (on every rank)."""
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return list(local)
    per = (n_items + world - 1) // world
    mine = torch.full((per,), -1, dtype=torch.int64, device=_device())
    if local:
        mine[:len(local)] = torch.tensor(list(local), dtype=torch.int64, device=_device())
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = [0] * n_items
    for r in range(world):
        for k, i in enumerate(shard_indices(n_items, r, world)):
            out[i] = int(parts[r][k].item())
    return out
