"""Multi-GPU sharding of a batch: instances / frequency points are independent, so rank g owns the
contiguous range [g * ceil(B / G), min(B, (g + 1) * ceil(B / G))) and runs it on its own GPU with no data-path
collective; torch.distributed (NCCL on the GPU box, gloo in the CPU tests) is used only to gather results and to
reduce the counters at the end of a run."""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def shard_values(values: np.ndarray, rank: int, world: int) -> np.ndarray:
    """slice the last axis (instances) of a per-instance parameter table"""
    lo, hi = shard_range(values.shape[-1], rank, world)
    return np.ascontiguousarray(values[..., lo:hi])


def gather_instances(local: np.ndarray, n_total: int, dist=None, device=None) -> np.ndarray | None:
    """all ranks call with their [n_local, ...] result block; returns the [n_total, ...] array on every rank
    (all_gather of equally padded blocks).  dist = torch.distributed (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch

    world = dist.get_world_size()
    per = -(-n_total // world)
    pad = np.zeros((per,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = torch.cat(outs, dim=0)[:n_total]
    return full.cpu().numpy()


def reduce_counters(solves: int, failed: int, dist=None, device=None) -> tuple[int, int]:
    """sum of solve_once-equivalents and of failed lanes over all ranks"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(solves), int(failed)
    import torch

    t = torch.tensor([solves, failed], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t[0].item()), int(t[1].item())
