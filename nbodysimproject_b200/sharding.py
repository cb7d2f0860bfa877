"""Multi-GPU partitioning.

Ensembles shard by SYSTEM: rank r analyses the contiguous block shard_range(B, r, P) of every (N, mode) bucket with
no data-path collective; the only communication is the gather of the [B_r, K] feature tables at the end
(`gather_rows`, an all_gather over torch.distributed -- NCCL on GPUs, gloo in the CPU tests).  Results are
per-system and therefore identical whatever the rank count.

The large-N path shards by i-BLOCK: rank r owns particles shard_range(N, r, P) and needs one in-place position
all-gather per force evaluation (largen.py)."""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) block of `n` items for `rank` of `world` (first n % world ranks get one more)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local: np.ndarray, n_total: int, group=None) -> np.ndarray:
    """All-gather row blocks produced with shard_range back into the full [n_total, K] table (every rank gets it)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local = np.ascontiguousarray(local, dtype=np.float64)
    K = local.shape[1] if local.ndim == 2 else 1
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros((cap, K), dtype=torch.float64, device=dev)
    buf[: local.shape[0]] = torch.as_tensor(local.reshape(local.shape[0], K)).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    full = np.concatenate([out[r][: hi - lo].cpu().numpy() for r, (lo, hi) in enumerate(sizes)], axis=0)
    assert full.shape[0] == n_total
    return full


def analyze_sharded(compute: Callable[[int, int], np.ndarray], n_total: int, group=None) -> np.ndarray:
    """Run `compute(lo, hi) -> rows[hi-lo, K]` on this rank's shard and return the gathered full table."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_range(n_total, rank, world)
    return gather_rows(compute(lo, hi), n_total, group)
