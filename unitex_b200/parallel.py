"""Multi-GPU plumbing for the path: ranks shard INDEPENDENT grids/assets -- all views of one grid share one
attention sequence (flux_piplines/texturing/pipeline.py:630-656), so views are never split -- and the only exchange is one
all-gather of the finished tiles before UV projection.  torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_grids(n_grids: int, rank: int, world: int) -> List[int]:
    """Grid indices owned by `rank`: round-robin, so seeds 63, 64, ... (run.py:5) land on ranks 0, 1, ... ."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_grids, world))


def grid_seed(base_seed: int, grid_index: int) -> int:
    return base_seed + grid_index


def all_gather_tiles(tile: torch.Tensor) -> List[torch.Tensor]:
    """One collective per batch of assets: every rank receives every rank's finished tile (same shape/dtype)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [tile]
    out = [torch.empty_like(tile) for _ in range(dist.get_world_size())]
    dist.all_gather(out, tile.contiguous())
    return out


def max_over_ranks(values: Sequence[float], device) -> List[float]:
    """Timing convention of bench.py: a multi-GPU number is the max over ranks."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
