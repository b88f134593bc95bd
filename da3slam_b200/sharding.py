"""Multi-GPU sharding of independent submap pairs.

One process per GPU (torchrun); pairs — consecutive submap pairs and loop-closure candidates
alike — are independent units, so each rank aligns a contiguous block of the pair list and the
ONLY exchange is one all_gather of the [n_local, 16] float64 Sim(3) rows (128 B per pair,
latency bound; NCCL over NVLink on GPUs, gloo on CPU for the tests).  The chain accumulation
(utils/geometry.py:73-119) then runs on every rank on the gathered table.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

ROW_LEN = 16


def shard_range(n_units: int, rank: int, world: int):
    """Contiguous block partition: the first (n_units % world) ranks get one extra unit."""
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_units: int, world: int):
    return [shard_range(n_units, r, world)[1] - shard_range(n_units, r, world)[0] for r in range(world)]


def gather_rows(rows_local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all_gather the per-rank row blocks into the full [n_total, 16] table (same on every rank).
    Blocks are padded to the largest shard so that a single fixed-size collective is used."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        assert rows_local.shape[0] == n_total
        return rows_local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    pad = max(sizes)
    buf = torch.zeros((pad, ROW_LEN), dtype=torch.float64, device=rows_local.device)
    buf[: rows_local.shape[0]] = rows_local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)
