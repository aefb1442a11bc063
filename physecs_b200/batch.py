"""Batched independent scenes across GPUs (BASELINE.json configs[4]).

Scenes never interact, so a batch shards by scene: rank r of G simulates a contiguous block of scenes in its own device
context (one process per GPU).  There is no collective on the data path; torch.distributed is used only for the barrier
around the timed region and the max-over-ranks reduction of the measured time (bench.py).
"""
from __future__ import annotations


def shard_range(n_scenes: int, world: int, rank: int):
    """Contiguous block [begin, end) of scene indices owned by `rank`; blocks differ in size by at most one scene."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world / rank")
    base, extra = divmod(n_scenes, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (a device time) over all ranks; identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def scene_seed(base_seed: int, scene_index: int) -> int:
    """Per-scene RNG seed: depends on the GLOBAL scene index, so a scene is the same whichever rank simulates it."""
    return (base_seed + 0x9E3779B1 * (scene_index + 1)) & 0xFFFFFFFF
