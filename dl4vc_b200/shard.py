"""Candidate sharding for multi-GPU inference: candidates are independent in eval mode, so rank g of G
processes the contiguous index range [g*N/G, (g+1)*N/G) with its own weight replica and no data-path collective. The
only communication is the optional rank-ordered gather of the per-candidate score rows at the end (the reference
gathers DataParallel outputs on GPU 0, main.py:117 / trainer.py:629-630, then sorts the VCF by position anyway,
call_variants.sh:151)."""
from __future__ import annotations

import torch


def shard_range(num_candidates: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most one), order-preserving split of range(num_candidates)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(int(num_candidates), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_in_rank_order(local_rows: torch.Tensor, num_candidates: int, group=None) -> torch.Tensor | None:
    """Concatenate the per-rank result rows (rank r holds shard_range(num_candidates, r, world) rows) on rank 0.
    Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU test). Returns None on other ranks."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local_rows
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(num_candidates, rank, world)
    if local_rows.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local_rows.shape[0]} rows, its shard has {hi - lo}")
    width = local_rows.shape[1:]
    cap = -(-num_candidates // world)                      # every rank sends a block of the largest shard size
    block = local_rows.new_zeros((cap, *width))
    block[: hi - lo] = local_rows
    out = [torch.empty_like(block) for _ in range(world)] if rank == 0 else None
    dist.gather(block, out, dst=0, group=group)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        a, b = shard_range(num_candidates, r, world)
        parts.append(out[r][: b - a])
    return torch.cat(parts, dim=0)
