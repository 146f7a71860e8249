"""Host-side logic of the multi-GPU path: which chains a rank owns and how per-rank results combine.

Chains never interact, so the path shards with no per-step collective: rank r owns a
contiguous range of GLOBAL chain ids (the Philox stream of a chain is keyed by its global id, so the
union of the shards is the single-GPU run), and the only exchanges are one sum of the 64-bit
model-visit histogram / counters and one max of the device time.  Works with any torch.distributed
backend (nccl on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


def shard_range(total_chains: int, world: int, rank: int) -> tuple[int, int]:
    """(first global chain id, count) for `rank`; ragged totals give the first ranks one more."""
    if world < 1 or not (0 <= rank < world) or total_chains < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total_chains, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def weak_range(chains_per_rank: int, rank: int) -> tuple[int, int]:
    """Weak scaling (bench.py): every rank owns the same number of chains."""
    return rank * chains_per_rank, chains_per_rank


def allreduce_sum_(t):
    """In-place sum over ranks of an integer/float tensor (histogram, counters, flops)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_max_(t):
    """In-place max over ranks (device time: every multi-GPU number is the max over ranks)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t
