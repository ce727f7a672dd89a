"""Row / series / chain sharding for one-process-per-GPU runs (torch.distributed).

The path shards embarrassingly: trajectories and chains are independent; series of
one chain are independent until the final sum of mcmc.py:486-491.  The ONLY collective is the
all-reduce(sum) of the per-rank [chains, 1+P+n_obs] log-likelihood/gradient block when series
are sharded (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def world() -> Tuple[int, int]:
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n` items owned by `rank` (sizes differ by at most 1)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t):
    """In-place sum over ranks of the log-likelihood block; no-op for a single process."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
