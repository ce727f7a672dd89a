"""Row / series / chain sharding for one-process-per-GPU runs (torch.distributed).

The path shards embarrassingly: trajectories and chains are independent; series of
one chain are independent until the final sum of mcmc.py:486-491.  The ONLY collective is the
all-reduce(sum) of the per-rank [chains, 1+P+n_obs] log-likelihood/gradient block when series
are sharded: on the GPUs of one node the engine's own one-shot kernel over NVLink peer memory
(``PeerAllReduce`` -> ctx_p2p_allreduce), NCCL otherwise; gloo in the CPU tests."""
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


class PeerAllReduce:
    """One-shot all-reduce(sum) over NVLink peer memory for latency-bound blocks (ctx_p2p_*).

    Every rank of the node allocates a symmetric buffer, the 64-byte cudaIpc handles are exchanged
    once through torch.distributed (plumbing), and each call is ONE kernel of the engine: copy,
    flag exchange, P2P loads of every peer's block summed in rank order.  Stream-ordered on the
    current torch stream and CUDA-graph capturable.  ``max_bytes``: largest message."""

    def __init__(self, max_bytes: int, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import engine

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerAllReduce needs an initialised torch.distributed process group")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._lib = engine.lib()
        self._h = C.c_void_p()
        handle = C.create_string_buffer(64)
        rc = self._lib.ctx_p2p_create(self.rank, self.world, int(max_bytes), C.byref(self._h), handle)
        # every rank takes part in every exchange and all ranks fail TOGETHER: a rank that left
        # early would leave the others waiting in a collective
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (rc, self._error() if rc else "", bytes(handle.raw)), group=group)
        bad = [(r, g[1]) for r, g in enumerate(gathered) if g[0] != 0]
        if bad:
            self.close()
            raise RuntimeError(f"ctx_p2p_create failed on rank(s) {bad}")
        rc = self._lib.ctx_p2p_connect(self._h, b"".join(g[2] for g in gathered))
        oks = [None] * self.world
        dist.all_gather_object(oks, (rc, self._error() if rc else ""), group=group)
        bad = [(r, g[1]) for r, g in enumerate(oks) if g[0] != 0]
        if bad:
            self.close()
            raise RuntimeError(f"ctx_p2p_connect failed on rank(s) {bad} (peer access over NVLink / "
                               f"cudaIpc needs the ranks to be GPUs of one node)")
        self.max_bytes = int(max_bytes)
        dist.barrier(group)        # every rank has mapped every buffer before the first call
        self._torch = torch

    def _error(self) -> str:
        import ctypes as C

        buf = C.create_string_buffer(4096)
        self._lib.ctx_p2p_last_error(buf, len(buf))
        return buf.value.decode(errors="replace")

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"ctx_p2p error {rc}: {self._error()}")

    def __call__(self, t):
        """In-place sum of the CUDA tensor ``t`` (float32 / float64, contiguous) over the ranks."""
        import ctypes as C

        torch = self._torch
        if not t.is_cuda or not t.is_contiguous() or t.dtype not in (torch.float32, torch.float64):
            raise ValueError("PeerAllReduce: contiguous float32 / float64 CUDA tensor expected")
        if t.numel() * t.element_size() > self.max_bytes:
            raise ValueError("PeerAllReduce: message larger than the symmetric buffer")
        stream = torch.cuda.current_stream(t.device).cuda_stream
        self._check(self._lib.ctx_p2p_allreduce(self._h, 1 if t.dtype == torch.float64 else 0, t.numel(),
                                                C.c_void_p(t.data_ptr()), C.c_void_p(stream)))
        return t

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ctx_p2p_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
