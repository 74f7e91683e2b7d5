"""Multi-GPU plumbing: one process per GPU, poses sharded by contiguous index ranges, no data-path collective.

The reference's only multi-GPU mechanism is single-process `torch.nn.DataParallel` (`MPL/run/valid_mpl.py:177-178`):
replicate + scatter + gather every step.  Every pose is independent through the whole network, so the B200-native
equivalent is a full weight replica per rank, a rank-sharded pose range, and one all-reduce of the metric sums.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """(rank, world_size, local_rank) from torchrun's environment; initialises the process group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(total: int, rank: int, world: int):
    """Contiguous slice [start, stop) of the global pose index owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_state(module: torch.nn.Module, src: int = 0):
    """Optional: make every rank's replica bit-identical to rank `src` (one broadcast per tensor at init)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t, src=src)
        # the lifter keeps a packed device copy of its weights: tell it they changed (a collective writes in place without
        # necessarily bumping Tensor._version)
        for m in module.modules():
            if hasattr(m, "invalidate"):
                m.invalidate()


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over ranks (timing rule: report the slowest rank)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
