"""Multi-GPU plumbing: reads shard embarrassingly across devices (one engine replica per GPU, the reference's
`api::create_basecall_runners` builds one CudaCaller per device, dorado/api/runner_creation.cpp:91-113), so the
only cross-rank operations are a barrier and scalar reductions for timing/statistics.  No data-path collective.

Works with any initialised torch.distributed backend (nccl on GPUs, gloo in the CPU tests); degrades to the
single-process identity when no process group exists."""
from __future__ import annotations

from typing import List


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _device():
    import torch
    d = _dist()
    if d is not None and d.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def shard_reads(num_reads: int, world: int, rank: int) -> List[int]:
    """Round-robin read ids for this rank (reads are independent; chunks of one read stay on one rank so that
    stitching never crosses ranks)."""
    return list(range(rank, num_reads, world))


def barrier() -> None:
    d = _dist()
    if d is not None:
        d.barrier()


def max_over_ranks(x: float) -> float:
    d = _dist()
    if d is None:
        return float(x)
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=_device())
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float) -> float:
    d = _dist()
    if d is None:
        return float(x)
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=_device())
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return float(t.item())


def all_gather_int(x: int) -> List[int]:
    d = _dist()
    if d is None:
        return [int(x)]
    import torch
    out = [torch.zeros(1, dtype=torch.int64, device=_device()) for _ in range(d.get_world_size())]
    d.all_gather(out, torch.tensor([x], dtype=torch.int64, device=_device()))
    return [int(t.item()) for t in out]
