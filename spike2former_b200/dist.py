"""Multi-GPU plumbing for batch-sharded inference: one process per GPU, replicated weights,
no collective on the data path.  The only communication is the timing protocol of bench.py (barrier + max over ranks)
and the result gather of an evaluation run.  Works with the `nccl` backend on GPUs and `gloo` on CPU (tests)."""
from __future__ import annotations

import os

import torch


def env_world():
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not launched by torchrun."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(total: int, world: int, rank: int):
    """Contiguous, balanced [begin, end) of `total` images for `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device="cpu") -> float:
    """Step time of the job = the slowest rank's (bench.py contract).  No-op without a process group."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_label_maps(labels: torch.Tensor, total: int):
    """All-gather per-rank label maps [b_r, H, W] (uint8) of a sharded evaluation into [total, H, W] on every rank."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return labels
    world = dist.get_world_size()
    sizes = [shard_range(total, world, r) for r in range(world)]
    biggest = max(e - b for b, e in sizes)
    pad = torch.zeros((biggest,) + tuple(labels.shape[1:]), dtype=labels.dtype, device=labels.device)
    pad[: labels.shape[0]] = labels
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: e - b] for o, (b, e) in zip(out, sizes)], 0)
