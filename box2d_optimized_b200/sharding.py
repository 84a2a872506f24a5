"""Multi-GPU plumbing for the batched-independent-worlds case: worlds are sharded
round-robin over ranks, one process per GPU, with NO data-path collective — the only
communication is the timing reduction at the end (barrier + max over ranks).  Backend-agnostic
(`nccl` on the GPU box, `gloo` in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_worlds(num_worlds, rank, world_size):
    """world w -> rank (w mod world_size); returns this rank's world ids"""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, num_worlds, world_size))


def aggregate(local_units, local_ms, device="cpu"):
    """whole-job throughput: sum of units over ranks / max elapsed over ranks.
    Returns (units_total, ms_max, units_per_second)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        u = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
        t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
        dist.barrier()
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        units, ms = float(u.item()), float(t.item())
    else:
        units, ms = float(local_units), float(local_ms)
    return units, ms, units / (ms / 1000.0) if ms > 0 else 0.0
