"""Batch sharding across the GPUs of one box.

Every MPC instance is independent, so the batch is cut into contiguous slices, one per rank
(one process per GPU), and NO collective runs inside the solve.  The only communication is an
optional gather of the per-instance results to rank 0, at most once per batch
(`torch.distributed.gather`: NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(B, rank, world):
    """Contiguous slice [lo, hi) of a batch of B owned by `rank`: ceil(B / world) per rank, the
    last ranks may own fewer (or none)."""
    per = -(-B // world)
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def shard_rows(a, rank, world):
    """Slice of an instance-major array [B, ...]."""
    lo, hi = shard_range(a.shape[0], rank, world)
    return a[lo:hi]


def shard_soa(a, rank, world):
    """Slice of a structure-of-arrays buffer [F, B] (contiguous copy: the kernels index [f*B + b])."""
    lo, hi = shard_range(a.shape[1], rank, world)
    return np.ascontiguousarray(a[:, lo:hi])


def gather_to_rank0(local, B_total, group=None):
    """Gathers instance-major per-rank results [B_local, F] to rank 0 -> [B_total, F] (None elsewhere).
    Ranks may own different counts: slices are padded to ceil(B_total / world) for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group); rank = dist.get_rank(group)
    per = -(-B_total // world)
    F = local.shape[1:]
    pad = torch.zeros((per,) + tuple(F), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    out = torch.cat(bufs, dim=0)[:per * world]
    keep = []
    for r in range(world):
        lo, hi = shard_range(B_total, r, world)
        keep.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(keep, dim=0)
