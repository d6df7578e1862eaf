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
    dist.gather(pad, bufs, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None
    out = torch.cat(bufs, dim=0)[:per * world]
    keep = []
    for r in range(world):
        lo, hi = shard_range(B_total, r, world)
        keep.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(keep, dim=0)


class ResultGather:
    """Once-per-batch gather of the per-robot results to rank 0 with every buffer allocated up front (no allocation,
    no host synchronisation per call): rank r contributes `local` [per, F] (its shard, zero-padded to per = ceil(B / world)
    rows), rank 0 receives [world, per, F].  On GPUs the backend is NCCL: one grouped ncclSend / ncclRecv per peer over
    NVLink / NVSwitch (torch.distributed.gather), enqueued on the current CUDA stream -- no collective inside the solve,
    only this exchange of <= 0.5 KB per robot behind it."""

    def __init__(self, per, F, dtype, device, group=None):
        import torch
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group); self.rank = dist.get_rank(group)
        self.dst = dist.get_global_rank(group, 0) if group is not None else 0
        self.per, self.F = per, F
        self.local = torch.zeros((per, F), dtype=dtype, device=device)
        self.all = torch.zeros((self.world, per, F), dtype=dtype, device=device) if self.rank == 0 else None
        self._views = [self.all[r] for r in range(self.world)] if self.rank == 0 else None

    def gather(self):
        """Enqueue the gather of `self.local`; rank 0 finds the result in `self.all` (row block r = rank r's shard)."""
        self.dist.gather(self.local, self._views, dst=self.dst, group=self.group)
        return self.all

    def rows(self, B_total):
        """rank 0: the gathered rows in batch order, padding removed -> [B_total, F]."""
        import torch
        keep = []
        for r in range(self.world):
            lo, hi = shard_range(B_total, r, self.world)
            keep.append(self.all[r, :hi - lo])
        return torch.cat(keep, dim=0)


def pin_to_gpu_numa_node(device_index):
    """Bind the calling process (all its present and future threads) to the CPU cores of the NUMA node the GPU hangs off, so
    that the pinned host buffers it allocates afterwards are first-touched in memory local to that GPU's PCIe root and its
    feeder threads run beside them.  Eight ranks streaming records over eight links otherwise meet on one socket's memory
    controllers.  Linux sysfs only (no libnuma); returns a description of what was done, or of why nothing was."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        addr = f"{dom:04x}:{bus:02x}:{dev:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{addr}/numa_node").read().strip())
        if node < 0:
            return f"GPU {device_index} ({addr}): no NUMA affinity reported"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return f"GPU {device_index} ({addr}): NUMA node {node} has no CPU this process may use"
        os.sched_setaffinity(0, allowed)
        return f"GPU {device_index} ({addr}): bound to NUMA node {node}, {len(allowed)} CPUs"
    except Exception as e:                      # containers without sysfs PCI entries, non-Linux hosts
        return f"not bound ({type(e).__name__}: {e})"
