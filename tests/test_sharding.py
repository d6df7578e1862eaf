"""Multi-GPU path on CPU: world_size-2 (and 3) gloo process groups run the sharded batch -- each
rank solves its contiguous slice (here with the CPU oracle standing in for the device kernels;
the host-side logic under test is the sharding and the single gather to rank 0) -- and rank 0
must hold exactly the unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quadrupedal_loco_b200 import sharding, synth


def test_shard_ranges_partition_the_batch():
    for B in (0, 1, 7, 8, 4096, 65537):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            for a, b in zip(r, r[1:]):
                assert a[1] == b[0]
            assert all(hi - lo <= -(-B // world) for lo, hi in r)
    a = np.arange(20).reshape(2, 10)
    assert np.array_equal(np.concatenate([sharding.shard_soa(a, k, 3) for k in range(3)], axis=1), a)


def _worker(rank, world, port, B, nh, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import oracle_lib
    orc = oracle_lib.Oracle()
    d = synth.body_mpc_inputs(B, nh, seed=5)
    lo, hi = sharding.shard_range(B, rank, world)
    cfg = orc.body_cfg(nh)
    theta = d["theta"][lo:hi].copy(); x = d["x_warm"][lo:hi].copy(); o14 = np.zeros((hi - lo, 14))
    r = orc.body_step_batch(cfg, d["tick"][lo:hi], np.ascontiguousarray(d["tx"][lo:hi]), theta,
                            np.ascontiguousarray(d["bstate"][lo:hi]), np.ascontiguousarray(d["refs"][lo:hi]), o14, x)
    local = torch.from_numpy(np.concatenate([o14, theta, x, r["status"][:, None].astype(float)], axis=1))
    full = sharding.gather_to_rank0(local, B)
    # the preallocated once-per-batch gather bench.py uses must give the same rows
    per = -(-B // world)
    rg = sharding.ResultGather(per, local.shape[1], local.dtype, local.device)
    rg.local.zero_(); rg.local[:hi - lo] = local
    rg.gather()
    if rank == 0:
        assert torch.equal(rg.rows(B), full)
        ret.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 101), (3, 64)])
def test_sharded_batch_equals_unsharded(oracle, world, B):
    nh = 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, nh, ret)) for r in range(world)]
    for p in procs:
        p.start()
    full = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = synth.body_mpc_inputs(B, nh, seed=5)
    cfg = oracle.body_cfg(nh)
    theta = d["theta"].copy(); x = d["x_warm"].copy(); o14 = np.zeros((B, 14))
    r = oracle.body_step_batch(cfg, d["tick"], d["tx"], theta, d["bstate"], d["refs"], o14, x)
    want = np.concatenate([o14, theta, x, r["status"][:, None].astype(float)], axis=1)
    np.testing.assert_array_equal(full, want)
