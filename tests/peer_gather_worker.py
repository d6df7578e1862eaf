"""Worker of tests/test_gpu_peer_gather.py (launched with torch.distributed.run, one process per GPU): every rank fills its
rows of each slot with a rank- and round-dependent pattern through a kernel writing to PeerGather.dest(), rank 0 checks the
gathered blocks over many rounds with slot reuse (acquire / publish / wait_all / release), ranks deliberately out of step."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quadrupedal_loco_b200 as q
from quadrupedal_loco_b200 import sharding


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = q.load_library()
    per, F, slots, rounds = 1000, 12, 3, 40
    pg = sharding.PeerGather(lib, local, per, F, slots)
    streams = [torch.cuda.Stream(device=dev) for _ in range(slots)]
    host = [torch.zeros(world, per, F, dtype=torch.float64).pin_memory() for _ in range(slots)]
    src = torch.arange(per * F, dtype=torch.float64, device=dev).view(per, F)
    bad = 0
    for k in range(rounds):
        s = k % slots
        st = streams[s]
        if rank == 1 and k % 7 == 3:
            time.sleep(0.02)                                  # a rank falling behind must not corrupt anything
        with torch.cuda.stream(st):
            pg.acquire(s, st.cuda_stream)
            dst = torch.as_tensor(sharding._DevView(pg.dest(s), (per, F)), device=dev)
            dst.copy_(src * (rank + 1) + 1000.0 * k)          # device kernel storing to rank 0's memory (NVLink on peers)
            pg.publish(s, st.cuda_stream)
            if rank == 0:
                pg.wait_all(s, st.cuda_stream)
                host[s].copy_(pg.block(s), non_blocking=True)
                pg.release(s, st.cuda_stream)
        if rank == 0 and k >= slots - 1:
            # check the round that used the slot we are about to reuse next
            kk = k
            streams[s].synchronize()
            want = np.stack([np.arange(per * F).reshape(per, F) * (r + 1) + 1000.0 * kk for r in range(world)])
            bad += int(not np.array_equal(host[s].numpy(), want))
    torch.cuda.synchronize()
    st_code = pg.status()
    dist.barrier()
    pg.close()
    if rank == 0:
        print(f"PEER_GATHER world={world} rounds={rounds} bad={bad} status={st_code}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if (bad == 0 and st_code == 0) else 1)


if __name__ == "__main__":
    main()
