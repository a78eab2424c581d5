"""What can this box's host take?  N ranks (one per GPU, torchrun) doing nothing but device -> pinned-host copies.

The end-to-end path returns 4 bytes per decoded pixel over PCIe (the Int32Array contract of IVideoCodec.hx:11-29), so at N GPUs
its ceiling is whatever N concurrent D2H streams reach into ONE host's memory -- a property of the box (PCIe roots, NUMA,
IOMMU / hypervisor), not of the decoder.  This measures it, first with the rank's thread and pinned buffer wherever the OS put
them ("unbound"), then after jsp_numa_bind_thread(device) ("bound": thread + pinned pages on the GPU's NUMA node).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/d2h_ceiling.py
prints one JSON line per mode on rank 0; profiles/r02_d2h_ceiling.json keeps the numbers bench.py quotes as e2e.host_ceiling_gbs.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jsplayer_b200 import _lib  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
GB = float(os.environ.get("D2H_GB", "2"))
REPS = int(os.environ.get("D2H_REPS", "6"))

torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _lib.load()
n = int(GB * (1 << 30))
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
dev.fill_(rank + 1)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(mode):
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.copy_(dev, non_blocking=True)                  # touch every page once
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(REPS):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    t_all = time.perf_counter() - t0                    # until the slowest rank is done
    mine = torch.tensor([n * REPS / dt / 1e9], dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    tmax = torch.tensor([t_all], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    assert int(host[12345]) == rank + 1
    if rank == 0:
        print(json.dumps({"mode": mode, "n_gpus": world, "gb_per_copy": GB, "copies": REPS,
                          "per_rank_gbs": [round(float(x.item()), 2) for x in allr],
                          "total_gbs": round(world * n * REPS / float(tmax.item()) / 1e9, 2),
                          "numa_node_of_gpu0": lib.jsp_numa_node_of_device(0),
                          "cpus": os.cpu_count()}), flush=True)
    del host


run("unbound")
node = lib.jsp_numa_bind_thread(local)
run("bound to NUMA node %d" % node if node >= 0 else "bound (no NUMA topology visible: nothing changed)")
if world > 1:
    dist.destroy_process_group()
