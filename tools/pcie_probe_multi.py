"""Host<->device bandwidth ceiling of a multi-GPU box: n_active = 1, 2, 4, ... ranks copy at the same time (pinned memory, one
direction and both directions at once), the other ranks wait at the barrier.  The e2e leg of bench.py at N GPUs cannot beat
    bytes_per_rank / (duplex GB/s per rank with N ranks active).
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py
Prints one JSON line per configuration on rank 0 (append to profiles/)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

MB = int(os.environ.get("PROBE_MB", "1024"))
n = MB * (1 << 20) // 4
hx = torch.empty(n, dtype=torch.float32, pin_memory=True); hx.fill_(1.0)
hy = torch.empty(n, dtype=torch.float32, pin_memory=True); hy.fill_(0.0)
dx = torch.empty(n, dtype=torch.float32, device=dev)
dy = torch.ones(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
gb = n * 4 / 1e9


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def h2d():
    with torch.cuda.stream(s1):
        dx.copy_(hx, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hy.copy_(dy, non_blocking=True)


def both():
    h2d(); d2h()


def run(fn, subset, reps=2):
    """the ranks of `subset` run fn at the same time; returns this rank's best time (inf when idle)"""
    best = float("inf")
    for i in range(reps + 1):
        barrier()
        t0 = time.perf_counter()
        if rank in subset:
            fn()
            torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if i > 0 and rank in subset:
            best = min(best, dt)
    return best


# which GPUs copy at the same time: PROBE_SUBSETS="0;0,1;0,4;..." (default: the first 1, 2, 4, 8 and strided picks, to expose GPUs
# that share a PCIe switch uplink)
default = ["0", "0,1", "0,2", "0,4", "0,1,2,3", "0,2,4,6", "0,1,4,5", "0,1,2,3,4,5,6,7"]
subsets = [sorted({int(x) for x in t.split(",")}) for t in os.environ.get("PROBE_SUBSETS", ";".join(default)).split(";")]
subsets = [sub for sub in subsets if max(sub) < world]
for subset in subsets:
    active = len(subset)
    row = {"gpus": subset, "n_active": active, "MB_per_rank_per_direction": MB, "world": world}
    for name, fn, k in (("h2d", h2d, 1), ("d2h", d2h, 1), ("duplex", both, 1)):
        t = run(fn, subset)
        tt = torch.tensor([t if t != float("inf") else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tmax = float(tt.item())
        row[name + "_GBps_per_rank_each_direction"] = round(gb / tmax, 2)
        row[name + "_GBps_aggregate_each_direction"] = round(active * gb / tmax, 2)
    if rank == 0:
        row["host"] = {"cpus": len(os.sched_getaffinity(0)), "omp": os.environ.get("OMP_NUM_THREADS")}
        print(json.dumps(row), flush=True)
if world > 1:
    dist.destroy_process_group()
