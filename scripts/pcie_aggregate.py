"""Aggregate host -> device bandwidth of the box when N ranks (one per GPU, torchrun) copy at the same time from pinned
memory: what bounds the end-to-end step at N > 1 (every rank's shard of the input crosses the same host).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_aggregate.py"""
import os, time, json
import torch
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 27   # 512 MiB of float32 per copy
h = torch.empty(n, dtype=torch.float32).pin_memory()
h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
d2 = torch.empty(n, dtype=torch.float32, device="cuda")
s2 = torch.cuda.Stream()
def run(reps, duplex):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
        if duplex:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
out = {}
for duplex in (False, True):
    run(2, duplex)
    dt = run(8, duplex)
    gbs = torch.tensor([n * 4 / 1e9 / dt], device="cuda", dtype=torch.float64)
    allv = [torch.zeros_like(gbs) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, gbs)
    else:
        allv = [gbs]
    out["h2d_with_d2h" if duplex else "h2d_alone"] = {"per_rank_GBps": [round(float(v.item()), 1) for v in allv],
                                                      "aggregate_GBps": round(float(sum(v.item() for v in allv)), 1)}
if rank == 0:
    print(json.dumps({"ranks": world, "cpus": len(os.sched_getaffinity(0)), **out}))
if world > 1:
    dist.destroy_process_group()
