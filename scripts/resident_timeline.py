"""Where the host gaps of the RESIDENT step are: every kernel of one engine on the timeline (ml_engine_profile_spans) for a
few steps of bench.py's step_resident on one GPU; prints the step's span, the sum of its kernels and the largest gaps with the
kernels on either side."""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from methylasso_b200 import api, shard
from methylasso_b200._native import MlParams

tables, ptr, cols = bench.make_workload(28_000_000)
eng = api.Engine(0)
eng.set("want_mu", 0)
params = MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
batch = eng.upload(ptr, cols)
job_ids = np.arange(len(tables), dtype=np.int32)
gather = shard.RegionGather(eng, 0, 1, 8 << 20, None, "cuda:0", "ipc")


def step():
    res = eng.detect(batch, params)
    gather.all_gather(res, job_ids, 0.01, 500.0, fetch=False)
    res.free()


for _ in range(4):
    step()
eng.set("pool_headroom_percent", 25)
eng.set("profile_reserve_events", 3000)
eng.set("profile", 1)
eng.set("profile_timeline", 1)
step()
out = []
for rep in range(3):
    eng.profile_reset()
    t0 = time.perf_counter()
    step()
    eng.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    sp = eng.profile_spans()
    span = sp[-1][2] - sp[0][1]
    ksum = sum(b - a for _, a, b in sp)
    gaps = sorted(((sp[i + 1][1] - sp[i][2], sp[i][0], sp[i + 1][0]) for i in range(len(sp) - 1)), reverse=True)
    out.append({"wall_ms": round(wall, 3), "first_to_last_kernel_ms": round(span, 3), "kernel_sum_ms": round(ksum, 3),
                "launches": len(sp), "gaps_over_20us": sum(1 for g in gaps if g[0] > 0.02),
                "sum_of_gaps_ms": round(sum(g[0] for g in gaps), 3),
                "largest_gaps": [(round(g, 3), a, b) for g, a, b in gaps[:14]]})
print(json.dumps(out, indent=1))
gather.close()
batch.free()
eng.close()
