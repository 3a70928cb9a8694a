"""Host wall time of each API call of the resident bench step next to the kernel time inside it (per-kernel CUDA
events): where the host gaps of a step are."""
import gc, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from methylasso_b200 import api
from methylasso_b200._native import MlParams

eng = api.Engine(0)
stream = torch.cuda.current_stream()
eng.set_stream(stream.cuda_stream)
tables, ptr, cols = bench.make_workload(28_000_000)
batch = eng.upload(ptr, cols)
params = MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
calls = [("detect", lambda st: st.__setitem__("res", eng.detect(batch, params))),
         ("segments", lambda st: st["res"].segments_count(0.01)),
         ("call_table", lambda st: st["res"].call_table(0.01, 500.0, fetch=False)),
         ("call_pmd_segments", lambda st: st["res"].call_pmd_segments(0.01, 500.0, 1000.0, fetch=False)),
         ("call_lmr_umr (not in the bench step)", lambda st: st["res"].call_lmr_umr(0.01, 500.0)),
         ("call_valleys (not in the bench step)", lambda st: st["res"].call_valleys(0.01, 500.0)),
         ("free", lambda st: st["res"].free())]

def step(record=None, prof=False):
    st = {}
    for name, fn in calls:
        if prof:
            eng.profile_reset()
        t0 = time.perf_counter()
        fn(st)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if record is not None:
            k = sum(v[0] for v in eng.profile().values()) if prof else 0.0
            n = sum(v[1] for v in eng.profile().values()) if prof else 0
            record.setdefault(name, []).append((dt, k, n))

for _ in range(4):
    step()
eng.set("pool_headroom_percent", 25)
gc.collect(); gc.disable()
rec = {}
for _ in range(10):
    step(rec)
eng.set("profile", 1)
step(); step()
recp = {}
for _ in range(10):
    step(recp, prof=True)
print(f"{'call':38s} {'host wall ms':>12s} {'(with events)':>14s} {'kernel ms':>10s} {'launches':>9s} {'gap ms':>8s}")
tot = [0, 0, 0]
for name, _ in calls:
    w = np.median([x[0] for x in rec[name]]); wp = np.median([x[0] for x in recp[name]])
    k = np.median([x[1] for x in recp[name]]); n = np.median([x[2] for x in recp[name]])
    print(f"{name[:38]:38s} {w:12.2f} {wp:14.2f} {k:10.2f} {n:9.0f} {w - k:8.2f}")
    tot[0] += w; tot[1] += k; tot[2] += n
print(f"{'step':38s} {tot[0]:12.2f} {'':14s} {tot[1]:10.2f} {tot[2]:9.0f} {tot[0] - tot[1]:8.2f}")
