"""Host wall time of each resident call of the step (debug aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from methylasso_b200 import api, synth
from methylasso_b200._native import MlParams
import bench

tables, ptr, cols = bench.make_workload(int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000)
eng = api.Engine(0)
bt = eng.upload(ptr, cols)
p = MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
for it in range(4):
    t = [time.perf_counter()]
    r = eng.detect(bt, p); eng.synchronize(); t.append(time.perf_counter())
    r.segments_count(0.01); eng.synchronize(); t.append(time.perf_counter())
    r.call_table(0.01, 500.0, fetch=False); eng.synchronize(); t.append(time.perf_counter())
    r.call_pmd_segments(0.01, 500.0, 1000.0, fetch=False); eng.synchronize(); t.append(time.perf_counter())
    r.pmd_segments(0.01, 1000.0, fetch=False); eng.synchronize(); t.append(time.perf_counter())
    r.free()
    print("detect %.2f  segments %.2f  call_table %.2f  call_pmd %.2f  old_pmd %.2f ms" % tuple((b - a) * 1e3 for a, b in zip(t, t[1:])), flush=True)
eng.set("profile", 1)
r = eng.detect(bt, p); r.segments_count(0.01); r.call_table(0.01, 500.0, fetch=False); r.call_pmd_segments(0.01, 500.0, 1000.0, fetch=False); eng.synchronize()
for k, v in sorted(eng.profile().items(), key=lambda kv: -kv[1][0])[:12]:
    print(k, round(v[0], 3), v[1])
