"""Where the time of ONE chromosome group of the e2e pipeline goes when nothing else runs on the GPU: upload, detect,
segments, call table, PMD segments, rule engine, region tables - host wall per call, kernel time and launches."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from methylasso_b200 import api
from methylasso_b200._native import MlParams

tables, ptr, cols = bench.make_workload(28_000_000)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 6
jobs, gptr, gcols = api.split_jobs(ptr, cols, G)[G - 1]
gcols = {k: np.ascontiguousarray(v) for k, v in gcols.items()}
eng = api.Engine(0)
eng.set("want_mu", 0)
p = MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
print("group of", len(jobs), "chromosomes,", int(gptr[-1]), "rows")


def once(profile):
    marks = []
    def mark(name, t0):
        eng.synchronize()
        marks.append((name, (time.perf_counter() - t0) * 1e3))
    t0 = time.perf_counter(); bt = eng.upload(gptr, gcols); mark("upload", t0)
    t0 = time.perf_counter(); res = eng.detect(bt, p); mark("detect", t0)
    t0 = time.perf_counter(); res.segments_count(0.01); mark("segments", t0)
    t0 = time.perf_counter(); res.call_table(0.01, 500.0, fetch=False); mark("call_table", t0)
    t0 = time.perf_counter(); res.call_pmd_segments(0.01, 500.0, 1000.0, fetch=False); mark("pmd_segments", t0)
    t0 = time.perf_counter(); res.call_lmr_umr(0.01, 500.0); mark("lmr_umr (+fetch)", t0)
    t0 = time.perf_counter(); res.call_pmd_split(0.01, 500.0, 1000.0); mark("valleys+split (+fetch)", t0)
    t0 = time.perf_counter(); res.call_pmds(0.01, 500.0); mark("pmds (+fetch)", t0)
    t0 = time.perf_counter(); res.call_pmd_status(0.01, 500.0); mark("status (+fetch)", t0)
    t0 = time.perf_counter(); res.segment_methylation_tables(0.01, 500.0); mark("region tables", t0)
    res.free(); bt.free()
    return marks

for _ in range(3):
    once(False)
m = [once(False) for _ in range(5)]
for i, (name, _) in enumerate(m[0]):
    print(f"{name:26s} {np.median([x[i][1] for x in m]):8.3f} ms")
# one call as the pipeline makes it
def fused():
    t0 = time.perf_counter(); bt = eng.upload(gptr, gcols); t1 = time.perf_counter()
    res = eng.detect(bt, p); t2 = time.perf_counter()
    res.segment_methylation_tables(0.01, 500.0, reuse_pinned=True); t3 = time.perf_counter()
    res.free(); bt.free()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3
for _ in range(3): fused()
f = np.median([fused() for _ in range(7)], axis=0)
print("as the pipeline calls it: upload %.2f  detect %.2f  segment_methylation_tables %.2f ms" % tuple(f))
eng.set("profile", 1); eng.profile_reset(); l0 = eng.launch_count
bt = eng.upload(gptr, gcols); res = eng.detect(bt, p); l1 = eng.launch_count
res.segment_methylation_tables(0.01, 500.0, reuse_pinned=True); eng.synchronize(); l2 = eng.launch_count
prof = eng.profile()
print("launches: detect", l1 - l0, "tables", l2 - l1, " kernel ms total %.3f" % sum(v[0] for v in prof.values()))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"    {k:26s} {v[0]:8.3f} ms  x{v[1]}")
