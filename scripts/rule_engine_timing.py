"""Host wall time and kernel time of the native rule engine of segment_methylation_singlechr (R/call.R:31-221) on the
resident fit of a config-2 genome: call table, LMR / UMR annotation, valleys, PMD candidate split, PMD calls, PMD status.
Every timed call is made with a new key (flank_dist_max toggled between 5000 and 5000.5: distances are integers, the
results do not change), so that it recomputes every step from the LMR / UMR annotation on: the rows are cumulative.  The
query-then-fetch pair inside one API call and the chain pmd_split -> pmds -> pmd_status on one key reuse the tables.

    python scripts/rule_engine_timing.py [cpgs] > gpurun_out/rule_engine_timing.log
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from methylasso_b200 import api
from methylasso_b200._native import MlParams

cpgs = int(sys.argv[1]) if len(sys.argv) > 1 else 28_000_000
tables, ptr, cols = bench.make_workload(cpgs)
eng = api.Engine(0)
batch = eng.upload(ptr, cols)
res = eng.detect(batch, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
res.call_pmd_segments(0.01, 500.0, 1000.0, fetch=False)      # call table and std_call resident: not timed below
eng.synchronize()
tick = [0]


def fresh():
    tick[0] += 1
    return dict(flank_dist_max=5000.0 + 0.5 * (tick[0] & 1))


calls = [("call_lmr_umr   :75-91", lambda: res.call_lmr_umr(0.01, 500.0, **fresh())),
         ("call_valleys   :75-133", lambda: res.call_valleys(0.01, 500.0, **fresh())),
         ("call_pmd_split :75-168", lambda: res.call_pmd_split(0.01, 500.0, 1000.0, **fresh())),
         ("call_pmds      :75-199", lambda: res.call_pmds(0.01, 500.0, **fresh())),
         ("call_pmd_status:75-215", lambda: res.call_pmd_status(0.01, 500.0, **fresh()))]
out = {}
print(f"{'entry (R/call.R lines it covers)':34s} {'host wall ms':>12s} {'kernel ms':>10s} {'launches':>9s}")
for name, fn in calls:
    for _ in range(2):
        out[name] = fn()
    eng.synchronize()
    t = []
    for _ in range(5):
        t0 = time.perf_counter(); fn(); eng.synchronize(); t.append((time.perf_counter() - t0) * 1e3)
    eng.set("profile", 1); eng.profile_reset(); fn(); eng.synchronize()
    prof = eng.profile(); eng.set("profile", 0)
    print(f"{name:34s} {np.median(t):12.2f} {sum(v[0] for v in prof.values()):10.3f} {sum(v[1] for v in prof.values()):9.0f}", flush=True)
    if name.startswith("call_pmd_status"):
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:10]:
            print(f"    {k:24s} {v[0]:8.3f} ms  x{v[1]}")
lu, val, (merged, pieces), pmd, fin = (out[n] for n, _ in calls)
print("call-table rows", lu["category"].size, "LMR", int((lu["category"] == 1).sum()), "UMR", int((lu["category"] == 2).sum()))
print("valleys", val["start"].size, "rows replaced", int(val["row_in_valley"].sum()))
print("merged rows", merged["job"].size, "pieces", pieces["job"].size)
print("pmd_call rows", pmd["job"].size, "PMD", int((pmd["category"] == 1).sum()), "positions in PMDs", int(pmd["num_cpgs"][pmd["category"] == 1].sum()))
print("final rows", fin["job"].size, "with category", int((fin["category"] != 0).sum()), "status PMD", int((fin["pmd_status"] == 1).sum()),
      "NA", int((fin["pmd_status"] < 0).sum()), "LMRs dropped inside PMDs", int(((lu["category"] == 1).sum()) - (fin["category"] == 1).sum()))
res.free(); batch.free()
