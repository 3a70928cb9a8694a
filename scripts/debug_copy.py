import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from methylasso_b200 import api, synth
from methylasso_b200._native import MlParams
eng = api.Engine()
tables = synth.genome_tables(2, total_cpgs=14_000_000, reps_per_condition=(3,))
ptr, cols = synth.concat_jobs(tables)
pin = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in cols.items()}
params = MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
b = eng.upload(ptr, pin); r = eng.detect(b, params)
N = int(ptr[-1]); EP = sum(r.sizes(j)[1] * r.sizes(j)[2] for j in range(r.njobs))
def pinned(n, dt): return torch.empty(n, dtype=dt).pin_memory().numpy()
mu = pinned(N, torch.float64); beta = pinned(EP, torch.float64); bh = pinned(EP, torch.float64); pw = pinned(EP, torch.float64)
st = pinned(EP, torch.float64); eid = pinned(EP, torch.int32); off = np.empty(3 * 24)
print("pinned?", torch.from_numpy(mu).is_pinned())
for name, kw in (("mu", dict(mu=mu)), ("beta", dict(beta=beta)), ("beta+bh+pw", dict(beta=beta, betahat=bh, pseudoweight=pw)),
                 ("start", dict(start=st)), ("est_id", dict(estimate_id=eid)), ("all", dict(mu=mu, offset=off, estimate_id=eid, start=st, beta=beta, betahat=bh, pseudoweight=pw))):
    for rep in range(2):
        t0 = time.perf_counter(); r.copy_all_async(**kw); t1 = time.perf_counter(); eng.synchronize(); t2 = time.perf_counter()
    print(f"{name:12s} enqueue {1e3*(t1-t0):7.2f} ms   complete {1e3*(t2-t0):7.2f} ms")

print("---- segmentation alone vs under a D2H in flight")
for mode in ("alone", "with_d2h"):
    r2 = eng.detect(b, params)
    eng.set("profile", 1); eng.profile_reset()
    t0 = time.perf_counter()
    if mode == "with_d2h":
        r2.copy_all_async(mu=mu, offset=None, estimate_id=eid, start=st, beta=beta, betahat=bh, pseudoweight=pw)
    t1 = time.perf_counter()
    r2.segments_count(0.01); r2.pmd_segments(0.01, 1000.0, fetch=False)
    t2 = time.perf_counter(); eng.synchronize(); t3 = time.perf_counter()
    prof = eng.profile(); eng.set("profile", 0)
    tot = sum(v[0] for v in prof.values())
    top = sorted(prof.items(), key=lambda kv: -kv[1][0])[:6]
    print(mode, f"enqueue {1e3*(t1-t0):.2f} compute-host {1e3*(t2-t1):.2f} ms, sync {1e3*(t3-t2):.2f}; kernel total {tot:.2f} ms;", ", ".join(f"{k}={v[0]:.2f}/{v[1]}" for k, v in top))
    r2.free()
