"""Per-kernel timing of the FLSA solver on controlled inputs (design evidence; run under gpurun)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from methylasso_b200 import api, synth

def pooled(t):
    u, inv = np.unique(t["pos"], return_inverse=True)
    cov = np.bincount(inv, t["coverage"], u.size); nm = np.bincount(inv, t["Nmeth"], u.size)
    return nm / cov, cov.astype(float)

eng = api.Engine()
t = synth.chromosome_table(int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 77, (3,))
y, w = pooled(t)
def run(tag, yy, ww, lam, chunk, warm, reps=3):
    eng.set("flsa_chunk", chunk); eng.set("flsa_warm", warm)
    for _ in range(2): eng.weighted_1d_flsa(yy, ww, lam)
    eng.set("profile", 1); eng.profile_reset()
    for _ in range(reps): eng.weighted_1d_flsa(yy, ww, lam)
    prof = eng.profile(); eng.set("profile", 0)
    tot = sum(v[0] for v in prof.values()) / reps
    items = ", ".join(f"{k.replace('flsa_','').replace('_kernel','')}={v[0]/reps:.3f}ms/{v[1]//reps}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:6])
    print(f"{tag:28s} n={yy.size:8d} lam={lam:6.0f} chunk={chunk:5d} warm={warm:5d} total={tot:7.3f}ms | {items}", flush=True)

n = y.size
for chunk, warm in ((512, 512), (256, 512), (1024, 512), (512, 256), (2048, 512), (128, 512)):
    run("genome-like lam25", y, w, 25.0, chunk, warm)
# latency probes: a short series -> few threads
ys, ws = y[:4096], w[:4096]
run("one thread sequential 4096", ys, ws, 25.0, 1 << 30, 512)
run("8 chunks of 512", ys, ws, 25.0, 512, 512)
ys, ws = y[:65536], w[:65536]
run("128 chunks (2 blocks)", ys, ws, 25.0, 512, 512)
ys, ws = y[:65536 * 8], w[:65536 * 8]
run("1024 chunks (16 blocks)", ys, ws, 25.0, 512, 512)
# PMD-like: long memory
rng = np.random.default_rng(1)
m = 300_000
yp = np.abs(rng.normal(0.08, 0.06, m)); wp = rng.lognormal(7.0, 1.5, m)
for chunk, warm in ((512, 512), (512, 2048), (256, 2048), (128, 4096)):
    run("pmd-like lam1000", yp, wp, 1000.0, chunk, warm)
