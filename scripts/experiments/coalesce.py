"""Design experiment: coalescence length of the FLSA DP state (see coalesce.c)."""
import ctypes as C, numpy as np, sys, subprocess, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from methylasso_b200 import synth
L = C.CDLL("/tmp/coalesce.so")
def run(y, w, lam, stride=5000, maxsteps=400000):
    n = y.size; m = n // stride + 2
    ln = np.zeros(m, np.int32); be = np.zeros(m, np.int32); cnt = C.c_int(0); st = np.zeros(4, np.int32)
    L.experiment(y.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), n, C.c_double(lam), stride, maxsteps,
                 ln.ctypes.data_as(C.c_void_p), be.ctypes.data_as(C.c_void_p), C.byref(cnt), st.ctypes.data_as(C.c_void_p))
    ln, be = ln[:cnt.value], be[:cnt.value]
    return ln, be, st
if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    for reps in ((1,), (3,)):
        t = synth.chromosome_table(n, 20242, reps)
        # pooled pseudodata per unique position (FAST path semantics)
        pos = t["pos"]; u, inv = np.unique(pos, return_inverse=True)
        cov = np.bincount(inv, t["coverage"], u.size); nm = np.bincount(inv, t["Nmeth"], u.size)
        y = nm / cov; w = cov.astype(float)
        for lam in (10., 25., 100., 400., 1000., 2000.):
            ln, be, st = run(y, w, lam)
            ok = ln > 0
            q = np.percentile(ln[ok], [50, 90, 99, 100]) if ok.any() else [0]*4
            print(f"reps={reps} lam={lam:7.1f} starts={ln.size} coalesced={ok.mean():.3f} len p50/p90/p99/max={q} biteq={be[ok].mean() if ok.any() else 0:.4f} max_knots={st[1]}")
