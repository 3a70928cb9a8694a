import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from methylasso_b200 import api, synth
from oracle import oracle as O
eng = api.Engine()
t = synth.chromosome_table(400_000, 20241, (1,))
u, inv = np.unique(t["pos"], return_inverse=True)
cov = np.bincount(inv, t["coverage"], u.size); nm = np.bincount(inv, t["Nmeth"], u.size)
y, w = nm / cov, cov.astype(float)
n = 20000
yy, ww = np.ascontiguousarray(y[:n]), np.ascontiguousarray(w[:n])
rb, rtm, rtp = O.tf_dp_weight(yy, ww, 1000.0, "port")
p = lambda a: a.ctypes.data_as(C.c_void_p)
for chunk, warm in ((512, 512), (512, 4096), (512, 1 << 20)):
    eng.set("flsa_chunk", chunk); eng.set("flsa_warm", warm)
    beta, tm, tp = np.zeros(n), np.zeros(n), np.zeros(n)
    rc = eng.lib.ml_debug_flsa_backpointers(eng.h, n, p(yy), p(ww), 1000.0, p(beta), p(tm), p(tp))
    btm = np.nonzero(tm[:n-1] != rtm)[0]; btp = np.nonzero(tp[:n-1] != rtp)[0]; bb = np.nonzero(beta != rb)[0]
    print(chunk, warm, "rc", rc, "tm bad", btm.size, btm[:5], "tp bad", btp.size, btp[:5], "beta bad", bb.size, bb[:3], flush=True)
    if btm.size:
        k = btm[0]
        print("  first bad tm at", k, "chunk", (k - 1) // 512, "offset", (k - 1) % 512, tm[k], rtm[k], "tp", tp[k], rtp[k])
        print("  bad tm chunks:", np.unique((btm - 1) // 512)[:40])
