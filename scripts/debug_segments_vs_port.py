"""Which chromosomes of the bench workload differ from the CPU port in their segment boundaries, and by how much the
fitted coefficients differ there."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from methylasso_b200 import api
from methylasso_b200._native import MlParams
from oracle import oracle as O
O.build(); O.use_reference_dp(True)
tables, ptr, cols = bench.make_workload(int(sys.argv[1]) if len(sys.argv) > 1 else 28_000_000)
eng = api.Engine(0)
res = eng.detect_batch(ptr, cols, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
seg = res.segments(0.01)
for j, t in enumerate(tables):
    o = O.detect(t, O.MODEL_FAST, lambda2=25.0)
    r = res.job(j, want_mu=False)
    db = np.abs(r["estimates"]["beta"] - o["estimates"]["beta"])
    bh = np.array_equal(r["estimates"]["betahat"], o["estimates"]["betahat"])
    pw = np.array_equal(r["estimates"]["pseudoweight"], o["estimates"]["pseudoweight"])
    oseg = O.segment_methylation_helper(o["estimates"], 0.01)
    m = seg["job"] == j
    same = np.array_equal(seg["start"][m], oseg["start"]) and np.array_equal(seg["end"][m], oseg["end"])
    k = int(np.argmax(db))
    print(j, "P", db.size, "max|dbeta| %.3g at %d" % (db.max(), k), "betahat eq", bh, "pw eq", pw, "segments", int(m.sum()), oseg["start"].size,
          "same" if same else "DIFFERENT", flush=True)
    if db.max() > 1e-9:
        bad = np.nonzero(db > 1e-9)[0]
        print("   first / last bad index", bad[0], bad[-1], "count", bad.size, "beta gpu/cpu", r["estimates"]["beta"][k], o["estimates"]["beta"][k],
              "pw around", o["estimates"]["pseudoweight"][max(0, bad[0] - 3):bad[0] + 4], "y", o["estimates"]["betahat"][max(0, bad[0] - 3):bad[0] + 4])
res.free()
