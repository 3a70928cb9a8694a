import sys, numpy as np
sys.path.insert(0, ".")
from methylasso_b200 import api, synth
from methylasso_b200._native import MlParams
t = synth.chromosome_table(2_260_000, 1, (3,))
eng = api.Engine(0)
res = eng.detect_batch(np.array([0, t["pos"].size], np.int64), t, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
call = res.call_table(0.01, 500.0)
pmd = res.call_pmd_segments(0.01, 500.0, 1000.0, want_fused=True)
f = pmd["row_fused_std"]
heads = np.nonzero(np.diff(f) != 0)[0]
runs = np.diff(np.concatenate([[0], heads + 1, [f.size]]))
print("rows", f.size, "fused runs", runs.size, "longest runs", np.sort(runs)[-8:], "pmd segments", pmd["start"].size)
print("width stats", np.percentile(call["width"], [5, 50, 95]), "std stats", np.percentile(call["std"], [5, 50, 95]))
