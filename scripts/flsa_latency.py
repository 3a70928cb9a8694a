"""Per-step latency of a lone DP chain: one series in sequential mode (one chunk: one thread of the speculative kernel
runs it exactly from the head; if its window outgrows the 16-knot ring a lone walker pair redoes it)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from methylasso_b200 import api, synth
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import pooled_series

eng = api.Engine(0)
t = synth.chromosome_table(400_000, 99, (3,))
y, w = pooled_series(t)
n = y.size
for mode in (1, 0):
    eng.set("flsa_sequential", mode)
    eng.weighted_1d_flsa(y, w, 25.0)
    eng.set("profile", 1)
    eng.profile_reset()
    eng.weighted_1d_flsa(y, w, 25.0)
    pr = eng.profile()
    print("sequential" if mode else "chunked", n, {k: round(v[0], 3) for k, v in pr.items() if k.startswith("flsa")})
    eng.set("profile", 0)
    if mode:
        for k, v in pr.items():
            if k in ("flsa_spec_kernel", "flsa_chain_kernel"):
                print(f"  {k}: {v[0] * 1e6 / n:.1f} ns per step = {v[0] * 1e-3 * 1.965e9 / n:.0f} cycles")
