"""Profiling target: a few resident steps of the bench workload (configs[1]) and nothing else, so that `ncu -k regex:...
-s <skip> -c <n>` lands on steady-state launches.  python scripts/prof_step.py [--cpgs N] [--steps K]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from methylasso_b200 import api, synth          # noqa: E402
from methylasso_b200._native import MlParams    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cpgs", type=int, default=synth.GENOME_CPGS)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--regions", action="store_true", help="the whole rule engine (segment_methylation tables) per step")
args = ap.parse_args()
tables = synth.genome_tables(2, total_cpgs=args.cpgs, reps_per_condition=(3,), seed=synth.BASE_SEED + 2, dmv=True)
ptr, cols = synth.concat_jobs(tables)
eng = api.Engine(0)
batch = eng.upload(ptr, cols)
params = MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
for _ in range(args.steps):
    res = eng.detect(batch, params)
    if args.regions:
        seg, pmd = res.segment_methylation_tables(0.01, 500.0)
        n = (seg["start"].size, pmd["start"].size)
    else:
        n = (res.segments_count(0.01), res.call_table(0.01, 500.0, fetch=False),
             res.call_pmd_segments(0.01, 500.0, 1000.0, fetch=False))
    res.free()
eng.synchronize()
print("ok", n, eng.launch_count)
