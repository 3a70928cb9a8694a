"""End-to-end step (host columns -> region tables on the host) for a few pipeline shapes: engines x edge share."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from methylasso_b200 import synth
from methylasso_b200.pipeline import GenomePipeline

tables = synth.genome_tables(2, total_cpgs=synth.GENOME_CPGS, reps_per_condition=(3,), seed=synth.BASE_SEED + 2, dmv=True)
ptr, cols = synth.concat_jobs(tables)
shapes = [tuple(float(x) for x in a.split(":")) for a in sys.argv[1:]] or [(6, 0.0)]
for G, edge in shapes:
    pipe = GenomePipeline(ptr, cols, n_engines=int(G), edge_share=(edge or None))
    for _ in range(4):
        pipe.segment_methylation()
    pipe.warm_pools(25)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        pipe.segment_methylation()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"engines": int(G), "edge": edge, "ms_median": round(float(np.median(ts)), 2), "ms_each": [round(t, 1) for t in ts],
                      "timeline": sorted(pipe.timeline)}), flush=True)
    pipe.close()
