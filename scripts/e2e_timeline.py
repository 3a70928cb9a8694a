"""Timeline of the end-to-end step (GenomePipeline.segment_methylation, bench.py's e2e): start / end of every kernel of every
engine on one clock (ml_engine_profile_spans), written as JSON next to a summary: how busy the GPU is between the first and
the last kernel, how long each engine's chain of launches takes against the sum of its kernels, which kernels run alone.

    python scripts/e2e_timeline.py [engines] [out.json]
"""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from methylasso_b200 import synth
from methylasso_b200.pipeline import GenomePipeline

G = int(sys.argv[1]) if len(sys.argv) > 1 else 6
out = sys.argv[2] if len(sys.argv) > 2 else None
tables = synth.genome_tables(2, total_cpgs=synth.GENOME_CPGS, reps_per_condition=(3,), seed=synth.BASE_SEED + 2, dmv=True)
ptr, cols = synth.concat_jobs(tables)
pipe = GenomePipeline(ptr, cols, n_engines=G)
for _ in range(4):
    pipe.segment_methylation()
pipe.warm_pools(25)
plain = []
for _ in range(5):
    t0 = time.perf_counter(); pipe.segment_methylation(); plain.append((time.perf_counter() - t0) * 1e3)
for e in pipe.pool.engines:
    e.set("profile_reserve_events", 6000)
    e.set("profile", 1)
    e.set("profile_timeline", 1)
pipe.segment_methylation()
for e in pipe.pool.engines:
    e.profile_reset()
t0 = time.perf_counter(); pipe.segment_methylation(); prof_ms = (time.perf_counter() - t0) * 1e3
spans = [e.profile_spans() for e in pipe.pool.engines]
host = sorted(pipe.timeline)
allk = sorted((a, b, i, n) for i, sp in enumerate(spans) for n, a, b in sp)
lo, hi = allk[0][0], max(b for _, b, _, _ in allk)
# union of the busy intervals
busy, cur_a, cur_b = 0.0, None, None
for a, b, _, _ in allk:
    if cur_b is None or a > cur_b:
        if cur_b is not None:
            busy += cur_b - cur_a
        cur_a, cur_b = a, b
    else:
        cur_b = max(cur_b, b)
busy += cur_b - cur_a
summary = {"engines": G, "ms_plain_each": [round(x, 2) for x in plain], "ms_profiled": round(prof_ms, 2),
           "first_kernel_to_last_ms": round(hi - lo, 3), "gpu_busy_union_ms": round(busy, 3),
           "sum_of_kernel_ms": round(sum(b - a for a, b, _, _ in allk), 3), "launches": len(allk), "host_marks_ms": host,
           "per_engine": []}
for i, sp in enumerate(spans):
    a0, b1 = min(a for _, a, _ in sp), max(b for _, _, b in sp)
    gaps = sorted(((sp[k + 1][1] - sp[k][2], sp[k][0], sp[k + 1][0]) for k in range(len(sp) - 1)), reverse=True)
    summary["per_engine"].append({"first": round(a0 - lo, 3), "last": round(b1 - lo, 3), "launches": len(sp),
                                  "kernel_ms": round(sum(b - a for _, a, b in sp), 3),
                                  "largest_gaps_ms": [(round(g, 3), x, y) for g, x, y in gaps[:6]]})
# kernels stretched by sharing the GPU: time in the pipeline against the resident step's table would need both; list the top
agg = {}
for a, b, _, n in allk:
    v = agg.setdefault(n, [0.0, 0]); v[0] += b - a; v[1] += 1
summary["top_kernels"] = [(n, round(v[0], 3), v[1]) for n, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]]
print(json.dumps(summary))
if out:
    with open(out, "w") as f:
        json.dump({"summary": summary, "spans": [[(n, round(a - lo, 4), round(b - lo, 4)) for n, a, b in sp] for sp in spans]}, f)
pipe.close()
