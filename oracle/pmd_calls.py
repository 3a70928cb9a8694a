"""TEST INFRASTRUCTURE - CPU restatement of R/call.R:169-199 (segment_methylation_singlechr) for one chromosome: the
pooled per-position data gathered along split_call, the PMD calls, their merge and the statistics of the merged regions.

data.table semantics followed (R is not in the image: NOT pinned against R):
  * foverlaps(mdata, split_call): mdata rows are [pos, pos + 1], so a position belongs to a piece when
    start - 1 <= pos <= end (closed intervals, same estimate_id); a position matched by two touching pieces appears in
    both; pieces without a position have no row in pmd_call;
  * the grouped j of :172 holds list() columns, so mean() is data.table's fastmean = R's two-pass long double mean;
    :190 / :196 likewise (mean over the pieces' fused_std; mean / sd over the concatenated betavals, sd as cov.c does);
  * :182-187 shift() with by = condition restarts per condition; pmd.diff is TRUE on a condition's first row;
  * groups of :190 come out in order of first appearance = row order.
Only tests/ may import this module."""
import numpy as np

from .diffcall import r_mean, r_sd

DEFAULTS = dict(pmd_std_threshold=0.15, pmd_max_beta=0.75, pmd_valley_min_width=5000.0, max_distance=500.0, merge_pmds=True)


def pmd_calls(pooled, split, **kw):
    """pooled: estimate_id, pos, Nmeth, coverage per position and condition (mdata, :43-46), sorted by (estimate_id,
    pos); split: split_call after :168 (estimate_id, start, end, fused_std), in key order.  Returns pmd_call after :198:
    condition, category (1 PMD, 0 other), start, end, width, num_cpgs, fused_std, beta, std."""
    p = dict(DEFAULTS, **kw)
    pvw = float(p["pmd_valley_min_width"])
    pe, pos = np.asarray(pooled["estimate_id"], np.int64), np.asarray(pooled["pos"], np.float64)
    val = np.asarray(pooled["Nmeth"], np.float64) / np.asarray(pooled["coverage"], np.float64)
    rows = []          # pmd_call after :175
    for e, s, t, f in zip(split["estimate_id"], split["start"], split["end"], split["fused_std"]):
        sel = np.nonzero((pe == e) & (pos >= s - 1) & (pos <= t))[0]
        if sel.size == 0:
            continue
        rows.append(dict(cond=int(e), start=pos[sel].min(), end=pos[sel].max() + 2, beta=r_mean(val[sel]), n=sel.size,
                         vals=val[sel], fused=float(f)))
    # :177-180
    for r in rows:
        w = r["end"] - r["start"] + 1
        with np.errstate(divide="ignore"):
            adm = (w > pvw) and (np.float64(w) / np.float64(r["n"] - 1) < p["max_distance"])
        r["cat"] = 1 if (adm and r["fused"] >= p["pmd_std_threshold"] and r["beta"] < p["pmd_max_beta"]) else 0
    # :182-189
    out = dict(condition=[], category=[], start=[], end=[], width=[], num_cpgs=[], fused_std=[], beta=[], std=[])
    h = 0
    while h < len(rows):
        t = h + 1
        if p["merge_pmds"]:
            while (t < len(rows) and rows[t]["cond"] == rows[h]["cond"] and rows[t]["cat"] == rows[t - 1]["cat"]
                   and not rows[t]["start"] - rows[t - 1]["end"] > pvw):
                t += 1
        g = rows[h:t]
        vals = np.concatenate([r["vals"] for r in g])
        s, e = min(r["start"] for r in g), max(r["end"] for r in g)
        sd = r_sd(vals)
        for k, v in zip(out, (g[0]["cond"], g[0]["cat"], s, e, e - s + 1, sum(r["n"] for r in g),
                              r_mean([r["fused"] for r in g]), r_mean(vals), 0.0 if np.isnan(sd) else sd)):
            out[k].append(v)
        h = t
    return {k: np.asarray(v, np.int64 if k in ("condition", "category", "num_cpgs") else np.float64) for k, v in out.items()}
