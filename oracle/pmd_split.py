"""TEST INFRASTRUCTURE - CPU restatement of R/call.R:131-164 (segment_methylation_singlechr) for one chromosome: the
valleys replace the call-table rows they overlap (:131-140), the table is cut into "isolate" / "complement" regions
(:143-156) and the PMD candidate segments of :93-95 (std_call) are split along those regions (:158-164, split_call).

data.table semantics followed (R is not in the image: NOT pinned against R):
  * rbind(seg_call, new_valleys, fill = T) then setkey(condition, start, end): a stable sort, call-table rows before
    valleys on equal keys; shift() with `by` restarts per condition / estimate_id;
  * :154 groups by (estimate_id, segment_id, category) where segment_id = cumsum(cat.chg) + cumsum(follows.large.gap):
    one group per run of rows between changes, in row order;
  * foverlaps(std_call, umr_complement): type "any" on closed intervals within an estimate_id, every match kept, a
    std_call row without match kept with NA columns; :159 then emits one row per match (j returns vectors), :161
    zeroes fused_std outside "complement", :163 sorts by (estimate_id, start, end) with NA first, :164 numbers the rows
    per estimate_id and :168 drops the rows without an interval.
Only tests/ may import this module."""
import numpy as np

NA_CAT, LMR, UMR, UMR_DMV = 0, 1, 2, 3


def merged_table(call, category, val, pmd_valley_min_width=5000.0):
    """:131-140.  call / category: the annotated call table; val: the result of valleys.valleys().  Returns the new
    seg_call: condition, start, end, category (0 NA, 1 LMR, 2 UMR, 3 UMR-DMV), src_row (call-table row, -1 for a valley),
    src_valley (-1 for a call-table row), segment_id (1-based per condition), follows_large_gap."""
    keep = np.nonzero(~np.asarray(val["row_in_valley"], bool))[0]
    v = val["valleys"]
    nv = v["start"].size
    cond = np.concatenate([np.asarray(call["estimate_id"], np.int64)[keep], v["condition"].astype(np.int64)])
    start = np.concatenate([np.asarray(call["start"], np.float64)[keep], v["start"]])
    end = np.concatenate([np.asarray(call["end"], np.float64)[keep], v["end"]])
    cat = np.concatenate([np.asarray(category, np.int64)[keep], np.full(nv, UMR_DMV, np.int64)])
    src_row = np.concatenate([keep, np.full(nv, -1, np.int64)])
    src_valley = np.concatenate([np.full(keep.size, -1, np.int64), np.arange(nv)])
    order = np.lexsort((end, start, cond))
    cond, start, end, cat, src_row, src_valley = (a[order] for a in (cond, start, end, cat, src_row, src_valley))
    n = cond.size
    seg_id = np.zeros(n, np.int64)
    gap = np.zeros(n, bool)
    for c in np.unique(cond):
        m = cond == c
        seg_id[m] = np.arange(1, m.sum() + 1)
    if n > 1:
        gap[1:] = (cond[1:] == cond[:-1]) & (start[1:] - end[:-1] > pmd_valley_min_width)
    return dict(condition=cond, start=start, end=end, category=cat, src_row=src_row, src_valley=src_valley,
                segment_id=seg_id, follows_large_gap=gap)


def regions(M, split_pmds=True):
    """:143-156  umr_complement: estimate_id, start, end, isolate (TRUE = "isolate", FALSE = "complement")."""
    cat = M["category"]
    iso = ((cat == UMR) | (cat == UMR_DMV)) if split_pmds else (cat == UMR_DMV)
    cond = M["condition"]
    n = cond.size
    out = dict(estimate_id=[], start=[], end=[], isolate=[])
    h = 0
    while h < n:
        t = h + 1
        while t < n and cond[t] == cond[h] and iso[t] == iso[t - 1] and not M["follows_large_gap"][t]:
            t += 1
        out["estimate_id"].append(int(cond[h]))
        out["start"].append(M["start"][h:t].min())
        out["end"].append(M["end"][h:t].max())
        out["isolate"].append(bool(iso[h]))
        h = t
    r = dict(estimate_id=np.asarray(out["estimate_id"], np.int64), start=np.asarray(out["start"], np.float64),
             end=np.asarray(out["end"], np.float64), isolate=np.asarray(out["isolate"], bool))
    order = np.lexsort((r["end"], r["start"], r["estimate_id"]))
    return {k: a[order] for k, a in r.items()}


def split_call(reg, std_call):
    """:158-168.  std_call: estimate_id, start, end, fused_std, pseudoweight (the helper's table of :95).  Returns
    estimate_id, segment_id, start, end, fused_std, pseudoweight."""
    e = np.asarray(std_call["estimate_id"], np.int64)
    xs, xe = np.asarray(std_call["start"], np.float64), np.asarray(std_call["end"], np.float64)
    rows = []
    for i in range(e.size):
        hit = np.nonzero((reg["estimate_id"] == e[i]) & (reg["start"] <= xe[i]) & (reg["end"] >= xs[i]))[0]
        if hit.size == 0:
            rows.append((e[i], np.nan, np.nan, float(std_call["fused_std"][i]), float(std_call["pseudoweight"][i])))
        for y in hit:
            rows.append((e[i], max(reg["start"][y], xs[i]), min(reg["end"][y], xe[i]),
                         0.0 if reg["isolate"][y] else float(std_call["fused_std"][i]), float(std_call["pseudoweight"][i])))
    a = np.asarray(rows, np.float64).reshape(-1, 5)
    # setkey: NA sorts first
    ks, ke = np.where(np.isnan(a[:, 1]), -np.inf, a[:, 1]), np.where(np.isnan(a[:, 2]), -np.inf, a[:, 2])
    a = a[np.lexsort((ke, ks, a[:, 0]))]
    seg_id = np.zeros(a.shape[0], np.int64)
    for c in np.unique(a[:, 0]):
        m = a[:, 0] == c
        seg_id[m] = np.arange(1, m.sum() + 1)
    with np.errstate(invalid="ignore"):
        ok = a[:, 1] <= a[:, 2]
    return dict(estimate_id=a[ok, 0].astype(np.int64), segment_id=seg_id[ok], start=a[ok, 1], end=a[ok, 2],
                fused_std=a[ok, 3], pseudoweight=a[ok, 4])
