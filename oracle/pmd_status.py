"""TEST INFRASTRUCTURE - CPU restatement of R/call.R:201-215 (segment_methylation_singlechr) for one chromosome: seg_call
(after :140) joined with the PMD table, duplicates removed, LMRs inside PMDs dropped, the UMRs' PMD status taken from
their neighbours.

data.table semantics followed (R is not in the image: NOT pinned against R):
  * foverlaps(seg_call, pmd_call[, .(condition, segment_id, start, end, category)]): type "any" on closed intervals
    within a condition, every match kept in the order of pmd_call, a row without match kept with NA;
  * after :204-207 the rows of one segment differ in PMD.status only, so unique() (:209) leaves its distinct statuses in
    order of first appearance;
  * :211 selects is.admissible == T (NA on valley rows: not selected); shift() walks the selected rows of a condition;
    `NA == x` is NA and ifelse(NA, ...) is NA.
Status codes: 1 "PMD", 0 "other", -1 NA; category 0 NA, 1 LMR, 2 UMR, 3 UMR-DMV.  Only tests/ may import this module."""
import numpy as np


def annotate(M, row_adm, pmd):
    """M: pmd_split.merged_table(); row_adm: is.admissible per call-table row (0 / 1); pmd: pmd_calls.pmd_calls().
    Returns merged (row of M), status, category, adm per row of the final table (before the drop of :214-218)."""
    cond, start, end = M["condition"], M["start"], M["end"]
    merged, status = [], []
    for m in range(cond.size):
        hit = np.nonzero((pmd["condition"] == cond[m]) & (pmd["start"] <= end[m]) & (pmd["end"] >= start[m]))[0]
        seen = []
        for y in hit:
            c = int(pmd["category"][y])
            if c not in seen:
                seen.append(c)
        for s in (seen or [-1]):
            merged.append(m)
            status.append(s)
    merged, status = np.asarray(merged, np.int64), np.asarray(status, np.int64)
    category = M["category"][merged].copy()
    category[(status == 1) & (category == 1)] = 0                                  # :210
    src = M["src_row"][merged]
    adm = np.where(src >= 0, np.asarray(row_adm, np.int64)[np.maximum(src, 0)], -1)
    new = status.copy()
    sel = np.nonzero(adm == 1)[0]                                                   # :211-212
    for r, i in enumerate(sel):
        if category[i] != 2:
            continue
        prev = status[sel[r - 1]] if r > 0 and cond[merged[sel[r - 1]]] == cond[merged[i]] else -1
        nxt = status[sel[r + 1]] if r + 1 < sel.size and cond[merged[sel[r + 1]]] == cond[merged[i]] else -1
        new[i] = -1 if (prev < 0 or nxt < 0) else (prev if prev == nxt else 0)
    return dict(merged=merged, status=new, category=category, adm=adm)
