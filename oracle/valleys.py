"""TEST INFRASTRUCTURE - CPU restatement of the valley (DMV) block of segment_methylation_singlechr, R/call.R:100-133,
on the annotated call table (oracle.call_table + lmr_umr.annotate), for one chromosome.

data.table semantics followed here (R is not in the image: NOT pinned against R):
  * the call table is keyed by (condition, start, end); `shift()` with by = condition restarts per condition;
  * `:=` with `by` on the rows `is.high == F` leaves NA (NaN here) on the other rows;
  * none of the grouped `j` expressions is GForce-eligible (each holds an arithmetic expression, a constant or %in%),
    so mean() is data.table's fastmean = R's two-pass long double mean and sum() of doubles is R's long double sum
    (summary.c rsum), rounded to double once; numpy.longdouble is the x87 80-bit type on x86-64 Linux;
  * groups come out in order of first appearance; setkey() is a stable sort on (condition, start, end);
  * foverlaps(x, y, type = "any"): x overlaps y when x.start <= y.end and x.end >= y.start, same condition.
Only tests/ may import this module."""
import numpy as np

from .diffcall import LD, r_mean

DEFAULTS = dict(valley_max_beta=0.1, pmd_valley_min_width=5000.0, max_distance=500.0, min_num_cpgs=4)
DMV, LOW, UMR = 1, 0, 2     # category codes of the intermediate table new_valleys


def r_sum(x):
    s = LD(0)
    for v in np.asarray(x, np.float64):
        s += LD(v)
    return float(s)


def _overlaps_any(cond, start, end, vcond, vstart, vend):
    out = np.zeros(cond.size, bool)
    for i in range(cond.size):
        out[i] = bool(np.any((vcond == cond[i]) & (vstart <= end[i]) & (vend >= start[i])))
    return out


def valleys(call, category, **kw):
    """call: the call table of ONE chromosome ordered by (condition, start, end): estimate_id, start, end, width, beta,
    coverage, pseudoweight, num_cpgs; category: 0 NA, 1 LMR, 2 UMR per row (R/call.R:83-86).
    Returns dict(valleys = the `new_valleys` table after :129 (category "UMR-DMV": condition, start, end, width, beta,
    coverage, pseudoweight, num_cpgs), row_in_valley = per call-table row, TRUE when :132-133 removes it,
    low = the intermediate per-row columns of :102-106 (is_high, follows_large_gap, low_id, low_id_width, low_id_beta))."""
    p = dict(DEFAULTS, **kw)
    vmax, pvw = float(p["valley_max_beta"]), float(p["pmd_valley_min_width"])
    cond = np.asarray(call["estimate_id"], np.int64)
    start, end = np.asarray(call["start"], np.float64), np.asarray(call["end"], np.float64)
    beta, cov = np.asarray(call["beta"], np.float64), np.asarray(call["coverage"], np.float64)
    pw, ncpg = np.asarray(call["pseudoweight"], np.float64), np.asarray(call["num_cpgs"], np.int64)
    category = np.asarray(category)
    n = cond.size
    # :102-105
    is_high = beta > vmax
    gap = np.zeros(n, bool)
    if n > 1:
        gap[1:] = (cond[1:] == cond[:-1]) & (start[1:] - end[:-1] > pvw)
    low_id = np.zeros(n, np.int64)
    for c in np.unique(cond):
        m = cond == c
        low_id[m] = np.cumsum(is_high[m]) + np.cumsum(gap[m])
    # :106  by (low.id, condition) over the rows that are not high
    low_w, low_b = np.full(n, np.nan), np.full(n, np.nan)
    groups = {}
    for i in np.nonzero(~is_high)[0]:
        groups.setdefault((int(cond[i]), int(low_id[i])), []).append(i)
    for rows in groups.values():
        rows = np.asarray(rows)
        low_w[rows] = end[rows].max() - start[rows].min() + 1
        low_b[rows] = r_mean(beta[rows])
    # :107-111  one row per (low.id, low.id.width, low.id.beta, condition), first appearance order
    nv = dict(condition=[], start=[], end=[], width=[], beta=[], coverage=[], pseudoweight=[], num_cpgs=[], category=[])

    def add(c, s, e, w, b, cv, pwv, m, cat):
        for k, v in zip(nv, (c, s, e, w, b, cv, pwv, m, cat)):
            nv[k].append(v)

    with np.errstate(invalid="ignore"):
        sel = (~is_high) & (low_b <= vmax)
    g2 = {}
    for i in np.nonzero(sel)[0]:
        g2.setdefault((int(low_id[i]), float(low_w[i]), float(low_b[i]), int(cond[i])), []).append(i)
    for (lid, w, b, c), rows in g2.items():
        rows = np.asarray(rows)
        add(c, start[rows].min(), end[rows].max(), w, b, float(cov[rows].sum()), r_sum(pw[rows]), int(ncpg[rows].sum()),
            DMV if w > pvw else LOW)
    # :113-116  UMRs that overlap none of these regions are added with their own low.id columns (NA when high)
    v = {k: np.asarray(a, np.float64 if k not in ("condition", "num_cpgs", "category") else np.int64) for k, a in nv.items()}
    umr = np.nonzero(category == 2)[0]
    if umr.size:
        hit = _overlaps_any(cond[umr], start[umr], end[umr], v["condition"], v["start"], v["end"])
        for i in umr[~hit]:
            add(int(cond[i]), start[i], end[i], low_w[i], low_b[i], float(cov[i]), float(pw[i]), int(ncpg[i]), UMR)
    v = {k: np.asarray(a, np.float64 if k not in ("condition", "num_cpgs", "category") else np.int64) for k, a in nv.items()}
    order = np.lexsort((v["end"], v["start"], v["condition"]))            # stable
    v = {k: a[order] for k, a in v.items()}
    # :119-129  merge consecutive valleys
    m = v["condition"].size
    sep = np.zeros(m, bool)
    if m > 1:
        sep[1:] = (v["condition"][1:] == v["condition"][:-1]) & (v["start"][1:] - v["end"][:-1] > pvw)
    out = dict(condition=[], start=[], end=[], width=[], beta=[], coverage=[], pseudoweight=[], num_cpgs=[])
    first = np.ones(m, bool)
    if m > 1:
        first[1:] = (v["condition"][1:] != v["condition"][:-1]) | sep[1:]
    heads = np.nonzero(first)[0]
    for h, t in zip(heads, np.append(heads[1:], m)):
        rows = np.arange(h, t)
        s, e = v["start"][rows].min(), v["end"][rows].max()
        w, k = e - s + 1, int(v["num_cpgs"][rows].sum())
        contains = bool(np.any(v["category"][rows] == DMV))
        with np.errstate(divide="ignore", invalid="ignore"):
            ok = (w > pvw) and (np.float64(w) / np.float64(k - 1) < p["max_distance"]) and (k >= p["min_num_cpgs"])
        if not (contains and ok):
            continue
        for key, val in zip(out, (int(v["condition"][h]), s, e, w, r_mean(v["beta"][rows]), float(v["coverage"][rows].sum()),
                                  r_sum(v["pseudoweight"][rows]), k)):
            out[key].append(val)
    res = {k: np.asarray(a, np.int64 if k in ("condition", "num_cpgs") else np.float64) for k, a in out.items()}
    # :132-133  call-table rows that overlap a valley are replaced by it
    row_in = _overlaps_any(cond, start, end, res["condition"], res["start"], res["end"]) if n else np.zeros(0, bool)
    return dict(valleys=res, row_in_valley=row_in,
                low=dict(is_high=is_high, follows_large_gap=gap, low_id=low_id, low_id_width=low_w, low_id_beta=low_b))
