"""TEST INFRASTRUCTURE — CPU restatement of the reference's DMR caller, R/call.R:254-333 (call_differences_singlechr)
and :348-365 (call_differences), line by line in numpy / plain Python for small cases.

R is not in this image, so this restatement is NOT pinned against R itself ("parity unpinned" above the Rcpp module,
as for the call table of oracle.call_table); it follows the R source and R's published semantics:
  * data.table foverlaps(type = "any") of [pos, pos + 1] with [start, end]  ->  start - 1 <= pos <= end;
  * mean(): long double sum / n refined by a second pass (summary.c real_mean); sd(): sqrt of the long double sum of
    squares about that mean / (n - 1) (cov.c).  numpy.longdouble is the x87 80-bit type on x86-64 Linux;
  * wilcox.test / p.adjust: oracle.wilcox_test / oracle.p_adjust_fdr (pinned against scipy, tests/test_wilcoxon.py).
Only tests/ and __graft_entry__.smoke() may import this module.
"""
import numpy as np

from . import oracle as O

LD = np.longdouble


def r_mean(x):
    x = np.asarray(x, np.float64)
    n = x.size
    if n == 0:
        return np.nan
    s = LD(0)
    for v in x:
        s += LD(v)
    s /= n
    if np.isfinite(float(s)):
        t = LD(0)
        for v in x:
            t += LD(v) - s
        s += t / n
    return float(s)


def r_sd(x):
    x = np.asarray(x, np.float64)
    n = x.size
    if n < 2:
        return np.nan
    m = LD(r_mean(x))           # cov.c stores the refined mean as a double
    ss = LD(0)
    for v in x:
        d = LD(v) - m
        ss += d * d
    return float(np.sqrt(float(ss / (n - 1))))


def pooled_mdata(table):
    """R/call.R:258-263: per (condition, pos): coverage and Nmeth summed, beta = mean of the per-row Nmeth / coverage."""
    out = {}
    beta_row = table["Nmeth"].astype(np.float64) / table["coverage"].astype(np.float64)
    for c in np.unique(table["condition"]):
        m = table["condition"] == c
        pos = table["pos"][m]
        u, inv = np.unique(pos, return_inverse=True)
        order = np.argsort(inv, kind="stable")           # rows of one position in data order
        cuts = np.searchsorted(inv[order], np.arange(u.size + 1))
        br = beta_row[m][order]
        out[int(c)] = dict(pos=u.astype(np.int64),
                           coverage=np.bincount(inv, table["coverage"][m], u.size),
                           Nmeth=np.bincount(inv, table["Nmeth"][m], u.size),
                           beta=np.array([r_mean(br[cuts[k]:cuts[k + 1]]) for k in range(u.size)]))
    return out


def call_differences_singlechr(table, estimates, ref_cond, min_diff=0.1, min_delta_diff=0.1, tol_val=0.01,
                               max_distance=500.0):
    """table: the six input columns of one chromosome; estimates: the table difference_detection_fast returns
    (estimate 1 = common signal, ids > 1 = differences); ref_cond: integer code of the reference condition.
    Returns the group_call table of R/call.R:302-332 as a dict of arrays, ordered by (estimate_id, segment_id)."""
    md = pooled_mdata(table)
    conds = sorted(md)
    nonref = [c for c in conds if c != ref_cond]
    keep = estimates["estimate_id"] > 1                                     # :266 drop the common signal estimate
    diff_df = {k: np.asarray(v)[keep] for k, v in estimates.items()}
    seg = O.segment_methylation_helper(diff_df, tol_val)                    # :273
    seg_ids = np.unique(seg["estimate_id"])
    assert seg_ids.size == len(nonref), "stopifnot(uniqueN(estimate_id) == uniqueN(condition))"   # :274
    n_datasets = int(np.max(table["dataset"]))
    # :326-329 score per position: number of data rows at that position * 100 / number of datasets
    upos, cnt = np.unique(table["pos"], return_counts=True)
    score = cnt.astype(np.float64) * 100 / n_datasets
    rows = []
    for rank, cond in enumerate(nonref):
        est_id = rank + 2                                                   # :270 as.integer(factor(condition)) + 1
        # :268-271 rows of this estimate: the condition's own positions, then the reference's, keyed by position
        # (setkey is a stable sort: at equal positions the condition's row comes first)
        a, r = md[cond], md.get(ref_cond, dict(pos=np.zeros(0, np.int64), coverage=np.zeros(0), Nmeth=np.zeros(0),
                                               beta=np.zeros(0)))
        pos = np.concatenate([a["pos"], r["pos"]])
        isref = np.concatenate([np.zeros(a["pos"].size, bool), np.ones(r["pos"].size, bool)])
        beta = np.concatenate([a["beta"], r["beta"]])
        cov = np.concatenate([a["coverage"], r["coverage"]]).astype(np.float64)
        o = np.argsort(pos, kind="stable")
        pos, isref, beta, cov = pos[o], isref[o], beta[o], cov[o]
        sm = seg["estimate_id"] == seg_ids[rank] if est_id in seg_ids else seg["estimate_id"] == est_id
        s_start, s_end = seg["start"][sm].astype(np.int64), seg["end"][sm].astype(np.int64)
        s_id, s_pw = seg["segment_id"][sm].astype(np.int64), seg["pseudoweight"][sm]
        parts = []          # seg_call rows after dcast: one per (segment_id after gap shifts)
        shift = 0
        for k in range(s_start.size):
            if s_start[k] > s_end[k]:                                        # :277
                continue
            idx = np.nonzero((pos >= s_start[k] - 1) & (pos <= s_end[k]))[0]  # :278 foverlaps
            if idx.size == 0:
                continue
            over = np.zeros(idx.size, bool)
            over[1:] = (pos[idx][1:] - pos[idx][:-1]) > max_distance         # :282-283
            cuts = np.nonzero(over)[0]
            for j, part in enumerate(np.split(idx, cuts)):
                if j > 0:
                    shift += 1                                               # :284 cumsum(is.oversize) by estimate
                p = pos[part]
                st, en = int(p.min()), int(p.max()) + 2                      # :287
                x, y = part[~isref[part]], part[isref[part]]
                parts.append(dict(segment_id=int(s_id[k]) + shift, start=st, end=en, width=en - st + 1.0,   # :288
                                  pseudoweight=float(s_pw[k]),
                                  beta=r_mean(beta[x]) if x.size else np.nan, coverage=float(cov[x].sum()),
                                  num=int(np.unique(pos[x]).size), vals=beta[x],
                                  beta_ref=r_mean(beta[y]) if y.size else np.nan, coverage_ref=float(cov[y].sum()),
                                  num_ref=int(np.unique(pos[y]).size), vals_ref=beta[y]))
        # dcast keys on (segment_id, start, end, width, pseudoweight): parts are unique there by construction
        for p in parts:
            p["diff"] = p["beta"] - p["beta_ref"]                             # :298
        # :301-304 group neighbouring segments
        gid, groups = 0, []
        for i, p in enumerate(parts):
            change = True
            if i > 0:
                q = parts[i - 1]
                ok = (abs(p["diff"] - q["diff"]) <= min_delta_diff and abs(p["diff"]) >= min_diff
                      and abs(q["diff"]) >= min_diff and p["start"] - q["end"] < max_distance)
                # NA comparisons make the whole condition NA -> group_change TRUE (:303)
                change = not ok if not (np.isnan(p["diff"]) or np.isnan(q["diff"])) else True
            if change:
                gid += 1
                groups.append([])
            groups[-1].append(p)
        for g, ps in enumerate(groups, start=1):
            vals = np.concatenate([p["vals"] for p in ps])
            vals_ref = np.concatenate([p["vals_ref"] for p in ps])
            num, num_ref = sum(p["num"] for p in ps), sum(p["num_ref"] for p in ps)
            covs, covs_ref = sum(p["coverage"] for p in ps), sum(p["coverage_ref"] for p in ps)
            b = r_mean(vals) if num > 0 else np.nan                           # :311
            sd = r_sd(vals) if num > 0 else np.nan
            br = r_mean(vals_ref) if num_ref > 0 else np.nan                  # :312
            sdr = r_sd(vals_ref) if num_ref > 0 else np.nan
            diff = b - br
            with np.errstate(all="ignore"):
                cohen = diff / np.sqrt(((num - 1) * sd ** 2 + (num_ref - 1) * sdr ** 2) / np.float64(num + num_ref - 2))  # :314
            pval = np.nan
            if covs > 0 and covs_ref > 0:                                     # :317
                pval = O.wilcox_test(vals, vals_ref)[1]
            if np.isnan(pval):
                pval = 1.0                                                    # :319
            start, end = min(p["start"] for p in ps), max(p["end"] for p in ps)
            inside = (upos >= start - 1) & (upos <= end)                      # :330 foverlaps with [pos, pos + 1]
            rows.append(dict(condition=cond, estimate_id=est_id, segment_id=g, start=start, end=end,
                             width=float(sum(p["width"] for p in ps)), beta=b, std=sd, coverage=covs, num_cpgs=num,
                             beta_ref=br, std_ref=sdr, coverage_ref=covs_ref, num_cpgs_ref=num_ref,
                             pseudoweight=float(sum(p["pseudoweight"] for p in ps)), diff=diff, cohen_d=float(cohen),
                             pval=pval, coverage_score=r_mean(score[inside])))
    keys = ("condition", "estimate_id", "segment_id", "start", "end", "width", "beta", "std", "coverage", "num_cpgs",
            "beta_ref", "std_ref", "coverage_ref", "num_cpgs_ref", "pseudoweight", "diff", "cohen_d", "pval",
            "coverage_score")
    ints = {"condition", "estimate_id", "segment_id", "start", "end", "num_cpgs", "num_cpgs_ref"}
    return {k: np.array([r[k] for r in rows], np.int64 if k in ints else np.float64) for k in keys}


def call_differences(group_calls, min_diff=0.1, pval_cutoff=0.05, fdr_cutoff=1.0, cov_score=70.0, min_num_cpgs=4):
    """R/call.R:360-363 on the per-chromosome tables concatenated in chromosome order: fdr over the genome, then the
    threshold filter.  Returns (table with `job` and `fdr` columns, boolean mask of the rows the reference keeps)."""
    keys = list(group_calls[0].keys())
    t = {k: np.concatenate([g[k] for g in group_calls]) for k in keys}
    t["job"] = np.concatenate([np.full(g["start"].size, j, np.int64) for j, g in enumerate(group_calls)])
    t["fdr"] = O.p_adjust_fdr(t["pval"])
    with np.errstate(invalid="ignore"):
        keep = (~np.isnan(t["diff"]) & (np.abs(t["diff"]) >= min_diff) & (t["coverage_score"] >= cov_score)
                & (t["pval"] <= pval_cutoff) & (t["fdr"] <= fdr_cutoff)
                & (np.minimum(t["num_cpgs"], t["num_cpgs_ref"]) >= min_num_cpgs))
    return t, keep
