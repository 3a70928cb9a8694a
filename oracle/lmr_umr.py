"""TEST INFRASTRUCTURE - CPU restatement of the LMR / UMR annotation block of segment_methylation_singlechr,
R/call.R:75-91, on the call table of R/call.R:43-72 (oracle.call_table), for one chromosome.

Every `seg_call[i, col := expr, by = ...]` line is restated with data.table's semantics: rows where `i` is NA are not
selected; shift() inside such a statement walks the SELECTED rows (of the `by` group), not the table; logical columns
are three-valued (here: 1 TRUE, 0 FALSE, -1 NA; numeric NA is NaN).  R is not in the image: not pinned against R.
Only tests/ may import this module."""
import numpy as np

T, F, NA = 1, 0, -1
DEFAULTS = dict(min_width=30.0, min_num_cpgs=4, max_distance=500.0, flank_width=300.0, flank_dist_max=5000.0,
                lmr_max_beta=0.5, flank_num_cpgs=10, flank_beta_min=0.5, umr_max_beta=0.1, umr_std_threshold=0.1,
                umr_large_width=500.0, umr_large_density=0.03)


def _and(a, b):
    """R's three-valued & on int8 codes."""
    out = np.full(a.shape, NA, np.int8)
    out[(a == F) | (b == F)] = F
    out[(a == T) & (b == T)] = T
    return out


def _or(a, b):
    out = np.full(a.shape, NA, np.int8)
    out[(a == T) | (b == T)] = T
    out[(a == F) & (b == F)] = F
    return out


def _lt(x, y):
    """x < y with NaN -> NA"""
    out = np.where(x < y, T, F).astype(np.int8)
    out[np.isnan(x) | np.isnan(y)] = NA
    return out


def _shift(v, idx, groups, lead, fill):
    """shift(v) / shift(v, type = 'lead') over the selected rows idx, restarting in every group (groups may be None)."""
    out = np.full(idx.size, fill, dtype=v.dtype if v.dtype.kind == "f" else np.float64 if fill != fill else v.dtype)
    if idx.size == 0:
        return out
    src = v[idx]
    if lead:
        out[:-1] = src[1:]
        same = np.ones(idx.size, bool)
        if groups is not None:
            same[:-1] = groups[idx][1:] == groups[idx][:-1]
        same[-1] = False
    else:
        out[1:] = src[:-1]
        same = np.ones(idx.size, bool)
        if groups is not None:
            same[1:] = groups[idx][1:] == groups[idx][:-1]
        same[0] = False
    out[~same] = fill
    return out


def annotate(call, **kw):
    """call: dict with condition (estimate_id), start, end, width, beta, num_cpgs, std of ONE chromosome, ordered by
    (condition, start, end).  Returns is_admissible, is_V, is_flank, is_sep (int8 codes), flank_dist, flank_beta
    (NaN = NA) and category (0 NA, 1 LMR, 2 UMR)."""
    p = dict(DEFAULTS, **kw)
    cond = np.asarray(call["estimate_id"])
    start, end = np.asarray(call["start"], np.float64), np.asarray(call["end"], np.float64)
    width, beta = np.asarray(call["width"], np.float64), np.asarray(call["beta"], np.float64)
    ncpg, std = np.asarray(call["num_cpgs"], np.float64), np.asarray(call["std"], np.float64)
    n = cond.size
    with np.errstate(divide="ignore", invalid="ignore"):
        # :77
        adm = (width > p["min_width"]) & (width / (ncpg - 1) < p["max_distance"]) & (ncpg >= p["min_num_cpgs"])
    # :78  among the admissible rows of each condition
    is_v = np.full(n, NA, np.int8)
    idx = np.nonzero(adm)[0]
    b = beta[idx]
    is_v[idx] = _and(_lt(b, _shift(beta, idx, cond, False, np.nan)), _lt(b, _shift(beta, idx, cond, True, np.nan)))
    # :79  rows where (is.V == T) | (width > flank.width & is.admissible == T); shift() has no `by` here
    sel = _or(np.where(is_v == NA, NA, (is_v == T)).astype(np.int8),
              np.where((width > p["flank_width"]) & adm, T, F).astype(np.int8)) == T
    idx = np.nonzero(sel)[0]
    is_flank = np.full(n, NA, np.int8)
    v = is_v.astype(np.int8)
    prev_v = _shift(v, idx, None, False, NA).astype(np.int8)
    next_v = _shift(v, idx, None, True, NA).astype(np.int8)
    eqT = lambda a: np.where(a == NA, NA, (a == T)).astype(np.int8)
    neT = lambda a: np.where(a == NA, NA, (a != T)).astype(np.int8)
    is_flank[idx] = _and(neT(v[idx]), _or(eqT(prev_v), eqT(next_v)))
    # :80-82  rows where is.V == T | is.flank == T, by condition
    sel = _or(eqT(is_v), eqT(is_flank)) == T
    idx = np.nonzero(sel)[0]
    flank_dist, flank_beta = np.full(n, np.nan), np.full(n, np.nan)
    is_sep = np.full(n, NA, np.int8)
    lead_start, lag_end = _shift(start, idx, cond, True, np.nan), _shift(end, idx, cond, False, np.nan)
    a1, a2 = lead_start - end[idx], start[idx] - lag_end
    fd = np.maximum(a1, a2)                                   # pmax: NA if either is NA (numpy propagates NaN)
    flank_dist[idx] = fd
    lead_beta, lag_beta = _shift(beta, idx, cond, True, np.nan), _shift(beta, idx, cond, False, np.nan)
    flank_beta[idx] = np.minimum(lead_beta, lag_beta)
    m = np.minimum(np.abs(beta[idx] - lag_beta), np.abs(beta[idx] - lead_beta))
    is_sep[idx] = np.where(np.isnan(m), NA, m > 2 * std[idx]).astype(np.int8)
    # :83-86
    category = np.zeros(n, np.int8)
    with np.errstate(invalid="ignore"):
        lmr = (is_v == T) & (is_sep == T) & (flank_dist < p["flank_dist_max"]) & (beta < p["lmr_max_beta"])
        category[lmr] = 1
        category[(category == 1) & (ncpg < p["flank_num_cpgs"]) & (flank_beta <= p["flank_beta_min"])] = 0
        low = (beta < p["umr_max_beta"]) & (std < p["umr_std_threshold"])
        category[low & (category == 1)] = 2
        category[low & (is_v == T) & (width > p["umr_large_width"]) & (ncpg / width > p["umr_large_density"])] = 2
    return dict(is_admissible=adm.astype(np.int8), is_V=is_v, is_flank=is_flank, flank_dist=flank_dist,
                flank_beta=flank_beta, is_sep=is_sep, category=category)
