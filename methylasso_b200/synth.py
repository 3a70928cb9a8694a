"""Synthetic WGBS data of human-genome shape.

The reference ships no data (its README example points at external URLs, README.md:129-135), so
every test and benchmark input is generated here from `numpy.random.default_rng(seed)`.  The
tables have the columns the reference's R layer hands to the Rcpp module
(`MethyLasso.R:246-297`): dataset, condition, replicate, pos, Nmeth, coverage — all int32, rows
of one chromosome sorted by (dataset, pos), coverage filtered at `>= min_cov`
(`MethyLasso.R:361`).
"""
from __future__ import annotations

import numpy as np

# hg38 chromosome lengths in Mb (chr1..22, X, Y)
HG38_MB = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80,
           135.09, 133.28, 114.36, 107.04, 101.99, 90.34, 83.26, 80.37, 58.62, 64.44,
           46.71, 50.82, 156.04, 57.23]
CHR_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]
GENOME_CPGS = 28_000_000
BASE_SEED = 20240


def chromosome_cpgs(total: int = GENOME_CPGS) -> list[int]:
    """CpGs per chromosome, proportional to length (chr1 ~ 2.26 M of 28 M)."""
    tot_mb = sum(HG38_MB)
    n = [int(round(total * mb / tot_mb)) for mb in HG38_MB]
    n[0] += total - sum(n)
    return n


def _runs(rng, n, mean_a, mean_b):
    """Alternating run lengths (state 0 then 1 ...) with geometric means, covering n items."""
    est = int(n / (mean_a + mean_b) * 1.3) + 16
    la = rng.geometric(1.0 / mean_a, est)
    lb = rng.geometric(1.0 / mean_b, est)
    lens = np.empty(2 * est, dtype=np.int64)
    lens[0::2] = la
    lens[1::2] = lb
    cs = np.cumsum(lens)
    k = int(np.searchsorted(cs, n)) + 1
    lens = lens[:k].copy()
    lens[-1] -= cs[k - 1] - n
    state = (np.arange(k) & 1).astype(np.int8)
    if lens[-1] == 0:
        lens, state = lens[:-1], state[:-1]
    return lens, state


def truth(n_cpgs: int, seed: int, dmv: bool = False):
    """Positions (strictly increasing int32) and true methylation per CpG.  dmv=True adds, from a random stream of its
    own (everything else is unchanged): unmethylated stretches of 6-15 kb, about one per 150 000 CpGs - wide enough for
    the reference's valley rule (R/call.R:100-133, pmd_valley_min_width = 5000) - and low-methylated regions inside
    the PMD blocks (R/call.R:201-212 drops LMRs that lie in PMDs)."""
    rng = np.random.default_rng(seed)
    # background runs (mean 340 CpGs) alternate with CpG-island clusters (mean 60): ~15 % in islands
    lens, state = _runs(rng, n_cpgs, 340.0, 60.0)
    island = np.repeat(state, lens).astype(bool)
    gaps = np.where(island, rng.geometric(1.0 / 12.0, n_cpgs), rng.geometric(1.0 / 125.0, n_cpgs))
    pos = np.cumsum(gaps.astype(np.int64)) + 10_000
    if pos[-1] >= 2**31 - 1:
        raise ValueError("synthetic chromosome too long for int32 positions")
    # background level: piecewise-constant regimes of ~800 CpGs in [0.80, 0.92]
    blens, _ = _runs(rng, n_cpgs, 800.0, 800.0)
    p = np.repeat(rng.uniform(0.80, 0.92, blens.size), blens)
    # UMRs: islands of 40..300 CpGs become unmethylated with probability 0.7
    run_id = np.repeat(np.arange(lens.size), lens)
    umr_run = (state == 1) & (lens >= 40) & (lens <= 300) & (rng.random(lens.size) < 0.7)
    umr_level = rng.uniform(0.01, 0.06, lens.size)
    is_umr = umr_run[run_id]
    p = np.where(is_umr, umr_level[run_id], p)
    # LMRs: 5..40 CpGs at [0.15, 0.45], about one per 3000 CpGs, outside UMRs
    n_lmr = max(1, n_cpgs // 3000)
    st = rng.integers(0, max(1, n_cpgs - 41), n_lmr)
    ln = rng.integers(5, 41, n_lmr)
    lv = rng.uniform(0.15, 0.45, n_lmr)
    for s, l, v in zip(st, ln, lv):
        sl = slice(int(s), int(s + l))
        p[sl] = np.where(is_umr[sl], p[sl], v)
    # PMD blocks: ~5 % of the chromosome, 100 kb..1 Mb, per-CpG p ~ Beta(mean .6, sd .2)
    span = int(pos[-1] - pos[0])
    target = 0.05 * span
    covered = 0
    m, sd = 0.6, 0.2
    k = m * (1 - m) / sd**2 - 1
    is_pmd = np.zeros(n_cpgs, dtype=bool)
    guard = 0
    while covered < target and guard < 10_000:
        guard += 1
        w = int(rng.integers(100_000, 1_000_001))
        if w >= span:
            break
        s = int(pos[0]) + int(rng.integers(0, span - w))
        i0, i1 = np.searchsorted(pos, [s, s + w])
        if i1 - i0 < 10:
            continue
        sl = slice(int(i0), int(i1))
        vals = rng.beta(m * k, (1 - m) * k, i1 - i0)
        p[sl] = np.where(is_umr[sl], p[sl], vals)
        is_pmd[sl] = True
        covered += w
        if dmv and i1 - i0 > 200 and rng.random() < 0.7:          # an LMR inside the block
            a = int(i0) + int(rng.integers(20, i1 - i0 - 60))
            p[a:a + int(rng.integers(8, 41))] = rng.uniform(0.12, 0.3)
    if dmv:
        rng2 = np.random.default_rng(seed * 31 + 7)
        for _ in range(max(1, n_cpgs // 150_000)):
            i = int(rng2.integers(0, max(1, n_cpgs - 2)))
            j = int(np.searchsorted(pos, pos[i] + int(rng2.integers(6_000, 15_001))))
            if j - i >= 8 and j < n_cpgs:
                p[i:j] = rng2.uniform(0.01, 0.05)
                is_umr[i:j] = True
    return pos.astype(np.int32), np.clip(p, 0.0, 1.0), is_umr, is_pmd


def sample_replicate(pos, p, seed, mean_cov=20.0, min_cov=5):
    rng = np.random.default_rng(seed)
    cov = rng.poisson(mean_cov, pos.size)
    nmeth = rng.binomial(cov, p)
    keep = cov >= min_cov
    return pos[keep], nmeth[keep].astype(np.int32), cov[keep].astype(np.int32)


def shift_condition(p, seed, n_regions=None):
    """Condition 2 = condition 1 with regions of 10..150 CpGs shifted by +-(0.2..0.5) (config 3)."""
    rng = np.random.default_rng(seed)
    n = p.size
    if n_regions is None:
        n_regions = max(1, int(round(5000 * n / GENOME_CPGS)))
    q = p.copy()
    st = rng.integers(0, max(1, n - 151), n_regions)
    ln = rng.integers(10, 151, n_regions)
    dl = rng.uniform(0.2, 0.5, n_regions) * rng.choice([-1.0, 1.0], n_regions)
    truth_regions = []
    for s, l, d in zip(st, ln, dl):
        q[int(s):int(s + l)] = np.clip(q[int(s):int(s + l)] + d, 0.0, 1.0)
        truth_regions.append((int(s), int(s + l), float(d)))
    return q, truth_regions


def chromosome_table(n_cpgs, seed, reps_per_condition=(1,), mean_cov=20.0, min_cov=5, dmv=False):
    """One chromosome's input table: dict of int32 columns sorted by (dataset, pos)."""
    pos, p, _, _ = truth(n_cpgs, seed, dmv)
    cols = {k: [] for k in ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage")}
    dset = 0
    for c, nrep in enumerate(reps_per_condition, start=1):
        pc = p if c == 1 else shift_condition(p, seed * 7 + c)[0]
        for r in range(1, nrep + 1):
            dset += 1
            ps, nm, cv = sample_replicate(pos, pc, seed * 100 + dset, mean_cov, min_cov)
            cols["dataset"].append(np.full(ps.size, dset, np.int32))
            cols["condition"].append(np.full(ps.size, c, np.int32))
            cols["replicate"].append(np.full(ps.size, r, np.int32))
            cols["pos"].append(ps)
            cols["Nmeth"].append(nm)
            cols["coverage"].append(cv)
    return {k: np.concatenate(v) for k, v in cols.items()}


def genome_tables(config_index=2, total_cpgs=GENOME_CPGS, reps_per_condition=(3,), n_chr=24,
                  mean_cov=20.0, seed=None, dmv=False):
    """List of per-chromosome tables for a whole synthetic genome (config 2/3 shapes)."""
    seed = BASE_SEED + config_index if seed is None else seed
    sizes = chromosome_cpgs(total_cpgs)[:n_chr]
    return [chromosome_table(n, seed * 50 + i, reps_per_condition, mean_cov, dmv=dmv)
            for i, n in enumerate(sizes)]


def concat_jobs(tables):
    """Concatenate per-chromosome tables into one SoA batch + CSR job offsets (int64)."""
    ptr = np.zeros(len(tables) + 1, np.int64)
    for i, t in enumerate(tables):
        ptr[i + 1] = ptr[i] + t["pos"].size
    out = {k: np.ascontiguousarray(np.concatenate([t[k] for t in tables]))
           for k in ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage")}
    return ptr, out
