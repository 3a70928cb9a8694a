"""wilcox.test / p.adjust('fdr') of R/call.R:315-317, :360: the oracle restatement is pinned against scipy's
independent implementations (CPU); the CUDA kernels are compared with the oracle (GPU).

R itself is not in the image.  scipy.stats.mannwhitneyu implements the same test (U of the first sample = R's W;
exact distribution without ties, normal approximation with continuity and tie correction) and
scipy.stats.false_discovery_control the same Benjamini-Hochberg adjustment; two known answers of R's documentation
examples are included as literals."""
import numpy as np
import pytest
from scipy import stats

TOL = 1e-12          # oracle vs CUDA / scipy: p-values (north_star asks 1e-6 absolute)


def groups(rng, n_groups, kind):
    xs, ys = [], []
    for g in range(n_groups):
        if kind == "exact":           # continuous values: no ties, both samples < 50
            nx, ny = rng.integers(1, 50, 2)
            x, y = rng.normal(0.0, 1.0, nx), rng.normal(rng.choice([0.0, 0.8, -1.5]), 1.0, ny)
        elif kind == "ties":          # methylation-like ratios: ties everywhere
            nx, ny = rng.integers(1, 80, 2)
            cx, cy = rng.integers(1, 12, nx), rng.integers(1, 12, ny)
            x, y = rng.binomial(cx, 0.8) / cx, rng.binomial(cy, rng.choice([0.8, 0.4])) / cy
        else:                         # large samples without ties: normal approximation by size
            nx, ny = rng.integers(50, 400, 2)
            x, y = rng.random(nx), rng.random(ny) + rng.choice([0.0, 0.1])
        xs.append(x)
        ys.append(y)
    return xs, ys


def csr(parts):
    ptr = np.zeros(len(parts) + 1, np.int64)
    ptr[1:] = np.cumsum([p.size for p in parts])
    return ptr, (np.concatenate(parts) if parts else np.zeros(0))


def scipy_p(x, y):
    x, y = x[~np.isnan(x)], y[~np.isnan(y)]
    ties = np.unique(np.concatenate([x, y])).size != x.size + y.size
    exact = x.size < 50 and y.size < 50 and not ties
    r = stats.mannwhitneyu(x, y, alternative="two-sided", method="exact" if exact else "asymptotic",
                           use_continuity=True)
    return r.statistic, r.pvalue


@pytest.mark.parametrize("kind", ["exact", "ties", "large"])
def test_oracle_wilcoxon_against_scipy(oracle, kind):
    rng = np.random.default_rng({"exact": 1, "ties": 2, "large": 3}[kind])
    xs, ys = groups(rng, 60, kind)
    for x, y in zip(xs, ys):
        w, p = oracle.wilcox_test(x, y)
        if np.unique(np.concatenate([x, y])).size == 1:
            assert np.isnan(p)                 # sigma = 0: R returns NaN (R/call.R:317 then sets 1)
            continue
        ws, ps = scipy_p(x, y)
        assert w == ws
        assert abs(p - ps) < TOL, (kind, x.size, y.size, p, ps)


def test_oracle_wilcoxon_known_answers(oracle):
    # R: wilcox.test(c(0.80, 0.83, 1.89, 1.04, 1.45, 1.38, 1.91, 1.64, 0.73, 1.46),
    #                c(1.15, 0.88, 0.90, 0.74, 1.21))  ->  W = 35, p-value = 0.2544  (exact)
    x = np.array([0.80, 0.83, 1.89, 1.04, 1.45, 1.38, 1.91, 1.64, 0.73, 1.46])
    y = np.array([1.15, 0.88, 0.90, 0.74, 1.21])
    w, p = oracle.wilcox_test(x, y)
    assert w == 35.0 and abs(p - 0.2544) < 5e-5
    # complete separation, 3 vs 3: p = 2 / choose(6, 3) = 0.1
    w, p = oracle.wilcox_test([1.0, 2.0, 3.0], [4.0, 5.0, 6.0])
    assert w == 0.0 and abs(p - 0.1) < 1e-15
    # NaN values are dropped; an empty sample has no p-value
    w, p = oracle.wilcox_test([np.nan, 1.0, 2.0, 3.0], [4.0, np.nan, 5.0, 6.0])
    assert w == 0.0 and abs(p - 0.1) < 1e-15
    assert np.isnan(oracle.wilcox_test([np.nan], [1.0, 2.0])[1])


def test_oracle_fdr_against_scipy(oracle):
    rng = np.random.default_rng(7)
    for n in (1, 2, 17, 1000):
        p = rng.random(n) ** 3
        p[rng.random(n) < 0.2] = 1.0                      # ties at 1 as R/call.R:317 produces them
        q = oracle.p_adjust_fdr(p)
        assert np.max(np.abs(q - stats.false_discovery_control(p, method="bh"))) < 1e-15
    p = np.array([0.01, np.nan, 0.04, 0.03, np.nan, 0.5])
    q = oracle.p_adjust_fdr(p)
    ok = ~np.isnan(p)
    assert np.all(np.isnan(q[~ok]))
    assert np.allclose(q[ok], stats.false_discovery_control(p[ok], method="bh"), rtol=0, atol=1e-15)
    # R: p.adjust(c(0.01, 0.02, 0.03, 0.04, 0.05), 'fdr') -> 0.05 0.05 0.05 0.05 0.05
    assert np.allclose(oracle.p_adjust_fdr([0.01, 0.02, 0.03, 0.04, 0.05]), 0.05, rtol=0, atol=1e-17)


# ---- GPU ------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["exact", "ties", "large"])
def test_gpu_wilcoxon_matches_oracle(engine, oracle, kind):
    rng = np.random.default_rng({"exact": 11, "ties": 12, "large": 13}[kind])
    xs, ys = groups(rng, 300, kind)
    if kind == "ties":
        xs[3], ys[3] = np.full(5, 0.5), np.full(4, 0.5)           # all equal: NaN
        xs[4] = np.array([np.nan, 0.25, 0.5])                    # NaN dropped
        xs[5], ys[5] = np.zeros(0), np.array([0.1, 0.2])           # empty sample
    if kind == "exact":
        xs[0], ys[0] = rng.normal(size=49), rng.normal(size=49) + 0.3   # the largest exact table
        xs[1], ys[1] = np.array([1.0]), np.array([2.0])
    xp, x = csr(xs)
    yp, y = csr(ys)
    p, w = engine.wilcoxon_groups(xp, x, yp, y, want_stat=True)
    for g, (a, b) in enumerate(zip(xs, ys)):
        wo, po = oracle.wilcox_test(a, b)
        if np.isnan(po):
            assert np.isnan(p[g]), g
        else:
            assert w[g] == wo, g
            assert abs(p[g] - po) < TOL, (kind, g, a.size, b.size, p[g], po)


@pytest.mark.gpu
def test_gpu_wilcoxon_one_large_group(engine, oracle):
    rng = np.random.default_rng(5)
    x = np.round(rng.random(30_000), 3)
    y = np.round(rng.random(20_000) * 1.01, 3)
    p, w = engine.wilcoxon_groups([0, x.size], x, [0, y.size], y, want_stat=True)
    wo, po = oracle.wilcox_test(x, y)
    assert w[0] == wo and abs(p[0] - po) < TOL


@pytest.mark.gpu
def test_gpu_fdr_matches_oracle(engine, oracle):
    rng = np.random.default_rng(9)
    for n in (1, 5, 1000, 200_000):
        p = rng.random(n) ** 2
        p[rng.random(n) < 0.3] = 1.0
        p[rng.random(n) < 0.01] = np.nan
        q, qo = engine.p_adjust_fdr(p), oracle.p_adjust_fdr(p)
        assert np.array_equal(np.isnan(q), np.isnan(qo))
        ok = ~np.isnan(qo)
        assert np.array_equal(q[ok], qo[ok])            # same operations in the same order: bit-equal
