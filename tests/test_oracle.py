"""CPU tests of the oracle itself: pinned to the reference's own DP kernel and known answers."""
import os

import numpy as np
import pytest

from methylasso_b200 import synth
from conftest import pooled_series

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_known_answers(oracle):
    # produced by the reference DP
    assert np.allclose(oracle.weighted_1d_flsa([0, 0, 1, 1], [1, 1, 1, 1], 0.5), [.25, .25, .75, .75], atol=1e-15)
    assert np.allclose(oracle.weighted_1d_flsa([0, 0, 1, 1], [1, 1, 1, 1], 10), [.5, .5, .5, .5], atol=1e-15)


def test_dp_port_matches_golden_reference_vectors(oracle):
    g = np.load(os.path.join(GOLD, "dp_vectors.npz"))
    for k in ("random", "zero_w", "monotone", "plateaus", "known"):
        beta, _, _ = oracle.tf_dp_weight(g[f"{k}_y"], g[f"{k}_w"], float(g[f"{k}_lam"]), "port")
        assert np.array_equal(beta, g[f"{k}_beta"], equal_nan=True), k


def test_dp_port_bit_equal_to_verbatim_reference(oracle):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(1)
    for trial in range(200):
        n = int(rng.integers(1, 600))
        y = rng.normal(size=n)
        w = np.where(rng.random(n) < 0.2, 0.0, rng.exponential(size=n))
        lam = float(rng.choice([0.0, 1e-3, 0.3, 5.0, 100.0]))
        a, tma, tpa = oracle.tf_dp_weight(y, w, lam, "port")
        b, tmb, tpb = oracle.tf_dp_weight(y, w, lam, "reference")
        assert np.array_equal(a, b, equal_nan=True)
        assert np.array_equal(tma, tmb, equal_nan=True) and np.array_equal(tpa, tpb, equal_nan=True)


def kkt_residual(y, w, beta, lam):
    """Optimality of the fused-lasso solution: c_k = sum_{i<=k} w_i (y_i - beta_i) satisfies
    |c_k| <= lam, c_k = -lam * sign(beta_{k+1} - beta_k) at jumps, c_n = 0.
    Returns the largest violation relative to lam."""
    c = np.cumsum(w * (y - beta))
    viol = max(0.0, np.max(np.abs(c[:-1])) - lam) if c.size > 1 else 0.0
    jump = np.sign(np.diff(beta))
    at = jump != 0
    if at.any():
        viol = max(viol, np.max(np.abs(c[:-1][at] + lam * jump[at])))
    viol = max(viol, abs(c[-1]))
    return viol / lam


def test_dp_kkt_property(oracle):
    t = synth.chromosome_table(50_000, 3, (1,))
    y, w = pooled_series(t)
    for lam in (10.0, 25.0, 1000.0):
        beta, _, _ = oracle.tf_dp_weight(y, w, lam, "port")
        assert kkt_residual(y, w, beta, lam) < 1e-9


def _check_fixture(g, prefix, r, seg):
    assert np.array_equal(r["estimates"]["estimate_id"], g[f"{prefix}_est_estimate_id"])
    assert np.array_equal(r["estimates"]["start"], g[f"{prefix}_est_start"])
    for c in ("beta", "betahat", "pseudoweight"):
        assert np.array_equal(r["estimates"][c], g[f"{prefix}_est_{c}"]), (prefix, c)
    assert np.array_equal(r["mu"], g[f"{prefix}_mu"])
    assert np.array_equal(r["offset"], g[f"{prefix}_offset"])
    for c in seg:
        assert np.array_equal(seg[c], g[f"{prefix}_seg_{c}"], equal_nan=True), (prefix, c)


def fixture_table(g, prefix):
    return {c: g[f"{prefix}_in_{c}"] for c in ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage")}


FIXTURES = {
    "fast_1x1": dict(model=0, lambda2=25.0),
    "fast_1x3": dict(model=0, lambda2=25.0),
    "fast_2x2": dict(model=0, lambda2=25.0),
    "fast_iter3": dict(model=0, lambda2=10.0, max_iter=3),
    "full_sig": dict(model=1, lambda2=25.0, max_iter=30, nouter=20, hyper=100.0, bf_per_kb=50.0, max_lambda2=1000.0),
    "full_diff": dict(model=2, lambda2=25.0, max_iter=30, nouter=20, hyper=100.0, bf_per_kb=50.0, max_lambda2=1000.0, ref=1),
}


@pytest.mark.parametrize("prefix", list(FIXTURES))
def test_port_reproduces_path_fixtures(oracle, prefix):
    """Fixtures were generated with the reference's verbatim DP inside the restatement; the pure
    port (own DP restatement) must reproduce them bit for bit."""
    g = np.load(os.path.join(GOLD, "path_fixtures.npz"))
    oracle.use_reference_dp(False)
    t = fixture_table(g, prefix)
    r = oracle.detect(t, **FIXTURES[prefix])
    seg = oracle.segment_methylation_helper(r["estimates"], 0.01)
    _check_fixture(g, prefix, r, seg)
    meta = g[f"{prefix}_meta"]
    assert [r["n_params"], r["n_est"], r["n_dsets"], int(r["converged"]), r["irls_steps"], r["dampings"],
            r["outer_iters"]] == meta.tolist()


def test_fast_semantics(oracle):
    """betahat = pooled methylation, pseudoweight = pooled coverage, beta centred,
    offset = per-dataset weighted mean, mu = beta[param] + offset."""
    t = synth.chromosome_table(4000, 5, (3,))
    r = oracle.detect(t, 0, lambda2=25.0)
    y, w = pooled_series(t)
    est = r["estimates"]
    assert np.allclose(est["pseudoweight"], w, rtol=0, atol=1e-9)
    assert np.allclose(est["betahat"], y, rtol=1e-13)
    assert abs(np.sum(est["beta"] * est["pseudoweight"]) / np.sum(est["pseudoweight"])) < 1e-12
    for d in (1, 2, 3):
        m = t["dataset"] == d
        assert np.isclose(r["offset"][d - 1], t["Nmeth"][m].sum() / t["coverage"][m].sum(), rtol=1e-12)
    u, inv = np.unique(t["pos"], return_inverse=True)
    assert np.allclose(r["mu"], est["beta"][inv] + r["offset"][t["dataset"] - 1], rtol=1e-13, atol=1e-15)
    assert r["converged"] and r["irls_steps"] == 1


def test_segment_helper_drift(oracle):
    """A.6: comparison is against the beta at the START of the segment, strict >."""
    beta = np.array([0.0, 0.006, 0.012, 0.018, 0.019, 0.5, 0.5])
    est = dict(estimate_id=np.ones(7, np.int32), start=np.arange(10, 80, 10, dtype=np.int32), beta=beta,
               betahat=beta + 0.1, pseudoweight=np.arange(1, 8, dtype=float))
    s = oracle.segment_methylation_helper(est, 0.01)
    # 0.012 - 0 > 0.01 opens a segment at row 2; then 0.018, 0.019 stay (diff to 0.012 <= 0.01); 0.5 opens
    assert s["start"].tolist() == [10, 30, 60]
    assert s["end"].tolist() == [29, 59, 70]
    assert s["segment_id"].tolist() == [1, 2, 3]
    assert np.allclose(s["pseudoweight"], [3, 12, 13])
    # estimate change closes at start(i-1) and restarts ids
    est2 = dict(estimate_id=np.array([1, 1, 2, 2], np.int32), start=np.array([5, 9, 5, 9], np.int32),
                beta=np.zeros(4), betahat=np.ones(4), pseudoweight=np.ones(4))
    s2 = oracle.segment_methylation_helper(est2, 0.01)
    assert s2["estimate_id"].tolist() == [1, 2] and s2["segment_id"].tolist() == [1, 1]
    assert s2["start"].tolist() == [5, 5] and s2["end"].tolist() == [9, 9]


def _tbl(**kw):
    base = dict(dataset=[1, 1, 1], condition=[1, 1, 1], replicate=[1, 1, 1], pos=[10, 20, 30], Nmeth=[1, 2, 3],
                coverage=[5, 5, 5])
    base.update(kw)
    return {k: np.asarray(v, np.int32) for k, v in base.items()}


VALIDATION_CASES = [
    (_tbl(dataset=[1], condition=[1], replicate=[1], pos=[1], Nmeth=[1], coverage=[2]), 1),
    (_tbl(dataset=[2, 2, 2]), 3),
    (_tbl(condition=[2, 2, 2]), 4),
    (_tbl(replicate=[2, 2, 2]), 5),
    (_tbl(pos=[10, 10, 30]), 6),
    (_tbl(pos=[10, 30, 20]), 6),
    (_tbl(dataset=[1, 3, 3]), 7),
    (_tbl(dataset=[1, 2, 2], replicate=[1, 3, 3]), 8),
    (_tbl(dataset=[1, 2, 2], condition=[1, 3, 3]), 9),
    (_tbl(dataset=[1, 2, 2], condition=[1, 2, 2], replicate=[1, 2, 2]), 10),
    (_tbl(Nmeth=[1, -2, 3]), 11),
    (_tbl(coverage=[5, 1, 5]), 12),
    # first offence wins: unsorted at row 1 beats negative count at row 0 (different stages)
    (_tbl(pos=[10, 5, 30], Nmeth=[-1, 2, 3]), 6),
    # within a dataset condition/replicate are never looked at
    (_tbl(condition=[1, 7, 1], replicate=[1, 9, 1]), 0),
]


@pytest.mark.parametrize("case", range(len(VALIDATION_CASES)))
def test_check_observations_codes(oracle, case):
    t, code = VALIDATION_CASES[case]
    assert oracle.check_observations(t) == code


def test_detect_error_sites(oracle):
    t = synth.chromosome_table(500, 1, (1,))
    with pytest.raises(oracle.OracleError) as e:
        oracle.detect(t, 0, lambda2=0.0)
    assert e.value.code == 13
    with pytest.raises(oracle.OracleError) as e:
        oracle.detect(t, 2, lambda2=25.0, ref=1)
    assert e.value.code == 14
    t2 = synth.chromosome_table(500, 1, (1, 1))
    with pytest.raises(oracle.OracleError) as e:
        oracle.detect(t2, 2, lambda2=25.0, ref=3)
    assert e.value.code == 15
    same = _tbl(dataset=[1, 2], condition=[1, 1], replicate=[1, 2], pos=[7, 7], Nmeth=[1, 1], coverage=[2, 2])
    with pytest.raises(oracle.OracleError) as e:
        oracle.detect(same, 0, lambda2=25.0)
    assert e.value.code == 18


def test_even_bin_rule_matches_bruteforce(oracle):
    """A.3: bin = (first boundary strictly greater than pos) - 1 with boundaries lower + i*size/Nbins."""
    t = synth.chromosome_table(3000, 21, (1,))
    r = oracle.detect(t, 1, lambda2=25.0, max_iter=1, nouter=1, bf_per_kb=50.0)
    pos = t["pos"].astype(float)
    lo, hi = pos.min(), pos.max()
    nb = int((hi - lo + 1) * 50.0 / 1000.0)
    assert r["n_params"] == nb
    bounds = lo + np.arange(nb + 1) * (hi - lo) / float(nb)
    assert np.array_equal(r["estimates"]["start"], np.trunc(bounds[:-1]))
    bins = np.searchsorted(bounds, pos, side="right") - 1
    bins[bins >= nb] = nb - 1
    cov = np.bincount(bins, t["coverage"], nb)
    # with a single dataset and beta-binomial weights, a bin has weight > 0 iff it holds data
    assert np.array_equal(r["estimates"]["pseudoweight"] > 0, cov > 0)
