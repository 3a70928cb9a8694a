"""The DMR caller of R/call.R:254-365 (call_differences): CPU restatement (oracle/diffcall.py) self-checks, and the
resident CUDA path against it (GPU)."""
import numpy as np
import pytest

from methylasso_b200 import synth

INT_COLS = ("condition", "estimate_id", "segment_id", "start", "end", "num_cpgs", "num_cpgs_ref")
EXACT_COLS = ("width", "coverage", "coverage_ref", "pseudoweight")          # integer-valued doubles
FLOAT_COLS = ("beta", "std", "beta_ref", "std_ref", "diff", "cohen_d", "pval", "coverage_score")


def thin(t, rng, drop=0.15):
    """drop rows at random so that replicates / conditions miss positions (validation needs >= 1 row per dataset)"""
    keep = rng.random(t["pos"].size) > drop
    for d in np.unique(t["dataset"]):
        keep[np.nonzero(t["dataset"] == d)[0][0]] = True
    return {k: np.asarray(v)[keep] for k, v in t.items()}


def make_tables(seed, sizes=(6000, 2500), reps=(3, 3), drop=0.15):
    rng = np.random.default_rng(seed)
    return [thin(synth.chromosome_table(n, seed * 10 + i, reps), rng, drop) for i, n in enumerate(sizes)]


def oracle_tables(O, DC, tables, ref, **kw):
    out = []
    for t in tables:
        o = O.detect(t, 0, lambda2=25.0)
        est = O.fast_diff_combine(o["estimates"], o["n_params"], o["n_est"])
        out.append(DC.call_differences_singlechr(t, est, ref, **kw))
    return out


def test_oracle_diffcall_invariants(oracle):
    from oracle import diffcall as DC
    tables = make_tables(3, sizes=(4000,))
    g = oracle_tables(oracle, DC, tables, ref=2)[0]
    n = g["start"].size
    assert n > 10
    assert np.array_equal(g["segment_id"], np.arange(1, n + 1))
    assert np.all(g["start"][1:] >= g["start"][:-1]) and np.all(g["end"] >= g["start"] + 2)
    ok = ~np.isnan(g["beta"]) & ~np.isnan(g["beta_ref"])
    assert np.allclose(g["diff"][ok], g["beta"][ok] - g["beta_ref"][ok], rtol=0, atol=0)
    assert np.all((g["pval"] >= 0) & (g["pval"] <= 1))
    assert np.all((g["coverage_score"] > 0) & (g["coverage_score"] <= 100))
    # every position of the condition is counted at least once (boundary positions at start - 1 twice)
    cond_pos = np.unique(tables[0]["pos"][tables[0]["condition"] == 1]).size
    assert cond_pos <= g["num_cpgs"].sum() <= cond_pos + n
    # swapping the reference swaps the sides
    h = oracle_tables(oracle, DC, tables, ref=1)[0]
    assert h["condition"][0] == 2 and np.isclose(np.nansum(h["coverage"]), np.nansum(g["coverage_ref"]))
    t, keep = DC.call_differences([g])
    assert keep.sum() >= 1 and np.all(np.abs(t["diff"][keep]) >= 0.1) and np.all(t["fdr"] >= t["pval"] - 1e-15)


def compare(got, want):
    m = got["job"].size
    assert m == sum(w["start"].size for w in want)
    for j, w in enumerate(want):
        sel = got["job"] == j
        for c in INT_COLS:
            assert np.array_equal(np.asarray(got[c][sel], np.int64), w[c]), (j, c)
        for c in EXACT_COLS:
            assert np.array_equal(got[c][sel], w[c]), (j, c)
        for c in FLOAT_COLS:
            a, b = got[c][sel], w[c]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (j, c)
            ok = ~np.isnan(b)
            tol = 1e-9 if c == "cohen_d" else 1e-12
            assert np.all(np.abs(a[ok] - b[ok]) <= tol * np.maximum(1.0, np.abs(b[ok]))), (j, c, np.max(np.abs(a[ok] - b[ok])))


@pytest.mark.gpu
@pytest.mark.parametrize("ref,reps,drop", [(2, (3, 3), 0.15), (1, (2, 3), 0.3), (2, (1, 1), 0.0)])
def test_gpu_call_differences_matches_oracle(engine, oracle, ref, reps, drop):
    from methylasso_b200 import api
    from methylasso_b200._native import MlParams
    from oracle import diffcall as DC
    tables = make_tables(11 + ref, reps=reps, drop=drop)
    want = oracle_tables(oracle, DC, tables, ref)
    ptr, cols = synth.concat_jobs(tables)
    batch = engine.upload(ptr, cols)
    res = engine.detect(batch, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        with pytest.raises(api.MethyLassoError):
            res.call_differences(batch, ref)            # not a difference result yet (R/call.R:256)
        res.fast_diff_combine()
        assert res.call_differences(batch, ref, fetch=False) == sum(w["start"].size for w in want)
        compare(res.call_differences(batch, ref), want)
        # other thresholds regroup the same parts
        want2 = oracle_tables(oracle, DC, tables, ref, min_diff=0.05, min_delta_diff=0.2, max_distance=300.0)
        compare(res.call_differences(batch, ref, min_diff=0.05, min_delta_diff=0.2, max_distance=300.0), want2)
    finally:
        res.free()
        batch.free()


@pytest.mark.gpu
def test_gpu_call_differences_three_conditions(engine, oracle):
    from methylasso_b200._native import MlParams
    from oracle import diffcall as DC
    tables = make_tables(21, sizes=(3000,), reps=(2, 2, 2), drop=0.25)
    ref = 2
    want = oracle_tables(oracle, DC, tables, ref)
    assert set(np.unique(want[0]["estimate_id"])) == {2, 3} and set(np.unique(want[0]["condition"])) == {1, 3}
    ptr, cols = synth.concat_jobs(tables)
    batch = engine.upload(ptr, cols)
    res = engine.detect(batch, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        res.fast_diff_combine()
        compare(res.call_differences(batch, ref), want)
    finally:
        res.free()
        batch.free()


@pytest.mark.gpu
def test_gpu_call_differences_genome_api(engine, oracle):
    """R-level mirror: difference_detection_fast(keep_resident=True) -> call_differences, FDR over all chromosomes."""
    from methylasso_b200 import api
    from oracle import diffcall as DC
    tables = make_tables(31, sizes=(5000, 3000, 1500))
    data = {k: np.concatenate([t[k] for t in tables]) for k in tables[0]}
    data["chr"] = np.concatenate([np.full(t["pos"].size, f"chr{i + 1}") for i, t in enumerate(tables)])
    ret = api.difference_detection_fast(data, ref=2, lambda2=25, engine=engine, keep_resident=True)
    try:
        got = api.call_differences(data, ret, return_all=True)
        want, keep = DC.call_differences(oracle_tables(oracle, DC, tables, 2))
        assert np.array_equal(got["chr"], np.array([f"chr{j + 1}" for j in want["job"]]))
        assert np.max(np.abs(got["fdr"] - want["fdr"])) < 1e-12
        assert np.array_equal(got["called"], keep) and keep.sum() >= 1
        called = api.call_differences(data, ret)
        assert called["start"].size == keep.sum()
    finally:
        ret["_resident"].free()
