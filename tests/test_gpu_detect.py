"""GPU parity of the batched methylation_detection engine (through the C ABI) against the oracle."""
import os

import numpy as np
import pytest

from methylasso_b200 import api, synth
from methylasso_b200._native import MlParams
from test_oracle import FIXTURES, VALIDATION_CASES, fixture_table, _tbl

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REL = 1e-10  # north_star: fitted betas within 1e-6 relative; asserted per element at 1e-10 (measured: ~1e-12 and below)


def params(model=0, lambda2=25.0, tol_val=0.01, max_iter=1, nouter=1, max_lambda2=None, lambda1=0.0,
           bf_per_kb=50.0, hyper=100.0, ref=1):
    return MlParams(lambda2, lambda2 + 1 if max_lambda2 is None else max_lambda2, lambda1, tol_val, bf_per_kb,
                    hyper, max_iter, nouter, ref, model, 0, 0)


def run_one(engine, t, **kw):
    res = engine.detect_batch(np.array([0, t["pos"].size], np.int64), t, params(**kw))
    try:
        return res.job(0)
    finally:
        res.free()


def close(a, b, rel=REL):
    """per element: |a - b| <= rel * max(|b|, 1e-3)"""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return a.shape == b.shape and (a.size == 0 or bool(np.all(np.abs(a - b) <= rel * np.maximum(np.abs(b), 1e-3))))


def seg_row_counts(est_id, est_start, seg):
    """rows of the estimates table inside each segment (start <= row start <= end within the estimate)."""
    est_id, est_start = np.asarray(est_id, np.int64), np.asarray(est_start, np.float64)
    n = np.zeros(len(seg["start"]), np.int64)
    for e in np.unique(est_id):
        st = est_start[est_id == e]
        m = np.nonzero(np.asarray(seg["estimate_id"]) == e)[0]
        lo = np.searchsorted(st, np.asarray(seg["start"], np.float64)[m], "left")
        hi = np.searchsorted(st, np.asarray(seg["end"], np.float64)[m], "right")
        n[m] = hi - lo
    return n


def assert_job_matches(r, o, exact_fast=True):
    est, oest = r["estimates"], o["estimates"]
    assert np.array_equal(est["estimate_id"], oest["estimate_id"])
    assert np.array_equal(est["start"], oest["start"])
    assert close(est["pseudoweight"], oest["pseudoweight"])
    assert close(est["betahat"], oest["betahat"])
    assert close(est["beta"], oest["beta"])
    assert close(r["mu"], o["mu"]) and close(r["offset"], o["offset"])
    assert r["converged"] == o["converged"]
    # The number of dampings is not an output of the reference and, at a fixed point of the IRLS
    # map, `mlogp > old_mlogp` compares two sums that are equal up to the rounding of their
    # association order (Eigen's, the oracle's and the GPU's all differ): not asserted.
    assert (r["irls_steps"], r["outer_iters"]) == (o["irls_steps"], o["outer_iters"])
    # a15: the -log posterior of the last evaluation (Posterior.ipp:63-69) itself, not only its side effects
    if "last_mlogp" in o:
        assert abs(r["mlogp"] - o["last_mlogp"]) <= 1e-9 * max(1.0, abs(o["last_mlogp"])), (r["mlogp"], o["last_mlogp"])


@pytest.mark.parametrize("prefix", list(FIXTURES))
def test_golden_path_fixtures(engine, oracle, prefix):
    g = np.load(os.path.join(GOLD, "path_fixtures.npz"))
    t = fixture_table(g, prefix)
    r = run_one(engine, t, **FIXTURES[prefix])
    o = dict(mu=g[f"{prefix}_mu"], offset=g[f"{prefix}_offset"],
             estimates={c: g[f"{prefix}_est_{c}"] for c in ("estimate_id", "start", "beta", "betahat", "pseudoweight")})
    meta = g[f"{prefix}_meta"].tolist()
    o.update(converged=bool(meta[3]), irls_steps=meta[4], dampings=meta[5], outer_iters=meta[6])
    assert_job_matches(r, o)
    if prefix.startswith("fast") and FIXTURES[prefix].get("max_iter", 1) == 1:
        # count data: pseudodata and the FLSA fit before centring are exact; what remains is the
        # association order of two long sums (centring mean, offset) -> a uniform shift ~1e-16
        assert np.array_equal(r["estimates"]["pseudoweight"], o["estimates"]["pseudoweight"])
        assert np.array_equal(r["estimates"]["betahat"], o["estimates"]["betahat"])
        assert np.max(np.abs(r["estimates"]["beta"] - o["estimates"]["beta"])) < 1e-13
    # segment boundaries identical
    seg = engine.segment_methylation_helper(r["estimates"], 0.01)
    for c in ("estimate_id", "segment_id", "start", "end"):
        assert np.array_equal(seg[c], g[f"{prefix}_seg_{c}"]), c
    assert close(seg["beta"], g[f"{prefix}_seg_beta"]) and close(seg["pseudoweight"], g[f"{prefix}_seg_pseudoweight"])


REFERENCE_FIXTURES = ("fast_1x1", "fast_1x3", "fast_2x2", "fast_iter3", "fast_lam1000", "full_sig", "full_sig_3c",
                      "full_diff", "full_diff_ref2")


@pytest.mark.parametrize("name", REFERENCE_FIXTURES)
def test_reference_generated_fixtures(engine, name):
    """tests/golden/reference_fixtures.npz holds outputs of the reference's OWN C++ path (unmodified sources
    compiled against stand-in Rcpp/Eigen headers, tests/golden/make_golden.py): no restatement in between."""
    g = np.load(os.path.join(GOLD, "reference_fixtures.npz"))
    t = {c: g[f"{name}_in_{c}"] for c in ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage")}
    model, lambda2, max_iter, nouter, ref = g[f"{name}_kw"].tolist()
    kw = dict(model=int(model), lambda2=lambda2, max_iter=int(max_iter), ref=int(ref))
    if model != 0:
        kw.update(nouter=int(nouter), max_lambda2=1000.0, hyper=100.0, bf_per_kb=50.0)
    r = run_one(engine, t, **kw)
    est = r["estimates"]
    assert r["converged"] == bool(g[f"{name}_converged"])
    assert np.array_equal(est["estimate_id"], g[f"{name}_est_estimate_id"])
    assert np.array_equal(est["start"], g[f"{name}_est_start"])
    for c in ("beta", "betahat", "pseudoweight"):
        assert close(est[c], g[f"{name}_est_{c}"]), c
    assert close(r["mu"], g[f"{name}_mu"]) and close(r["offset"], g[f"{name}_offset"])
    if model == 0 and max_iter == 1:        # count data: bit-identical pseudodata, beta to the rounding of one mean
        assert np.array_equal(est["betahat"], g[f"{name}_est_betahat"])
        assert np.array_equal(est["pseudoweight"], g[f"{name}_est_pseudoweight"])
        assert np.max(np.abs(est["beta"] - g[f"{name}_est_beta"])) < 1e-13
    seg = engine.segment_methylation_helper({c: g[f"{name}_est_{c}"] for c in ("estimate_id", "start", "beta", "betahat",
                                                                               "pseudoweight")}, 0.01)
    for c in ("estimate_id", "segment_id", "start", "end"):
        assert np.array_equal(np.asarray(seg[c], np.int64), g[f"{name}_seg_{c}"].astype(np.int64)), c
    # segments of up to SEG_LONG = 256 rows: same additions in the same order, bit-identical.  Longer ones are summed
    # by a block in a fixed tree (k_seg_emit_long): beta (a copy) stays bit-identical, the sums agree to rounding.
    rows = seg_row_counts(g[f"{name}_est_estimate_id"], g[f"{name}_est_start"], seg)
    short = rows <= 256
    assert np.array_equal(seg["beta"], g[f"{name}_seg_beta"], equal_nan=True)
    for c in ("betahat", "pseudoweight", "stdev"):
        ref = g[f"{name}_seg_{c}"]
        assert np.array_equal(seg[c][short], ref[short], equal_nan=True), c
        if c == "stdev":
            # sqrt(var / w - (res / w)^2) is NaN or tiny when the true variance is 0 up to rounding (the reference
            # itself returns NaN there): compare the variances, NaN standing for 0
            a, b = np.nan_to_num(seg[c][~short]) ** 2, np.nan_to_num(ref[~short]) ** 2
            assert np.all(np.abs(a - b) <= 1e-10 * np.maximum(1.0, b)), c
        else:
            assert np.array_equal(np.isnan(seg[c]), np.isnan(ref)), c
            ok = ~short & ~np.isnan(ref)
            assert np.all(np.abs(seg[c][ok] - ref[ok]) <= 1e-12 * np.maximum(1.0, np.abs(ref[ok]))), c


@pytest.mark.parametrize("reps,kw", [((1,), {}), ((3,), {}), ((3, 3), {}), ((2,), dict(lambda2=1000.0)),
                                     ((1,), dict(max_iter=4, lambda2=10.0))])
def test_fast_against_oracle(engine, oracle, reps, kw):
    t = synth.chromosome_table(60_000, 41 + len(reps), reps)
    o = oracle.detect(t, 0, **kw)
    r = run_one(engine, t, **kw)
    assert_job_matches(r, o)
    so = oracle.segment_methylation_helper(o["estimates"], 0.01)
    sr = engine.segment_methylation_helper(r["estimates"], 0.01)
    for c in ("estimate_id", "segment_id", "start", "end"):
        assert np.array_equal(sr[c], so[c]), c


@pytest.mark.parametrize("model,reps,kw", [(1, (1,), {}), (1, (3,), {}), (2, (2, 2), dict(ref=1)),
                                           (2, (1, 2), dict(ref=2))])
def test_full_against_oracle(engine, oracle, model, reps, kw):
    t = synth.chromosome_table(6000, 51 + model, reps)
    full = dict(max_iter=30, nouter=20, hyper=100.0, bf_per_kb=50.0, max_lambda2=1000.0)
    o = oracle.detect(t, model, **full, **kw)
    r = run_one(engine, t, model=model, **full, **kw)
    # device exp/log/lgamma differ from glibc in the last ulp; a damping decision on a knife edge
    # could in principle flip, so the control-flow counters are compared first and reported
    print("control flow gpu/oracle:", (r["irls_steps"], r["dampings"], r["outer_iters"]),
          (o["irls_steps"], o["dampings"], o["outer_iters"]))
    assert_job_matches(r, o)


def test_batch_of_chromosomes_equals_single_calls(engine, oracle):
    tables = [synth.chromosome_table(n, 70 + i, (2,)) for i, n in enumerate((30_000, 500, 12_345, 2, 80_000))]
    ptr, cols = synth.concat_jobs(tables)
    res = engine.detect_batch(ptr, cols, params())
    try:
        assert res.status.tolist() == [0] * 5
        for j, t in enumerate(tables):
            assert_job_matches(res.job(j), oracle.detect(t, 0))
        # resident segmentation of the whole batch == per-job segment helper
        seg = res.segments(0.01)
        for j, t in enumerate(tables):
            so = oracle.segment_methylation_helper(oracle.detect(t, 0)["estimates"], 0.01)
            m = seg["job"] == j
            for c in ("estimate_id", "segment_id", "start", "end"):
                assert np.array_equal(seg[c][m], so[c]), (j, c)
            assert close(seg["betahat"][m], so["betahat"]) and close(seg["pseudoweight"][m], so["pseudoweight"])
        # PMD std fusion + re-segmentation on the resident segments == the same chain on the oracle
        pmd = res.pmd_segments(0.01, 1000.0)
        for j, t in enumerate(tables):
            so = oracle.segment_methylation_helper(oracle.detect(t, 0)["estimates"], 0.01)
            y = np.where(np.isfinite(so["stdev"]), so["stdev"], 0.0)
            w = so["end"].astype(float) - so["start"].astype(float) + 1.0
            fused = oracle.weighted_1d_flsa(y, w, 1000.0)
            po = oracle.segment_methylation_helper(dict(estimate_id=so["estimate_id"], start=so["start"].astype(np.int32),
                                                        beta=fused, betahat=y, pseudoweight=so["pseudoweight"]), 0.01)
            m = pmd["job"] == j
            for c in ("estimate_id", "segment_id", "start", "end"):
                assert np.array_equal(pmd[c][m], po[c]), (j, c)
            assert close(pmd["fused_std"][m], po["beta"], 1e-6)
    finally:
        res.free()


def test_validation_codes_per_job_and_batch_survives(engine):
    """One bad chromosome does not fail the batch (foreach .errorhandling='remove')."""
    good = synth.chromosome_table(5000, 5, (1,))
    tables = [good] + [t for t, _ in VALIDATION_CASES] + [good]
    want = [0] + [c for _, c in VALIDATION_CASES] + [0]
    ptr, cols = synth.concat_jobs(tables)
    res = engine.detect_batch(ptr, cols, params())
    try:
        got = res.status.tolist()
        # valid 3-row tables have 3 distinct positions -> fine
        assert got == want
        a, b = res.job(0), res.job(len(tables) - 1)
        assert np.array_equal(a["estimates"]["beta"], b["estimates"]["beta"])
        with pytest.raises(api.MethyLassoError) as e:
            res.job(1)
        assert e.value.code == 1
    finally:
        res.free()


def test_error_sites(engine):
    t = synth.chromosome_table(500, 1, (1,))
    for kw, code in ((dict(lambda2=0.0), 13), (dict(model=2, ref=1), 14)):
        with pytest.raises(api.MethyLassoError) as e:
            run_one(engine, t, **kw)
        assert e.value.code == code
    t2 = synth.chromosome_table(500, 1, (1, 1))
    with pytest.raises(api.MethyLassoError) as e:
        run_one(engine, t2, model=2, ref=3)
    assert e.value.code == 15
    same = _tbl(dataset=[1, 2], condition=[1, 1], replicate=[1, 2], pos=[7, 7], Nmeth=[1, 1], coverage=[2, 2])
    with pytest.raises(api.MethyLassoError) as e:
        run_one(engine, same)
    assert e.value.code == 18
    with pytest.raises(api.MethyLassoError):
        api.signal_detection_fast_singlechr({"pos": np.arange(3)}, engine=engine)


def test_segment_helper_generic_tables(engine, oracle):
    rng = np.random.default_rng(4)
    for trial in range(30):
        n = int(rng.integers(1, 4000))
        ids = np.sort(rng.integers(1, 4, n)).astype(np.int32)
        beta = np.cumsum(rng.normal(scale=0.004, size=n)) if trial % 2 else np.repeat(rng.random(n // 9 + 1), 9)[:n]
        est = dict(estimate_id=ids, start=np.cumsum(rng.integers(1, 300, n)).astype(np.int32), beta=beta,
                   betahat=beta + rng.normal(scale=0.1, size=n), pseudoweight=rng.integers(0, 60, n).astype(float))
        so = oracle.segment_methylation_helper(est, 0.01)
        sr = engine.segment_methylation_helper(est, 0.01)
        for c in so:
            assert np.array_equal(sr[c], so[c], equal_nan=True), (trial, c)   # rows added in row order: bit-equal


def test_segment_helper_long_segments(engine, oracle):
    """Segments of more than 256 rows are summed by a block (k_seg_emit_long): boundaries, ids, beta and integer
    pseudo-weights bit-equal, betahat / stdev equal to rounding."""
    rng = np.random.default_rng(14)
    sizes = [3, 257, 5000, 1, 100_000, 256, 40_000]
    levels = np.arange(len(sizes)) * 0.05
    beta = np.concatenate([np.full(n, v) for n, v in zip(sizes, levels)])
    n = beta.size
    ids = np.ones(n, np.int32)
    ids[sum(sizes[:5]):] = 2
    est = dict(estimate_id=ids, start=np.cumsum(rng.integers(1, 40, n)).astype(np.int32), beta=beta,
               betahat=beta + rng.normal(scale=0.1, size=n), pseudoweight=rng.integers(0, 60, n).astype(float))
    so = oracle.segment_methylation_helper(est, 0.01)
    sr = engine.segment_methylation_helper(est, 0.01)
    assert so["start"].size == len(sizes)
    for c in ("estimate_id", "segment_id", "start", "end", "beta", "pseudoweight"):
        assert np.array_equal(sr[c], so[c]), c
    for c in ("betahat", "stdev"):
        assert np.all(np.abs(sr[c] - so[c]) <= 1e-12 * np.maximum(1.0, np.abs(so[c]))), c
    short = np.array(sizes) <= 256
    for c in ("betahat", "stdev"):
        assert np.array_equal(sr[c][short], so[c][short]), c


def test_fast_diff_combine(engine, oracle):
    t = synth.chromosome_table(30_000, 88, (3, 3))
    o = oracle.detect(t, 0)
    d = oracle.fast_diff_combine(o["estimates"], o["n_params"], o["n_est"])
    r = api.difference_detection_fast_singlechr(t, ref=2, engine=engine)
    assert r["detection.type"] == "difference" and r["ref.cond"] == 2
    for c in ("beta", "betahat", "pseudoweight"):
        assert close(r["estimates"][c], d[c])


def test_genome_wide_wrapper_mirrors_r_layer(engine, oracle):
    tables = [synth.chromosome_table(8000, 90 + i, (2,)) for i in range(3)]
    names = ["chr2", "chr10", "chrX"]
    data = {k: np.concatenate([t[k] for t in tables]) for k in tables[0]}
    data["chr"] = np.concatenate([np.repeat(nm, t["pos"].size) for nm, t in zip(names, tables)])
    ret = api.signal_detection_fast(data, lambda2=25, verbose=False, engine=engine)
    assert ret["chr"].tolist() == names and ret["detection.type"] == "signal" and ret["model"] == "normal"
    at = 0
    for nm, t in zip(names, tables):
        o = oracle.detect(t, 0)
        n = o["estimates"]["beta"].size
        assert close(ret["estimates"]["beta"][at:at + n], o["estimates"]["beta"])
        assert (ret["estimates"]["chr"][at:at + n] == nm).all()
        at += n


def test_placement_independence(engine):
    """The same chromosome gives bit-identical results alone or inside any batch (what makes
    1-GPU and 8-GPU runs identical)."""
    a = synth.chromosome_table(50_000, 7, (3,))
    b = synth.chromosome_table(20_000, 8, (3,))
    r1 = run_one(engine, a)
    ptr, cols = synth.concat_jobs([b, a, b])
    res = engine.detect_batch(ptr, cols, params())
    try:
        r2 = res.job(1)
    finally:
        res.free()
    for c in ("beta", "betahat", "pseudoweight"):
        assert np.array_equal(r1["estimates"][c], r2["estimates"][c])
    assert np.array_equal(r1["mu"], r2["mu"]) and np.array_equal(r1["offset"], r2["offset"])


def test_chr1_three_replicates_against_the_port(engine, oracle):
    """The largest job of the bench workload (chr1: 2.26 M CpGs x 3 replicates) against the CPU port with the
    reference's verbatim DP: support, pseudodata and segment boundaries identical, estimates to 1e-10."""
    t = synth.chromosome_table(2_260_000, 20242 * 50, (3,))
    r = run_one(engine, t)
    o = oracle.detect(t, oracle.MODEL_FAST, lambda2=25.0)
    assert_job_matches(r, o)
    assert np.array_equal(r["estimates"]["pseudoweight"], o["estimates"]["pseudoweight"])
    assert np.array_equal(r["estimates"]["betahat"], o["estimates"]["betahat"])
    seg = engine.segment_methylation_helper(r["estimates"], 0.01)
    oseg = oracle.segment_methylation_helper(o["estimates"], 0.01)
    for c in ("estimate_id", "segment_id", "start", "end"):
        assert np.array_equal(seg[c], oseg[c]), c
    assert close(seg["beta"], oseg["beta"]) and close(seg["pseudoweight"], oseg["pseudoweight"])


def test_full_genome_shape_properties(engine):
    """Config 2 at 1/4 scale per chromosome count (24 jobs, 3 replicates): size-independent checks."""
    tables = synth.genome_tables(2, total_cpgs=7_000_000, reps_per_condition=(3,))
    ptr, cols = synth.concat_jobs(tables)
    res = engine.detect_batch(ptr, cols, params())
    try:
        assert res.status.tolist() == [0] * 24
        for j in (0, 11, 23):
            r = res.job(j)
            t = tables[j]
            est = r["estimates"]
            u, inv = np.unique(t["pos"], return_inverse=True)
            assert np.array_equal(est["start"], u.astype(float))
            assert np.array_equal(est["pseudoweight"], np.bincount(inv, t["coverage"], u.size))
            assert abs(np.sum(est["beta"] * est["pseudoweight"]) / np.sum(est["pseudoweight"])) < 1e-12
            assert np.allclose(r["mu"], est["beta"][inv] + r["offset"][t["dataset"] - 1], rtol=0, atol=1e-14)
            # FLSA optimality of beta + mean against (betahat, pseudoweight)
            from test_oracle import kkt_residual
            shift = np.sum(est["betahat"] * est["pseudoweight"]) / np.sum(est["pseudoweight"])
            assert kkt_residual(est["betahat"] - shift, est["pseudoweight"], est["beta"], 25.0) < 1e-8
    finally:
        res.free()


def test_cohort_batch_config4(engine, oracle):
    """BASELINE config 4 at reduced size: many independent samples x chromosomes in one batched call
    (on 8 GPUs the same job list is sharded by methylasso_b200.shard.lpt_assign)."""
    tables = [synth.chromosome_table(n, 1000 * s + c, (1,)) for s in range(8) for c, n in enumerate((9000, 4000, 1500))]
    ptr, cols = synth.concat_jobs(tables)
    res = engine.detect_batch(ptr, cols, params())
    try:
        assert res.status.tolist() == [0] * len(tables)
        for j in (0, 5, 13, 23):
            assert_job_matches(res.job(j), oracle.detect(tables[j], 0))
        seg = res.segments(0.01)
        so = oracle.segment_methylation_helper(oracle.detect(tables[13], 0)["estimates"], 0.01)
        m = seg["job"] == 13
        assert np.array_equal(seg["start"][m], so["start"]) and np.array_equal(seg["end"][m], so["end"])
    finally:
        res.free()


def test_engine_pool_grouping_is_result_neutral(engine):
    """EnginePool / split_jobs: any grouping of the chromosomes gives the same bits."""
    tables = [synth.chromosome_table(n, 400 + i, (2,)) for i, n in enumerate((20_000, 3000, 12_000, 700, 9000))]
    ptr, cols = synth.concat_jobs(tables)
    res = engine.detect_batch(ptr, cols, params())
    whole = [res.job(j) for j in range(5)]
    res.free()
    pool = api.EnginePool(2)
    try:
        def run(eng, grp):
            jobs, gptr, gcols = grp
            r = eng.detect_batch(gptr, gcols, params())
            out = [(j, r.job(i)) for i, j in enumerate(jobs)]
            r.free()
            return out
        got = dict(sum(pool.map(run, api.split_jobs(ptr, cols, 2)), []))
    finally:
        pool.close()
    for j in range(5):
        for c in ("beta", "betahat", "pseudoweight"):
            assert np.array_equal(got[j]["estimates"][c], whole[j]["estimates"][c])
        assert np.array_equal(got[j]["mu"], whole[j]["mu"])


def test_count_quotient_without_the_division_has_its_bits(engine):
    """counts_div.cuh on the device: all 67 M pairs (0 <= Nmeth <= 65536, 1 <= coverage < 1024), both signatures of
    div_counts against the device's IEEE division."""
    import ctypes as C
    pairs, bad = C.c_longlong(), C.c_longlong()
    engine._check(engine.lib.ml_debug_div_counts(engine.h, 65536, 1024, C.byref(pairs), C.byref(bad)))
    assert pairs.value == 65537 * 1023 and bad.value == 0
