"""The per-segment call table of R/call.R:43-68: oracle restatement (CPU) and the CUDA builder (GPU)."""
import numpy as np
import pytest

from methylasso_b200 import synth


def pooled_from_table(t):
    """R/call.R:43-46: coverage and Nmeth summed per (condition, pos); estimate_id = condition."""
    est, pos, nm, cv = [], [], [], []
    for c in np.unique(t["condition"]):
        m = t["condition"] == c
        u, inv = np.unique(t["pos"][m], return_inverse=True)
        est.append(np.full(u.size, c, np.int32))
        pos.append(u.astype(np.int32))
        nm.append(np.bincount(inv, t["Nmeth"][m], u.size).astype(np.int32))
        cv.append(np.bincount(inv, t["coverage"][m], u.size).astype(np.int32))
    return dict(estimate_id=np.concatenate(est), pos=np.concatenate(pos), Nmeth=np.concatenate(nm),
                coverage=np.concatenate(cv))


def brute_force(pooled, seg, max_distance):
    """Row-by-row python statement of the same R lines (membership start-1 <= pos <= end, gap split, shrink,
    numpy mean / sample sd), for small cases."""
    rows = []
    for e in np.unique(seg["estimate_id"]):
        shift = 0
        pm = pooled["estimate_id"] == e
        P, X = pooled["pos"][pm].astype(np.int64), pooled["Nmeth"][pm] / pooled["coverage"][pm]
        Cv = pooled["coverage"][pm]
        for s in np.nonzero(seg["estimate_id"] == e)[0]:
            a, b = int(seg["start"][s]), int(seg["end"][s])
            if a > b:
                continue
            idx = np.nonzero((P >= a - 1) & (P <= b))[0]
            if idx.size == 0:
                continue
            cuts = np.nonzero(np.diff(P[idx]) > max_distance)[0] + 1
            for k, part in enumerate(np.split(idx, cuts)):
                if k > 0:
                    shift += 1
                x = X[part]
                rows.append((e, seg["segment_id"][s] + shift, P[part].min(), P[part].max() + 2,
                             P[part].max() + 2 - P[part].min() + 1, x.mean(), Cv[part].sum(), seg["pseudoweight"][s],
                             part.size, x.std(ddof=1) if part.size > 1 else 0.0))
    return np.array(rows, dtype=float)


def as_matrix(t):
    keys = ("estimate_id", "segment_id", "start", "end", "width", "beta", "coverage", "pseudoweight", "num_cpgs", "std")
    return np.stack([np.asarray(t[k], float) for k in keys], axis=1)


def random_case(rng, n_est=2, n=400):
    est, pos, nm, cv = [], [], [], []
    sg = dict(estimate_id=[], segment_id=[], start=[], end=[], pseudoweight=[])
    for e in range(1, n_est + 1):
        p = np.cumsum(rng.choice([1, 1, 2, 7, 40, 90, 501, 1500], n)).astype(np.int32) + 1000
        c = rng.integers(1, 60, n).astype(np.int32)
        m = rng.binomial(c, rng.random(n)).astype(np.int32)
        est.append(np.full(n, e, np.int32)); pos.append(p); nm.append(m); cv.append(c)
        heads = np.sort(rng.choice(np.arange(1, n), int(min(n - 1, rng.integers(1, 30))), replace=False))
        heads = np.concatenate([[0], heads])
        for k, h in enumerate(heads):
            sg["estimate_id"].append(e)
            sg["segment_id"].append(k + 1)
            sg["start"].append(p[h])
            sg["end"].append(p[heads[k + 1]] - 1 if k + 1 < heads.size else p[-1])
            sg["pseudoweight"].append(float(c[h:(heads[k + 1] if k + 1 < heads.size else n)].sum()))
    pooled = dict(estimate_id=np.concatenate(est), pos=np.concatenate(pos), Nmeth=np.concatenate(nm),
                  coverage=np.concatenate(cv))
    seg = dict(estimate_id=np.array(sg["estimate_id"], np.int32), segment_id=np.array(sg["segment_id"], np.int32),
               start=np.array(sg["start"], np.uint32), end=np.array(sg["end"], np.uint32),
               pseudoweight=np.array(sg["pseudoweight"]))
    return pooled, seg


# ---- CPU: the oracle -------------------------------------------------------------------------------
def test_oracle_hand_example(oracle):
    pos = np.array([10, 20, 30, 700, 710, 800, 805], np.int32)
    pooled = dict(estimate_id=np.ones(7, np.int32), pos=pos, Nmeth=np.array([1, 2, 3, 4, 5, 0, 10], np.int32),
                  coverage=np.full(7, 10, np.int32))
    seg = dict(estimate_id=np.array([1, 1], np.int32), segment_id=np.array([1, 2], np.int32),
               start=np.array([10, 800], np.uint32), end=np.array([799, 805], np.uint32),
               pseudoweight=np.array([50., 20.]))
    t = oracle.call_table(pooled, seg, 500)
    assert t["segment_id"].tolist() == [1, 2, 3]          # the 670 bp gap splits segment 1; segment 2 shifts to 3
    assert t["start"].tolist() == [10, 700, 800] and t["end"].tolist() == [32, 712, 807]
    assert t["width"].tolist() == [23., 13., 8.]
    assert np.allclose(t["beta"], [0.2, 0.45, 0.5], atol=1e-16)
    assert t["coverage"].tolist() == [30., 20., 20.] and t["pseudoweight"].tolist() == [50., 50., 20.]
    assert t["num_cpgs"].tolist() == [3, 2, 2]
    assert np.allclose(t["std"], [0.1, np.sqrt(0.005), np.sqrt(0.5)], rtol=1e-15)


def test_oracle_single_cpg_and_adjacent_positions(oracle):
    # a CpG at start-1 of the next segment belongs to both segments (foverlaps of [pos, pos+1] with [start, end])
    pooled = dict(estimate_id=np.ones(4, np.int32), pos=np.array([5, 9, 10, 50], np.int32),
                  Nmeth=np.array([1, 2, 3, 4], np.int32), coverage=np.full(4, 4, np.int32))
    seg = dict(estimate_id=np.array([1, 1, 1], np.int32), segment_id=np.array([1, 2, 3], np.int32),
               start=np.array([5, 10, 50], np.uint32), end=np.array([9, 49, 50], np.uint32), pseudoweight=np.ones(3))
    t = oracle.call_table(pooled, seg, 500)
    assert t["num_cpgs"].tolist() == [2, 2, 1]            # pos 9 counted in segments 1 and 2
    assert t["start"].tolist() == [5, 9, 50]
    assert t["std"][2] == 0.0                             # sd of one value is NA -> 0 (R/call.R:72)


def test_oracle_matches_brute_force(oracle):
    rng = np.random.default_rng(11)
    for trial in range(25):
        pooled, seg = random_case(rng, n_est=int(rng.integers(1, 4)), n=int(rng.integers(5, 300)))
        got = as_matrix(oracle.call_table(pooled, seg, 500))
        want = brute_force(pooled, seg, 500)
        assert got.shape == want.shape
        assert np.array_equal(got[:, :5], want[:, :5]) and np.array_equal(got[:, 6:9], want[:, 6:9])
        assert np.allclose(got[:, 5], want[:, 5], rtol=1e-13, atol=1e-15)
        assert np.allclose(got[:, 9], want[:, 9], rtol=1e-9, atol=1e-12)


# ---- GPU ---------------------------------------------------------------------------------------------
def assert_call_tables_match(got, want):
    for k in ("estimate_id", "segment_id", "start", "end", "num_cpgs"):
        assert np.array_equal(np.asarray(got[k], np.int64), np.asarray(want[k], np.int64)), k
    for k in ("width", "coverage", "pseudoweight"):
        assert np.array_equal(got[k], want[k]), k
    # R accumulates in 80-bit long double, the device in double-double: both round the same real number
    assert np.max(np.abs(got["beta"] - want["beta"]), initial=0.0) < 1e-15
    assert np.allclose(got["std"], want["std"], rtol=1e-12, atol=1e-15)


@pytest.mark.gpu
def test_gpu_call_table_host_tables(engine, oracle):
    rng = np.random.default_rng(12)
    for trial in range(12):
        pooled, seg = random_case(rng, n_est=int(rng.integers(1, 4)), n=int(rng.integers(5, 3000)))
        assert_call_tables_match(engine.segment_call_table(pooled, seg, 500.0), oracle.call_table(pooled, seg, 500.0))


@pytest.mark.gpu
@pytest.mark.parametrize("reps", [(1,), (3,), (2, 2)])
def test_gpu_resident_call_table_and_pmd(engine, oracle, reps):
    """Whole R-layer chain on resident data: detect -> segments -> call table -> std fusion -> PMD segments."""
    from methylasso_b200._native import MlParams
    tables = [synth.chromosome_table(n, 300 + i + len(reps), reps) for i, n in enumerate((40_000, 9_000))]
    ptr, cols = synth.concat_jobs(tables)
    res = engine.detect_batch(ptr, cols, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        call = res.call_table(0.01, 500.0)
        pmd = res.call_pmd_segments(0.01, 500.0, 1000.0, want_fused=True)
        for j, t in enumerate(tables):
            o = oracle.detect(t, 0, lambda2=25.0)
            seg = oracle.segment_methylation_helper(o["estimates"], 0.01)
            want = oracle.call_table(pooled_from_table(t), seg, 500.0)
            mine = {k: v[call["job"] == j] for k, v in call.items()}
            assert_call_tables_match(mine, want)
            # PMD fusion on the oracle's call table, condition by condition (R/call.R:93-95)
            fused = np.concatenate([oracle.weighted_1d_flsa(want["std"][want["estimate_id"] == e],
                                                            want["width"][want["estimate_id"] == e], 1000.0)
                                    for e in np.unique(want["estimate_id"])])
            got_fused = pmd["row_fused_std"][call["job"] == j]
            assert np.max(np.abs(got_fused - fused)) < 1e-9
            want_pmd = oracle.segment_methylation_helper(dict(estimate_id=want["estimate_id"], start=want["start"].astype(np.int32),
                                                              beta=fused, betahat=want["std"],
                                                              pseudoweight=want["pseudoweight"]), 0.01)
            mine_pmd = {k: v[pmd["job"] == j] for k, v in pmd.items() if k != "row_fused_std"}
            for k in ("estimate_id", "segment_id", "start", "end"):
                assert np.array_equal(np.asarray(mine_pmd[k], np.int64), np.asarray(want_pmd[k], np.int64)), k
    finally:
        res.free()


@pytest.mark.gpu
def test_gpu_resident_call_table_needs_one_step_fast(engine):
    from methylasso_b200._native import MlParams
    from methylasso_b200.api import MethyLassoError
    t = synth.chromosome_table(3000, 5, (1,))
    res = engine.detect_batch(np.array([0, t["pos"].size], np.int64), t, MlParams(10.0, 11.0, 0.0, 0.01, 50.0, 1.0, 3, 1, 1, 0, 0, 0))
    try:
        with pytest.raises(MethyLassoError):
            res.call_table()
    finally:
        res.free()
