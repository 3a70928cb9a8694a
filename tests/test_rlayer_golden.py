"""tests/golden/rlayer_fixtures.npz (made by tests/golden/make_golden_rlayer.py): the restatements of the R-layer steps
reproduce their committed outputs (CPU), and the CUDA path reproduces them too (GPU), without running the oracle."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rlayer_fixtures.npz")
COLS = ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage")
INT_COLS = ("condition", "estimate_id", "segment_id", "start", "end", "num_cpgs", "num_cpgs_ref")
FLOAT_COLS = ("width", "coverage", "coverage_ref", "pseudoweight", "beta", "std", "beta_ref", "std_ref", "diff", "cohen_d",
              "pval", "coverage_score", "fdr")


def close(a, b, tol):
    a, b = np.asarray(a, float), np.asarray(b, float)
    ok = ~np.isnan(b)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.all(np.abs(a[ok] - b[ok]) <= tol * np.maximum(1.0, np.abs(b[ok])))


def test_oracle_reproduces_fixtures(oracle):
    from oracle import codec as OC, diffcall as DC, lmr_umr as LU
    from test_calltable import pooled_from_table
    g = np.load(GOLD)
    t = {k: g["dmr_in_" + k] for k in COLS}
    o = oracle.detect(t, 0, lambda2=25.0)
    tab, called = DC.call_differences([DC.call_differences_singlechr(t, oracle.fast_diff_combine(o["estimates"], o["n_params"], o["n_est"]), 2)])
    for k in INT_COLS:
        assert np.array_equal(tab[k], g["dmr_out_" + k]), k
    for k in FLOAT_COLS:
        assert close(tab[k], g["dmr_out_" + k], 1e-13), k
    assert np.array_equal(called, g["dmr_out_called"]) and called.sum() >= 1
    t = {k: g["lu_in_" + k] for k in COLS}
    o = oracle.detect(t, 0, lambda2=25.0)
    call = oracle.call_table(pooled_from_table(t), oracle.segment_methylation_helper(o["estimates"], 0.01), 500.0)
    for k, v in LU.annotate(call).items():
        assert np.array_equal(v, g["lu_out_" + k], equal_nan=True), k
    text = g["codec_text"].tobytes()
    for mode, kw in (("counts", dict(mC=5, uC=6)), ("level", dict(cov=5, meth=4))):
        r = OC.read_replicate(text, min_coverage=5, **kw)
        assert np.array_equal(r["chr"].astype("U"), g[f"codec_{mode}_chr"])
        for k in ("pos", "coverage", "Nmeth", "beta"):
            assert np.array_equal(r[k], g[f"codec_{mode}_{k}"]), k


@pytest.mark.gpu
def test_gpu_reproduces_fixtures(engine):
    from methylasso_b200 import codec
    from methylasso_b200._native import MlParams
    g = np.load(GOLD)
    p = MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
    # DMR caller + FDR + threshold mask
    t = {k: g["dmr_in_" + k] for k in COLS}
    batch = engine.upload(np.array([0, t["pos"].size], np.int64), t)
    res = engine.detect(batch, p)
    try:
        res.fast_diff_combine()
        tab = dict(res.call_differences(batch, 2))
        tab["fdr"] = engine.p_adjust_fdr(tab["pval"])
        for k in INT_COLS:
            assert np.array_equal(np.asarray(tab[k], np.int64), g["dmr_out_" + k]), k
        for k in FLOAT_COLS:
            assert close(tab[k], g["dmr_out_" + k], 1e-9 if k == "cohen_d" else 1e-12), k
        with np.errstate(invalid="ignore"):
            called = (~np.isnan(tab["diff"]) & (np.abs(tab["diff"]) >= 0.1) & (tab["coverage_score"] >= 70)
                      & (tab["pval"] <= 0.05) & (tab["fdr"] <= 1) & (np.minimum(tab["num_cpgs"], tab["num_cpgs_ref"]) >= 4))
        assert np.array_equal(called, g["dmr_out_called"])
    finally:
        res.free()
        batch.free()
    # LMR / UMR annotation
    t = {k: g["lu_in_" + k] for k in COLS}
    res = engine.detect_batch(np.array([0, t["pos"].size], np.int64), t, p)
    try:
        for k, v in res.call_lmr_umr(0.01, 500.0).items():
            if k == "flank_beta":      # a neighbour's mean methylation: the call table's double-double mean against the
                assert close(v, g["lu_out_" + k], 1e-13), k       # oracle's long double one, equal to the last bit or so
            else:
                assert np.array_equal(v, g["lu_out_" + k], equal_nan=True), k
    finally:
        res.free()
    # codec
    text = g["codec_text"].tobytes()
    for mode, kw in (("counts", dict(mC=5, uC=6)), ("level", dict(cov=5, meth=4))):
        r = codec.read_replicate(engine, text, min_coverage=5, **kw)
        assert np.array_equal(r["chr"].astype("U"), g[f"codec_{mode}_chr"])
        for k in ("pos", "coverage", "Nmeth", "beta"):
            assert np.array_equal(r[k], g[f"codec_{mode}_{k}"]), k


VGOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "valley_fixtures.npz")
VALLEY_CASES = (("default", {}), ("narrow", dict(pmd_valley_min_width=1200.0)))


def test_oracle_reproduces_valley_fixtures(oracle):
    """tests/golden/valley_fixtures.npz (made by tests/golden/make_golden_valleys.py): R/call.R:100-133 restated."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden_valleys as MG
    v = np.load(VGOLD)
    got = MG.compute(np.load(GOLD))
    assert set(got) == set(v.files)
    for k in v.files:
        assert np.array_equal(got[k], v[k], equal_nan=True), k
    assert v["narrow_start"].size >= 10


@pytest.mark.gpu
def test_gpu_reproduces_valley_fixtures(engine):
    from methylasso_b200._native import MlParams
    g, v = np.load(GOLD), np.load(VGOLD)
    t = {k: g["lu_in_" + k] for k in COLS}
    res = engine.detect_batch(np.array([0, t["pos"].size], np.int64), t, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        for name, kw in VALLEY_CASES:
            got = res.call_valleys(0.01, 500.0, **kw)
            for k in ("condition", "start", "end", "width", "coverage", "num_cpgs"):
                assert np.array_equal(np.asarray(got[k], np.float64), np.asarray(v[f"{name}_{k}"], np.float64)), (name, k)
            # means of the call table's per-segment means: the call table's double-double mean against the oracle's long
            # double one can differ in the last bit, and these inherit it
            for k in ("beta", "pseudoweight"):
                assert close(got[k], v[f"{name}_{k}"], 1e-13), (name, k)
            assert np.array_equal(got["row_in_valley"] != 0, v[f"{name}_row_in_valley"]), name
            assert np.array_equal(got["low_id_width"], v[f"{name}_low_id_width"], equal_nan=True), name
            assert close(got["low_id_beta"], v[f"{name}_low_id_beta"], 1e-13), name
    finally:
        res.free()


RE_CASES = (("default", {}), ("narrow", dict(pmd_lambda2=100.0, pmd_valley_min_width=1200.0, pmd_std_threshold=0.1)))


@pytest.mark.gpu
def test_gpu_reproduces_rule_engine_fixtures(engine):
    """The tables of the whole rule engine (R/call.R:31-221, drop = F) in valley_fixtures.npz against the resident path."""
    from methylasso_b200._native import MlParams
    g, v = np.load(GOLD), np.load(VGOLD)
    t = {k: g["lu_in_" + k] for k in COLS}
    res = engine.detect_batch(np.array([0, t["pos"].size], np.int64), t, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        call = res.call_table(0.01, 500.0)
        for name, kw in RE_CASES:
            vkw = {k: x for k, x in kw.items() if k == "pmd_valley_min_width"}
            val = res.call_valleys(0.01, 500.0, **vkw)
            merged, _ = res.call_pmd_split(0.01, 500.0, kw.get("pmd_lambda2", 1000.0), True, **vkw)
            pmd = res.call_pmds(0.01, 500.0, **kw)
            fin = res.call_pmd_status(0.01, 500.0, **kw)
            w = {k[len(f"re_{name}_lmr_umr_valley_"):]: v[k] for k in v.files if k.startswith(f"re_{name}_lmr_umr_valley_")}
            assert fin["start"].size == w["start"].size
            for k, gk in (("condition", "condition"), ("start", "start"), ("end", "end"), ("category", "category"), ("pmd_status", "pmd_status")):
                assert np.array_equal(np.asarray(fin[gk], np.float64), np.asarray(w[k], np.float64)), (name, k)
            assert np.array_equal(merged["segment_id"][fin["merged_row"]], w["segment_id"]), name
            is_row = fin["src_row"] >= 0
            for k in ("width", "beta", "coverage", "pseudoweight", "num_cpgs"):
                got = np.where(is_row, call[k][np.maximum(fin["src_row"], 0)],
                               val[k][np.maximum(fin["src_valley"], 0)] if val[k].size else np.nan)
                assert close(got, w[k], 1e-12), (name, k)
            wp = {k[len(f"re_{name}_pmd_"):]: v[k] for k in v.files if k.startswith(f"re_{name}_pmd_")}
            assert pmd["start"].size == wp["start"].size
            for k in ("condition", "category", "start", "end", "width", "num_cpgs"):
                assert np.array_equal(np.asarray(pmd[k], np.float64), np.asarray(wp[k], np.float64)), (name, k)
            for k in ("fused_std", "beta", "std"):
                assert close(pmd[k], wp[k], 1e-12), (name, k)
            assert (wp["category"] == 1).sum() >= 5 and (w["category"] > 0).sum() >= 50
    finally:
        res.free()
