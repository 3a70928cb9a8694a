"""LMR / UMR annotation of R/call.R:75-91: the numpy restatement (oracle/lmr_umr.py) on hand-made tables (CPU), and the
CUDA kernels against it on resident call tables (GPU)."""
import numpy as np
import pytest

from methylasso_b200 import synth

T, F, NA = 1, 0, -1


def table(beta, width=None, ncpg=None, std=None, cond=None, gap=100):
    n = len(beta)
    width = np.full(n, 400.0) if width is None else np.asarray(width, float)
    start = np.cumsum(np.concatenate([[1000], width[:-1] + gap])).astype(np.uint32)
    return dict(estimate_id=np.ones(n, np.int32) if cond is None else np.asarray(cond, np.int32),
                start=start, end=(start + width - 1).astype(np.uint32), width=width, beta=np.asarray(beta, float),
                num_cpgs=np.full(n, 12, np.int32) if ncpg is None else np.asarray(ncpg, np.int32),
                std=np.full(n, 0.02) if std is None else np.asarray(std, float))


def test_oracle_hand_example(oracle):
    from oracle import lmr_umr as LU
    # high - low (LMR-like) - high - very low (UMR) - high; all admissible (width 400 > 300: all selected by :79)
    a = LU.annotate(table([0.9, 0.3, 0.9, 0.03, 0.9]))
    assert a["is_admissible"].tolist() == [1] * 5
    assert a["is_V"].tolist() == [F, T, F, T, F]                # at the ends NA & FALSE is FALSE
    assert a["is_flank"].tolist() == [T, F, T, F, T]            # != T and a V next to it (NA | TRUE is TRUE)
    assert a["is_sep"].tolist() == [NA, T, T, T, NA]            # the first / last selected row has an NA neighbour
    assert np.isnan(a["flank_dist"][[0, 4]]).all() and a["flank_dist"][1:4].tolist() == [101.0] * 3
    assert a["category"].tolist() == [0, 1, 0, 2, 0]
    a = LU.annotate(table([0.9, 0.2, 0.9, 0.3, 0.9, 0.03, 0.9]))
    assert a["is_V"].tolist() == [F, T, F, T, F, T, F] and a["category"].tolist() == [0, 1, 0, 1, 0, 2, 0]
    # a small LMR whose lowest neighbour is itself low is dropped (:84)
    a = LU.annotate(table([0.9, 0.6, 0.45, 0.3, 0.45, 0.9, 0.9], ncpg=[12, 12, 12, 5, 12, 12, 12]))
    assert a["is_V"].tolist() == [F, F, F, T, F, F, F]
    assert a["is_flank"].tolist() == [NA, F, T, F, T, F, NA]    # TRUE & (NA | FALSE) is NA at the ends
    assert a["flank_beta"][3] == 0.45 and a["is_sep"][3] == T and a["category"][3] == 0
    a = LU.annotate(table([0.9, 0.6, 0.45, 0.3, 0.45, 0.9, 0.9]))
    assert a["category"][3] == 1                                # the same with 12 CpGs stays an LMR
    # a noisy one is not separated from its neighbours (:82)
    a = LU.annotate(table([0.9, 0.8, 0.9, 0.3, 0.9, 0.8, 0.9], std=[0.02, 0.02, 0.02, 0.4, 0.02, 0.02, 0.02]))
    assert a["is_V"][3] == T and a["is_sep"][3] == F and a["category"][3] == 0
    # inadmissible rows (one CpG: width / 0 = Inf) are invisible to :78 - the V is formed across them
    a = LU.annotate(table([0.9, 0.5, 0.3, 0.5, 0.9], ncpg=[12, 1, 12, 1, 12]))
    assert a["is_admissible"].tolist() == [1, 0, 1, 0, 1] and a["is_V"].tolist() == [F, NA, T, NA, F]
    # conditions do not see each other in :78
    a = LU.annotate(table([0.9, 0.3, 0.9, 0.9, 0.3, 0.9], cond=[1, 1, 1, 2, 2, 2]))
    assert a["is_V"].tolist() == [F, T, F, F, T, F]
    # a condition's first row with a higher successor: NA & TRUE is NA
    a = LU.annotate(table([0.9, 0.3, 0.9, 0.2, 0.9, 0.9], cond=[1, 1, 1, 2, 2, 2]))
    assert a["is_V"].tolist() == [F, T, F, NA, F, F]
    # :79 has no `by`: the last row of condition 1 sees the first row of condition 2 as its successor, so its
    # is.flank is TRUE & (FALSE | FALSE) = FALSE, where a per-condition shift would give TRUE & (FALSE | NA) = NA
    a = LU.annotate(table([0.9] * 6, cond=[1, 1, 1, 2, 2, 2]))
    assert a["is_V"].tolist() == [F] * 6 and a["is_flank"].tolist() == [NA, F, F, F, F, NA]


def oracle_call_tables(O, tables):
    from test_calltable import pooled_from_table
    out = []
    for t in tables:
        o = O.detect(t, 0, lambda2=25.0)
        seg = O.segment_methylation_helper(o["estimates"], 0.01)
        out.append(O.call_table(pooled_from_table(t), seg, 500.0))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("reps", [(2,), (2, 2)])
def test_gpu_lmr_umr_matches_oracle(engine, oracle, reps):
    from methylasso_b200._native import MlParams
    from oracle import lmr_umr as LU
    tables = [synth.chromosome_table(n, 500 + i, reps) for i, n in enumerate((60_000, 9_000, 25_000))]
    ptr, cols = synth.concat_jobs(tables)
    res = engine.detect_batch(ptr, cols, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        call = res.call_table(0.01, 500.0)
        for kw in ({}, dict(min_num_cpgs=2, flank_width=100.0, lmr_max_beta=0.6, umr_large_density=0.01)):
            got = res.call_lmr_umr(0.01, 500.0, **kw)
            assert got["category"].size == call["start"].size
            n_cat = 0
            for j in range(len(tables)):
                m = call["job"] == j
                want = LU.annotate({k: v[m] for k, v in call.items()}, **kw)
                for k, w in want.items():
                    assert np.array_equal(got[k][m], w, equal_nan=True), (j, k)
                n_cat += int((want["category"] > 0).sum())
            assert n_cat > 20
        # the call table of the resident path is the oracle's, so the annotation is the oracle's end to end
        want_tables = oracle_call_tables(oracle, tables)
        got = res.call_lmr_umr(0.01, 500.0)
        for j, ct in enumerate(want_tables):
            m = call["job"] == j
            assert np.array_equal(got["category"][m], LU.annotate(ct)["category"])
    finally:
        res.free()
