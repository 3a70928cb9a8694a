"""Parity of the code path bench.py's end-to-end numbers are measured on: `pipeline.GenomePipeline` (several engines of
one GPU, page-locked host buffers, run-table upload, want_mu = 0, asynchronous downloads, engine-owned pinned result
buffers).  Its outputs are compared with single-engine calls column by column (array_equal) and with the CPU
restatement of the reference (oracle.detect, the segment helper and the rule-engine chain)."""
import numpy as np
import pytest

from methylasso_b200 import synth
from methylasso_b200._native import MlParams

SIZES = (60_000, 35_000, 35_000, 20_000, 12_000, 9_000, 5_000)


def _genome(reps=(3,)):
    # dmv=True: 6-15 kb unmethylated stretches and LMRs inside PMD blocks, so that the valley (UMR-DMV) and the
    # LMR-in-PMD rules of R/call.R:100-133, 201-212 are exercised
    tables = [synth.chromosome_table(n, 4100 + i, reps, dmv=True) for i, n in enumerate(SIZES)]
    ptr, cols = synth.concat_jobs(tables)
    return tables, ptr, cols


def _close(a, b, rel=1e-10):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= rel * np.maximum(np.abs(b), 1e-3)))


@pytest.mark.gpu
def test_pipeline_region_tables(engine, oracle):
    """GenomePipeline.segment_methylation (the timed e2e call of bench.py): equal to one single-engine call on the whole
    batch, bit for bit, and to the chain of CPU restatements per chromosome (integers exactly, floats to 1e-10)."""
    from methylasso_b200.pipeline import GenomePipeline
    from oracle import rule_engine as RE
    from test_calltable import pooled_from_table
    tables, ptr, cols = _genome()
    pipe = GenomePipeline(ptr, cols, n_engines=3)
    try:
        pipe.segment_methylation()                        # a first call sizes the engines' pinned buffers
        seg, pmd, status = pipe.segment_methylation()
        seg = {k: v.copy() for k, v in seg.items()}
        pmd = {k: v.copy() for k, v in pmd.items()}
        assert pipe.h2d_bytes() == 6 * int(ptr[-1])     # pos + two one-byte counts per row
    finally:
        pipe.close()
    assert not status.any()
    # one engine, one batch, mu stored: the same tables
    engine.set("want_mu", 1)
    res = engine.detect_batch(ptr, cols, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        seg1, pmd1 = res.segment_methylation_tables(0.01, 500.0)
    finally:
        res.free()
    for k in seg1:
        assert np.array_equal(seg[k], seg1[k], equal_nan=seg1[k].dtype.kind == "f"), k
    for k in pmd1:
        assert np.array_equal(pmd[k], pmd1[k], equal_nan=pmd1[k].dtype.kind == "f"), k
    n_rows = 0
    for j, t in enumerate(tables):
        o = oracle.detect(t, oracle.MODEL_FAST, lambda2=25.0)
        want = RE.segment_methylation_singlechr(pooled_from_table(t), o["estimates"])
        m, w = seg["job"] == j, want["lmr_umr_valley"]
        assert int(m.sum()) == w["start"].size, j
        for k in ("condition", "start", "end", "segment_id", "num_cpgs", "category", "pmd_status", "width", "coverage"):
            assert np.array_equal(np.asarray(seg[k][m], np.float64), np.asarray(w[k], np.float64)), (j, k)
        for k in ("beta", "pseudoweight", "std"):
            a, b = seg[k][m], np.asarray(w[k], float)
            assert np.array_equal(np.isnan(a), np.isnan(b)) and _close(a[~np.isnan(a)], b[~np.isnan(b)]), (j, k)
        pm, wp = pmd["job"] == j, want["pmd"]
        assert int(pm.sum()) == wp["start"].size, j
        for k in ("condition", "start", "end", "width", "num_cpgs"):
            assert np.array_equal(np.asarray(pmd[k][pm], np.float64), np.asarray(wp[k], np.float64)), (j, k)
        for k in ("beta", "std", "fused_std"):
            assert _close(pmd[k][pm], wp[k]), (j, k)
        n_rows += int(m.sum())
    assert n_rows > 100 and pmd["start"].size > 3
    assert int((seg["category"] == 3).sum()) >= 2          # valleys were called (and equal the restatement's)


@pytest.mark.gpu
def test_pipeline_per_cpg_outputs(engine, oracle):
    """GenomePipeline.signal_detection_fast (copy_all_async into page-locked buffers, reuse_pinned table fetches): every
    output column equals Result.job() of a single-engine call bit for bit; ids, starts and segment boundaries equal the
    CPU restatement's, estimates agree to 1e-10."""
    from methylasso_b200.pipeline import GenomePipeline
    tables, ptr, cols = _genome()
    pipe = GenomePipeline(ptr, cols, n_engines=3)
    try:
        pipe.signal_detection_fast()
        outs = pipe.signal_detection_fast()
        engine.set("want_mu", 1)
        res = engine.detect_batch(ptr, cols, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
        try:
            seen = 0
            for o in outs:
                r_at = e_at = d_at = 0
                for gj, j in enumerate(o["jobs"]):
                    ref = res.job(int(j))
                    n, P = tables[j]["pos"].size, ref["estimates"]["start"].size
                    assert np.array_equal(o["mu"][r_at:r_at + n], ref["mu"]), j
                    assert np.array_equal(o["offset"][d_at:d_at + 3], ref["offset"]), j
                    for k in ("estimate_id", "start", "beta", "betahat", "pseudoweight"):
                        assert np.array_equal(o["estimates"][k][e_at:e_at + P], ref["estimates"][k]), (j, k)
                    want = oracle.detect(tables[j], oracle.MODEL_FAST, lambda2=25.0)
                    assert np.array_equal(o["estimates"]["start"][e_at:e_at + P], want["estimates"]["start"]), j
                    assert np.array_equal(o["estimates"]["estimate_id"][e_at:e_at + P], want["estimates"]["estimate_id"]), j
                    for k in ("beta", "betahat", "pseudoweight"):
                        assert _close(o["estimates"][k][e_at:e_at + P], want["estimates"][k]), (j, k)
                    assert _close(o["mu"][r_at:r_at + n], want["mu"]), j
                    wseg = oracle.segment_methylation_helper(want["estimates"], 0.01)
                    sm = o["segments"]["job"] == gj
                    assert np.array_equal(o["segments"]["start"][sm], wseg["start"]), j
                    assert np.array_equal(o["segments"]["end"][sm], wseg["end"]), j
                    r_at += n
                    e_at += P
                    d_at += 3
                    seen += 1
                assert not o["status"].any()
            assert seen == len(tables)
        finally:
            res.free()
    finally:
        pipe.close()


@pytest.mark.gpu
def test_want_mu_off_refuses_mu(engine):
    """want_mu = 0: a one-step FAST fit does not store mu, asking for it is an error, everything else is unchanged."""
    from methylasso_b200 import api
    t = synth.chromosome_table(8_000, 77, (2,))
    ptr = np.array([0, t["pos"].size], np.int64)
    p = MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
    engine.set("want_mu", 1)
    a = engine.detect_batch(ptr, t, p)
    ja = a.job(0)
    a.free()
    engine.set("want_mu", 0)
    try:
        b = engine.detect_batch(ptr, t, p)
        try:
            jb = b.job(0, want_mu=False)
            for k in ("beta", "betahat", "pseudoweight", "start"):
                assert np.array_equal(ja["estimates"][k], jb["estimates"][k]), k
            assert np.array_equal(ja["offset"], jb["offset"])
            assert b.job(0, want_mu=True)["mu"] is None
            with pytest.raises(api.MethyLassoError):
                b.copy_all(mu=np.empty(t["pos"].size))
        finally:
            b.free()
    finally:
        engine.set("want_mu", 1)


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [1, 30, 4000])
def test_packed_upload_modes(engine, oracle, scale):
    """ml_batch_stage packs Nmeth / coverage with one byte per count (all < 256), two (all < 65536) or not at all: the
    fit is the same as the CPU restatement's in every mode, through upload and through stage + commit."""
    t = synth.chromosome_table(30_000, 606, (2,))
    t = dict(t, Nmeth=(t["Nmeth"] * scale).astype(np.int32), coverage=(t["coverage"] * scale).astype(np.int32))
    assert (scale == 1) == (int(t["coverage"].max()) < 256) and (scale <= 30) == (int(t["coverage"].max()) < 65536)
    ptr = np.array([0, t["pos"].size], np.int64)
    p = MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
    want = oracle.detect(t, oracle.MODEL_FAST, lambda2=25.0)
    engine.set("want_mu", 1)
    for two_calls in (False, True):
        bt = engine.commit(engine.stage(ptr, t)) if two_calls else engine.upload(ptr, t)
        res = engine.detect(bt, p)
        try:
            got = res.job(0)
        finally:
            res.free()
            bt.free()
        assert np.array_equal(got["estimates"]["pseudoweight"], want["estimates"]["pseudoweight"])
        assert np.array_equal(got["estimates"]["betahat"], want["estimates"]["betahat"])
        assert _close(got["estimates"]["beta"], want["estimates"]["beta"]) and _close(got["mu"], want["mu"])


@pytest.mark.gpu
def test_packed_upload_keeps_the_validation_codes(engine):
    """A negative count or Nmeth > coverage must still be refused with the reference's own message code
    (core/Observations.hpp:48-54): the packing falls back to the plain columns / keeps the offending values."""
    from methylasso_b200 import api
    t = synth.chromosome_table(5_000, 11, (2,))
    ptr = np.array([0, t["pos"].size], np.int64)
    p = MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
    for col, val, code in (("Nmeth", -1, 11), ("coverage", -3, 11), ("Nmeth", 250, 12)):
        bad = {k: v.copy() for k, v in t.items()}
        bad[col][1234] = val
        res = engine.detect_batch(ptr, bad, p)
        try:
            assert res.status.tolist() == [code], (col, val, res.status.tolist())
        finally:
            res.free()
