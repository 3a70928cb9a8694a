"""GPU parity of the batched weighted 1-D FLSA (through the C ABI) against the oracle."""
import os

import numpy as np
import pytest

from methylasso_b200 import synth
from conftest import pooled_series
from test_oracle import kkt_residual

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_known_answers(engine):
    assert np.allclose(engine.weighted_1d_flsa([0, 0, 1, 1], [1, 1, 1, 1], 0.5), [.25, .25, .75, .75], atol=1e-15)
    assert np.allclose(engine.weighted_1d_flsa([0, 0, 1, 1], [1, 1, 1, 1], 10), [.5, .5, .5, .5], atol=1e-15)


def test_golden_reference_vectors(engine):
    """Vectors produced by the reference's verbatim tf_dp_weight (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, "dp_vectors.npz"))
    for k in ("random", "zero_w", "monotone", "plateaus", "known"):
        beta = engine.weighted_1d_flsa(g[f"{k}_y"], g[f"{k}_w"], float(g[f"{k}_lam"]))
        want = np.clip(g[f"{k}_beta"], -35, 35)
        assert np.array_equal(beta, want, equal_nan=True), k


def test_reference_flsa_fixture(engine):
    """weighted_1d_flsa outputs of the reference's own C++ entry point (src/weighted_1d_flsa.cpp:7),
    tests/golden/reference_fixtures.npz: integer weights with zeros, with and without lambda1."""
    g = np.load(os.path.join(GOLD, "reference_fixtures.npz"))
    for lam2, lam1 in ((25.0, 0.0), (3.0, 0.05)):
        got = engine.weighted_1d_flsa(g["flsa_y"], g["flsa_w"], lam2, lam1)
        assert np.array_equal(got, g[f"flsa_{lam2:g}_{lam1:g}"])


def test_trivial_and_error_cases(engine, oracle):
    from methylasso_b200.api import MethyLassoError
    assert engine.weighted_1d_flsa([], [], 5.0).size == 0
    assert np.array_equal(engine.weighted_1d_flsa([0.3], [2.0], 5.0), [0.3])
    y = np.array([1.0, 50.0, -70.0])
    assert np.array_equal(engine.weighted_1d_flsa(y, np.ones(3), 0.0), [1.0, 35.0, -35.0])  # lam == 0: beta = y, clamped
    assert np.array_equal(engine.weighted_1d_flsa(y, np.ones(3), 0.0, 2.0), [0.0, 33.0, -33.0])  # soft threshold
    with pytest.raises(MethyLassoError) as e:
        engine.weighted_1d_flsa([1.0, 2.0], [1.0], 1.0)
    assert e.value.code == 2


@pytest.mark.parametrize("reps", [(1,), (3,)])
def test_count_data_bit_exact(engine, oracle, reps):
    t = synth.chromosome_table(400_000, 20241, reps)
    y, w = pooled_series(t)
    for lam in (10.0, 25.0, 1000.0, 2000.0):
        ref = oracle.weighted_1d_flsa(y, w, lam)
        got = engine.weighted_1d_flsa(y, w, lam)
        assert np.array_equal(got, ref), lam


def test_routes_on_count_data_bit_exact(engine, oracle):
    """Both forward passes on the same series: the speculative route forced onto a short series and the scan route
    (chunks as maps on messages, flsa_core.cuh) forced onto a long one.  Count data, integer penalties: bit-equal to the
    reference's verbatim DP either way; a fractional penalty within rounding."""
    t = synth.chromosome_table(400_000, 20241, (3,))
    y, w = pooled_series(t)
    ys, ws = y[:150_000], w[:150_000]
    try:
        for route, (yy, ww) in ((2, (y, w)), (1, (ys, ws)), (0, (ys, ws))):
            engine.set("flsa_route", route)
            for lam in (25.0, 1000.0):
                assert np.array_equal(engine.weighted_1d_flsa(yy, ww, lam), oracle.weighted_1d_flsa(yy, ww, lam)), (route, lam)
            got, ref = engine.weighted_1d_flsa(yy, ww, 17.3), oracle.weighted_1d_flsa(yy, ww, 17.3)
            assert float(np.max(np.abs(got - ref))) < 1e-10, route
    finally:
        engine.set("flsa_route", 0)


def test_walkers_hand_a_series_over_to_the_scan_route(engine, oracle):
    """A long dense series takes the speculative route; when its walkers meet a chain of more than flsa_chain_limit
    chunks the series is handed over to the scan route ON THE DEVICE (its tables are built next to the speculative ones).
    Count data with the limit at 1 (every chain of two chunks triggers it): still bit-equal to the verbatim DP.  Real-valued
    data (windows from opposite bounds agree numerically long before they agree bit for bit): per element to 1e-10."""
    import ctypes as C

    def stats():
        out = (C.c_int * 7)()
        engine._check(engine.lib.ml_debug_flsa_last_stats(engine.h, out))
        return dict(zip(("verified", "redone", "longest", "chains", "handed_over", "scan_chunks", "fallback"), out))
    t = synth.chromosome_table(400_000, 20241, (3,))
    y, w = pooled_series(t)
    ref = oracle.weighted_1d_flsa(y, w, 25.0)
    try:
        longest = None
        for limit in (0, 1, 6):
            engine.set("flsa_chain_limit", limit)
            assert np.array_equal(engine.weighted_1d_flsa(y, w, 25.0), ref), limit
            st = stats()
            assert st["fallback"] == 0 and st["verified"] > 0, st
            if limit == 0:                                  # no hand-over, no scan tables: the walkers finish every chain
                assert st["handed_over"] == 0 and st["scan_chunks"] == 0, st
                longest = st["longest"]
            else:                                           # handed over exactly when some chain is longer than the limit
                assert st["handed_over"] == int(longest > limit), (limit, longest, st)
        rng = np.random.default_rng(5)
        n = 300_000
        yr = np.cumsum(rng.normal(0, 0.01, n)) + rng.normal(0, 0.3, n)
        wr = rng.exponential(20.0, n)
        want = oracle.weighted_1d_flsa(yr, wr, 20.0)
        engine.set("flsa_chain_limit", 0)
        got0 = engine.weighted_1d_flsa(yr, wr, 20.0)
        longest = stats()["longest"]
        engine.set("flsa_chain_limit", 6)
        got = engine.weighted_1d_flsa(yr, wr, 20.0)
        assert stats()["handed_over"] == int(longest > 6), (longest, stats())
        assert longest > 6                                  # real-valued data: windows rarely turn bit-identical in time
    finally:
        engine.set("flsa_chain_limit", 6)
    assert np.all(np.abs(got0 - want) <= 1e-10 * max(1.0, float(np.max(np.abs(want)))))
    assert np.all(np.abs(got - want) <= 1e-10 * max(1.0, float(np.max(np.abs(want)))))


def test_scan_route_on_pmd_like_series(engine, oracle):
    """The PMD step's shape (R/call.R:93): a few 10^4 real-valued standard deviations weighted by segment widths, lambda2 =
    1000 - the series whose memory no warm-up covers.  Scan route (the default for short series) against the verbatim DP:
    per element to 1e-12 of the scale, whatever the chunk length; with zero weights at the head and inside."""
    rng = np.random.default_rng(77)
    n = 60_000
    y = np.abs(rng.normal(0.15, 0.08, n)) * (1 + (np.arange(n) // 7000) % 2)
    w = np.round(rng.lognormal(5.5, 1.0, n))
    ref = oracle.weighted_1d_flsa(y, w, 1000.0)
    try:
        for chunk in (0, 32, 448):
            engine.set("flsa_chunk", chunk)
            got = engine.weighted_1d_flsa(y, w, 1000.0)
            assert np.all(np.abs(got - ref) <= 1e-12 * max(1.0, float(np.max(np.abs(ref))))), chunk
    finally:
        engine.set("flsa_chunk", 0)
    w2 = w.copy()
    w2[:3] = 0.0
    w2[rng.random(n) < 0.7] = 0.0
    got, ref = engine.weighted_1d_flsa(y, w2, 1000.0), oracle.weighted_1d_flsa(y, w2, 1000.0)
    ok = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), ok)
    assert np.all(np.abs(got[ok] - ref[ok]) <= 1e-12 * max(1.0, float(np.max(np.abs(ref[ok])))))


def test_generic_inputs_and_ragged_batch(engine, oracle):
    rng = np.random.default_rng(0)
    ys, ws, lams, ptr = [], [], [], [0]
    for trial in range(300):
        n = int(rng.integers(0, 3000)) if trial % 7 else int(rng.integers(0, 4))
        kind = trial % 5
        if kind == 0:
            y, w = rng.random(n), rng.integers(0, 3, n).astype(float)
        elif kind == 1:
            y, w = np.sort(rng.random(n)), rng.random(n)
        elif kind == 2:
            y, w = np.repeat(rng.random(n // 7 + 1), 7)[:n], rng.integers(0, 50, n).astype(float)
        elif kind == 3:
            y = rng.integers(0, 2, n).astype(float)
            w = np.where(rng.random(n) < 0.8, 0.0, rng.integers(1, 30, n)).astype(float)
        else:
            y, w = rng.normal(size=n), rng.exponential(size=n)
        ys.append(y); ws.append(w); ptr.append(ptr[-1] + n)
        lams.append(float(rng.choice([0.0, 1e-3, 0.1, 0.5, 5.0, 25.0, 1000.0])))
    yy, ww = np.concatenate(ys), np.concatenate(ws)
    got = engine.weighted_1d_flsa_batch(np.array(ptr), yy, ww, np.array(lams))
    worst = 0.0
    for s in range(len(ys)):
        ref = oracle.weighted_1d_flsa(ys[s], ws[s], lams[s])
        g = got[ptr[s]:ptr[s + 1]]
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(g), ok), s
        if ok.any():
            worst = max(worst, np.max(np.abs(g[ok] - ref[ok])) / max(1.0, np.max(np.abs(ref[ok]))))
    assert worst < 1e-11, worst   # tolerance: rounding of re-associated sums; north_star allows 1e-6


def test_ring_overflow_falls_back_to_global_scratch(engine, oracle):
    n = 5000
    y = (np.arange(n) / n) ** 2 * 500
    w = np.ones(n)
    assert np.array_equal(engine.weighted_1d_flsa(y, w, 200.0), oracle.weighted_1d_flsa(y, w, 200.0))


def test_sequential_mode_is_bit_faithful_on_real_valued_data(engine, oracle):
    """flsa_sequential = 1: one thread walks each series start to end -> the reference's operation
    order exactly, for any input."""
    rng = np.random.default_rng(9)
    y, w = rng.normal(size=20_000), rng.exponential(size=20_000)
    engine.set("flsa_sequential", 1)
    try:
        got = engine.weighted_1d_flsa(y, w, 3.0)
    finally:
        engine.set("flsa_sequential", 0)
    assert np.array_equal(got, oracle.weighted_1d_flsa(y, w, 3.0))


def test_full_size_chromosome_against_the_verbatim_dp(engine, oracle):
    """chr1-sized series (2.26 M CpGs, count data): bit-equal to the reference's own tf_dp_weight
    (src/core/gfl/gfl_tf.c:184-321, compiled verbatim under oracle/_ref when present, else its restatement) at the
    UMR/LMR penalty and at the PMD penalty; optimal (KKT); chunked == sequential mode."""
    t = synth.chromosome_table(2_260_000, 99, (1,))
    y, w = pooled_series(t)
    for lam in (25.0, 1000.0):
        beta = engine.weighted_1d_flsa(y, w, lam)
        assert np.array_equal(beta, oracle.weighted_1d_flsa(y, w, lam)), lam
        assert kkt_residual(y, w, beta, lam) < 1e-8
    beta = engine.weighted_1d_flsa(y, w, 25.0)
    engine.set("flsa_sequential", 1)
    try:
        seq = engine.weighted_1d_flsa(y, w, 25.0)
    finally:
        engine.set("flsa_sequential", 0)
    assert np.array_equal(beta, seq)


def test_lambda_sweep_config5(engine, oracle):
    """BASELINE config 5: 16 lambda2 values in [10, 2000] on the same series, one batched call."""
    t = synth.chromosome_table(150_000, 5, (3,))
    y, w = pooled_series(t)
    lams = np.geomspace(10, 2000, 16)
    lams[::4] = np.round(lams[::4])          # every 4th value an integer, as the reference's own defaults are
    n = y.size
    ptr = np.arange(17) * n
    got = engine.weighted_1d_flsa_batch(ptr, np.tile(y, 16), np.tile(w, 16), lams)
    for i, lam in enumerate(lams):
        ref = oracle.weighted_1d_flsa(y, w, float(lam))
        g = got[i * n:(i + 1) * n]
        # counts and an integer penalty: every partial sum of the DP is exact -> bit-identical to the
        # sequential reference; a fractional penalty re-associates a few sums (measured 7.5e-12 at
        # lambda2 = 14.24; the north star's tolerance is 1e-6 relative)
        if float(lam).is_integer():
            assert np.array_equal(g, ref), lam
        assert float(np.max(np.abs(g - ref))) < 1e-10, lam
        assert kkt_residual(y, w, g, float(lam)) < 1e-8


def test_lambda_sweep_on_the_resident_fit(engine, oracle):
    """BASELINE config 5 as the engine runs it: ml_result_lambda_sweep - the 16 lambda2 lanes of every chromosome read ONE
    resident pseudodata stream (no tiling of y / w).  Every lane equals the single-lambda2 fit bit for bit (same FLSA, same
    centring), its segment count the single fit's, and the port's coefficients to 1e-10."""
    from methylasso_b200._native import MlParams
    tables = [synth.chromosome_table(n, 770 + i, (3,)) for i, n in enumerate((40_000, 25_000))]
    ptr, cols = synth.concat_jobs(tables)
    lams = np.geomspace(10, 2000, 16)
    lams[::4] = np.round(lams[::4])
    res = engine.detect_batch(ptr, cols, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        nseg, beta = res.lambda_sweep(lams, 0.01, want_beta=True)
        sizes = [res.sizes(j)[1] for j in range(res.njobs)]
    finally:
        res.free()
    for i in (0, 5, 12, 15):
        lam = float(lams[i])
        one = engine.detect_batch(ptr, cols, MlParams(lam, lam + 1, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
        try:
            at = 0
            for j, P in enumerate(sizes):
                assert np.array_equal(beta[i, at:at + P], one.job(j)["estimates"]["beta"]), (lam, j)
                want = oracle.detect(tables[j], oracle.MODEL_FAST, lambda2=lam)["estimates"]["beta"]
                assert np.all(np.abs(beta[i, at:at + P] - want) <= 1e-10 * np.maximum(np.abs(want), 1e-3)), (lam, j)
                at += P
            assert int(nseg[i]) == one.segments_count(0.01), lam
        finally:
            one.free()
    assert np.all(np.diff(nseg) <= 0) or nseg[0] > nseg[-1]       # a larger penalty fuses more


def test_hypothesis_kkt_and_oracle(engine, oracle):
    """Property test on arbitrary inputs: optimality conditions + agreement with the oracle."""
    from hypothesis import given, settings, strategies as st
    import hypothesis.extra.numpy as hnp

    @settings(max_examples=60, deadline=None)
    @given(st.integers(2, 300).flatmap(lambda n: st.tuples(
        hnp.arrays(np.float64, n, elements=st.floats(-5, 5, allow_nan=False, width=32)),
        hnp.arrays(np.float64, n, elements=st.floats(0.0, 50.0, allow_nan=False, width=32)),
        st.booleans(),
        st.sampled_from([0.01, 0.5, 7.0, 300.0]))))
    def check(args):
        y, w, zeros_allowed, lam = args
        if not zeros_allowed:
            w = np.maximum(w, 0.015625)
        if not (w > 0).any():
            return
        got = engine.weighted_1d_flsa(y, w, lam)
        ref = oracle.weighted_1d_flsa(y, w, lam)
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), ok)
        assert np.allclose(got[ok], ref[ok], rtol=1e-9, atol=1e-9)
        # optimality is a property of the problem with positive weights; with zero weights the
        # reference DP itself is only checked for agreement (e.g. y=[0,0,0], w=[0,0,1] gives -lambda)
        if not zeros_allowed and ok.all() and np.all(np.abs(ref) < 35):
            assert kkt_residual(y, w, got, lam) < 1e-6 * max(1.0, float(np.max(w)) * float(np.max(np.abs(y))) / lam)
    check()
