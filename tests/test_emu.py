"""CPU emulation of the device logic (tests/emu): the same inline functions the CUDA kernels call,
compiled with g++ -ffp-contract=off, checked against the oracle.  Catches logic errors in the
chunked / sandwiched DP before any GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from methylasso_b200 import synth
from conftest import pooled_series

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_flsa.cpp")
LIB = os.path.join(HERE, "emu", "libemu_flsa.so")


@pytest.fixture(scope="module")
def emu():
    deps = [SRC] + [os.path.join(HERE, "..", "methylasso_b200", "csrc", f)
                    for f in ("flsa_core.cuh", "irls_core.cuh", "hd.h")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", LIB, SRC])
    L = C.CDLL(LIB)
    L.emu_set_fine(0)
    L.emu_clamp35_soft.restype = C.c_double
    L.emu_clamp35_soft.argtypes = [C.c_double, C.c_double]

    def run(y, w, lam, chunk=512, warm=512, tile=2048):
        y, w = np.ascontiguousarray(y, float), np.ascontiguousarray(w, float)
        n = y.size
        beta, tm, tp = np.zeros(n), np.zeros(n + 1), np.zeros(n + 1)
        st = np.zeros(8, np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = L.emu_flsa(n, p(y), p(w), C.c_double(lam), chunk, warm, p(beta), p(tm), p(tp), tile, p(st))
        assert rc == 0
        return beta, st
    def scan(y, w, lam, chunk=64, group=8):
        """the scan route (flsa_core.cuh): the five steps of flsa.cu's scan kernels, thread by thread"""
        y, w = np.ascontiguousarray(y, float), np.ascontiguousarray(w, float)
        n = y.size
        beta, tm, tp = np.zeros(n), np.zeros(n + 1), np.zeros(n + 1)
        st = np.zeros(8, np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = L.emu_flsa_scan(n, p(y), p(w), C.c_double(lam), chunk, group, p(beta), p(tm), p(tp), 2048, p(st))
        assert rc == 0
        return beta, st
    run.lib = L
    run.scan = scan
    return run


@pytest.mark.parametrize("reps", [(1,), (3,)])
@pytest.mark.parametrize("lam", [10.0, 25.0, 1000.0])
def test_chunked_dp_bit_exact_on_count_data(emu, oracle, reps, lam):
    """On count-valued pseudodata (the FAST path) every partial sum is exact, so the chunked run
    reproduces the sequential reference bit for bit whatever the chunking."""
    t = synth.chromosome_table(120_000, 31, reps)
    y, w = pooled_series(t)
    ref, _, _ = oracle.tf_dp_weight(y, w, lam, "port")
    for chunk, warm in ((512, 512), (256, 64), (4096, 2048)):
        beta, st = emu(y, w, lam, chunk, warm)
        assert np.array_equal(beta, ref), (chunk, warm, st.tolist())
        assert st[4] == 0  # no ring overflow on WGBS-like data


def test_chunked_dp_generic_inputs_within_rounding(emu, oracle):
    """Real-valued weights: exact in exact arithmetic; association of a few sums may differ."""
    rng = np.random.default_rng(0)
    worst = 0.0
    for trial in range(400):
        n = int(rng.integers(2, 500))
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
        lam = float(rng.choice([1e-3, 0.1, 0.5, 5.0, 25.0, 1000.0]))
        ref, _, _ = oracle.tf_dp_weight(y, w, lam, "port")
        beta, st = emu(y, w, lam, int(rng.choice([1, 2, 4, 16, 64])), int(rng.choice([1, 2, 8, 32])),
                       int(rng.choice([1, 3, 8, 64])))
        scale = max(1.0, np.max(np.abs(ref[np.isfinite(ref)])) if np.isfinite(ref).any() else 1.0)
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(beta), ok)
        if ok.any():
            worst = max(worst, np.max(np.abs(beta[ok] - ref[ok])) / scale)
    assert worst < 1e-11, worst


@pytest.mark.parametrize("reps", [(1,), (3,)])
def test_scan_route_bit_exact_on_count_data(emu, oracle, reps):
    """Chunks as maps on messages, composed by merges of knot lists: on count data with an integer penalty every sum is
    exact and every crossing the correctly rounded quotient of the numbers the sequential run divides - the same bits,
    whatever the chunk and group lengths.  A fractional penalty re-associates sums like the speculative route does."""
    t = synth.chromosome_table(120_000, 31, reps)
    y, w = pooled_series(t)
    for lam in (10.0, 25.0, 1000.0):
        ref, _, _ = oracle.tf_dp_weight(y, w, lam, "port")
        for chunk, group in ((64, 16), (448, 12), (7, 3)):
            beta, st = emu.scan(y, w, lam, chunk, group)
            assert st[4] == 0, (lam, chunk, group)          # no fallback on WGBS-like data
            assert np.array_equal(beta, ref), (lam, chunk, group)
    ref, _, _ = oracle.tf_dp_weight(y, w, 17.3, "port")
    beta, st = emu.scan(y, w, 17.3, 64, 16)
    assert st[4] == 0 and np.max(np.abs(beta - ref)) < 1e-11


def test_fine_chains_of_the_speculative_route(emu, oracle):
    """The chains of unverified chunks by composition (eight sub-chunks per chunk, bounds / start messages / redo) instead of
    a walk: bit-equal on count data with integer penalties whatever the chunking, within rounding on arbitrary inputs; a
    chain the merges cannot take (irregular start window, a message of more than 32 knots) is left to the walker."""
    emu.lib.emu_set_fine(1)
    try:
        t = synth.chromosome_table(120_000, 31, (3,))
        y, w = pooled_series(t)
        for lam in (25.0, 1000.0):
            ref, _, _ = oracle.tf_dp_weight(y, w, lam, "port")
            for chunk, warm in ((448, 512), (256, 64), (64, 16)):
                beta, st = emu(y, w, lam, chunk, warm)
                assert np.array_equal(beta, ref), (lam, chunk, warm)
                assert st[4] == 0
        rng = np.random.default_rng(0)
        worst = 0.0
        for trial in range(300):
            n = int(rng.integers(2, 500))
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
            lam = float(rng.choice([1e-3, 0.1, 0.5, 5.0, 25.0, 1000.0]))
            ref, _, _ = oracle.tf_dp_weight(y, w, lam, "port")
            beta, st = emu(y, w, lam, int(rng.choice([1, 2, 4, 16, 64])), int(rng.choice([1, 2, 8, 32])), int(rng.choice([1, 3, 8, 64])))
            ok = np.isfinite(ref)
            assert np.array_equal(np.isfinite(beta), ok), (trial, kind)
            if ok.any():
                worst = max(worst, np.max(np.abs(beta[ok] - ref[ok])) / max(1.0, np.max(np.abs(ref[ok]))))
        assert worst < 1e-11, worst
    finally:
        emu.lib.emu_set_fine(0)


def test_scan_route_generic_inputs_within_rounding(emu, oracle):
    """Real-valued data, zero weights inside the series and AT ITS HEAD (the reference's first-element formulas with
    w = 0 leave knots at infinity and then a NaN knot: the scan keeps that window literally), every chunk / group length;
    monotone series outgrow a message's capacity and are finished start to end."""
    rng = np.random.default_rng(0)
    worst, fallbacks, zero_heads = 0.0, 0, 0
    for trial in range(500):
        n = int(rng.integers(2, 500))
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
        lam = float(rng.choice([1e-3, 0.1, 0.5, 5.0, 25.0, 1000.0]))
        ref, _, _ = oracle.tf_dp_weight(y, w, lam, "port")
        beta, st = emu.scan(y, w, lam, int(rng.choice([1, 2, 4, 16, 64])), int(rng.choice([1, 2, 3, 8])))
        if kind != 1:
            assert st[4] == 0, (trial, kind)
            zero_heads += int(w[0] == 0)
        fallbacks += int(st[4])
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(beta), ok), (trial, kind)
        if ok.any():
            worst = max(worst, np.max(np.abs(beta[ok] - ref[ok])) / max(1.0, np.max(np.abs(ref[ok]))))
    assert worst < 1e-11, worst
    assert zero_heads > 50 and 0 < fallbacks < 100


def test_side_symmetric_step_is_bit_exact_on_any_input(emu, oracle):
    """One chunk = the whole series: the side-symmetric step (both scans as `a*x + b > -lam`, knot
    (-lam - b)/a, crossing redo) must reproduce the reference's mirrored formulas bit for bit."""
    rng = np.random.default_rng(11)
    for trial in range(300):
        n = int(rng.integers(2, 400))
        y = rng.normal(size=n) * float(rng.choice([1e-3, 1.0, 1e3]))
        w = np.where(rng.random(n) < 0.25, 0.0, rng.exponential(size=n))
        lam = float(rng.choice([1e-6, 1e-2, 0.7, 30.0, 1e4]))
        ref, _, _ = oracle.tf_dp_weight(y, w, lam, "port")
        beta, st = emu(y, w, lam, 1 << 30, 1)
        assert np.array_equal(beta, ref, equal_nan=True), (trial, n, lam)


def test_chunked_dp_edge_sizes(emu, oracle):
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 4, 5, 6):
        y, w = rng.random(n), rng.integers(1, 9, n).astype(float)
        for lam in (0.0, 0.5, 50.0):
            ref, _, _ = oracle.tf_dp_weight(y, w, lam, "port")
            beta, _ = emu(y, w, lam, 2, 1, 2)
            assert np.array_equal(beta, ref)


def test_knot_ring_overflow_falls_back(emu, oracle):
    """Strictly convex data with a tiny penalty keeps every knot alive: the 64-knot ring overflows and
    the series is redone start-to-end with the reference's 2n scratch layout."""
    n = 400
    y = (np.arange(n) / n) ** 2 * 50
    w = np.ones(n)
    ref, _, _ = oracle.tf_dp_weight(y, w, 200.0, "port")
    beta, st = emu(y, w, 200.0, 64, 16)
    assert st[4] > 0
    assert np.array_equal(beta, ref)


def test_even_bin_matches_oracle_rule(emu):
    rng = np.random.default_rng(5)
    for lower, upper, bf in ((10_000.0, 3_400_000.0, 50.0), (5.0, 6.0, 2000.0), (0.0, 1e9, 0.004), (7.0, 7.0, 5000.0)):
        size = upper - lower
        nb = int((size + 1) * bf / 1000.0)
        if nb < 1:
            continue
        v = np.concatenate([rng.integers(int(lower), int(upper) + 1, 5000).astype(float),
                            [lower, upper], lower + np.arange(min(nb, 3000)) * size / nb])
        v = v[(v >= lower) & (v <= upper)]
        out = np.zeros(v.size, np.uint32)
        emu.lib.emu_even_bin(v.size, v.ctypes.data_as(C.c_void_p), C.c_double(lower), C.c_double(size),
                             C.c_uint(nb), out.ctypes.data_as(C.c_void_p))
        bounds = lower + np.arange(nb + 1) * size / float(nb)
        want = np.searchsorted(bounds, v, side="right").astype(np.int64) - 1
        want[np.searchsorted(bounds, v, side="right") > nb] = nb - 1
        if size > 0:
            assert np.array_equal(out.astype(np.int64), want)
        else:
            assert np.all(out == nb - 1)


def test_clamp_soft_threshold_semantics(emu):
    f = emu.lib.emu_clamp35_soft
    assert f(40.0, 0.0) == 35.0 and f(-40.0, 0.0) == -35.0 and f(0.3, 0.0) == 0.3
    assert f(0.3, 0.5) == 0.0 and f(-0.3, 0.5) == 0.0 and f(2.0, 0.5) == 1.5 and f(-2.0, 0.5) == -1.5
    assert f(float("nan"), 0.0) == -35.0   # std::max(-35, NaN) keeps -35 (reference behaviour)


def test_x87_mean_matches_host_long_double():
    """methylasso_b200/csrc/x87.cuh (R's mean() with its 80-bit roundings, used for the per-position means of the DMR
    caller) against the host's long double on methylation-like ratios and on random doubles."""
    src = os.path.join(HERE, "emu", "emu_x87.cpp")
    lib = os.path.join(HERE, "emu", "libemu_x87.so")
    dep = os.path.join(HERE, "..", "methylasso_b200", "csrc", "x87.cuh")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", lib, src])
    L = C.CDLL(lib)
    if L.ref_long_double_digits() != 64:
        pytest.skip("the host's long double is not the x87 80-bit format")
    L.emu_x87_compare.restype = C.c_longlong
    L.emu_x87_compare.argtypes = [C.c_longlong, C.c_int, C.c_void_p, C.POINTER(C.c_longlong)]
    rng = np.random.default_rng(1)
    for n in range(1, 9):
        m = 100_000
        cov = rng.integers(1, 60, (m, n))
        x = np.ascontiguousarray(rng.binomial(cov, rng.random((m, n))) / cov)
        first = C.c_longlong()
        assert L.emu_x87_compare(m, n, x.ctypes.data, C.byref(first)) == 0, (n, x[first.value])
        x = np.ascontiguousarray(rng.random((m, n)) * rng.choice([1e-3, 1.0, 1e3], (m, 1)))
        assert L.emu_x87_compare(m, n, x.ctypes.data, C.byref(first)) == 0, (n, x[first.value])


def test_device_tile_tables_match_host_loops():
    """methylasso_b200/csrc/tiles.cuh: the span -> tile arithmetic of k_fill_tiles against the host loops it replaced."""
    src = os.path.join(HERE, "emu", "emu_tiles.cpp")
    lib = os.path.join(HERE, "emu", "libemu_tiles.so")
    deps = [src] + [os.path.join(HERE, "..", "methylasso_b200", "csrc", f) for f in ("tiles.cuh", "hd.h")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", lib, src])
    L = C.CDLL(lib)
    L.emu_tiles_check.restype = C.c_longlong
    L.emu_tiles_check.argtypes = [C.c_int, C.c_void_p]
    rng = np.random.default_rng(5)
    cases = [np.array([[0, 0, 256, 0]]), np.array([[1, 0, 448, 1]]),                        # nothing / one empty chunk
             np.array([[1, 7, 2**31 - 1, 1]]),                                           # sequential mode: one chunk
             np.array([[5, 1024, 1024, 0], [0, 0, 1024, 0], [9, 1025, 1024, 0]])]         # exact fit, empty, one over
    for _ in range(200):
        k = int(rng.integers(1, 40))
        d = np.stack([rng.integers(0, 10**9, k), rng.choice([0, 1, 2, 255, 256, 257, 5000, 123457], k),
                      rng.choice([1, 64, 256, 448, 1024, 2048], k), rng.integers(0, 2, k)], axis=1)
        cases.append(d)
    for d in cases:
        d = np.ascontiguousarray(d, np.int64)
        assert L.emu_tiles_check(d.shape[0], d.ctypes.data) == 0, d.tolist()
