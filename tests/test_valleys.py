"""Valleys / DMV of R/call.R:100-133: the numpy restatement (oracle/valleys.py) on hand-made tables, the device logic
(csrc/valley_core.cuh) run on the CPU by tests/emu/emu_valleys.cpp against it, and the CUDA path against it on resident
call tables (GPU)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from methylasso_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_valleys.cpp")
LIB = os.path.join(HERE, "emu", "libemu_valleys.so")
VCOLS = (("job", np.int32), ("condition", np.int32), ("start", np.uint32), ("end", np.uint32), ("width", np.float64),
         ("beta", np.float64), ("coverage", np.float64), ("pseudoweight", np.float64), ("num_cpgs", np.int32))


def table(beta, width=None, gap=None, ncpg=None, cond=None, job=None, pw=None, cov=None):
    n = len(beta)
    width = np.full(n, 3000.0) if width is None else np.asarray(width, float)
    gap = np.full(n, 100.0) if gap is None else np.asarray(gap, float)        # gap[i]: distance from row i-1's end
    start = np.zeros(n)
    end = np.zeros(n)
    cond = np.ones(n, np.int32) if cond is None else np.asarray(cond, np.int32)
    job = np.zeros(n, np.int32) if job is None else np.asarray(job, np.int32)
    for i in range(n):
        new = i == 0 or cond[i] != cond[i - 1] or job[i] != job[i - 1]
        start[i] = 1000 if new else end[i - 1] + gap[i]
        end[i] = start[i] + width[i] - 1
    return dict(job=job, estimate_id=cond, start=start.astype(np.uint32), end=end.astype(np.uint32), width=width,
                beta=np.asarray(beta, float), num_cpgs=np.full(n, 12, np.int32) if ncpg is None else np.asarray(ncpg, np.int32),
                coverage=np.full(n, 240.0) if cov is None else np.asarray(cov, float),
                pseudoweight=np.full(n, 100.0) if pw is None else np.asarray(pw, float))


def test_oracle_hand_examples(oracle):
    from oracle import valleys as V
    # high | low low (6099 wide: a DMV) | high | low (3000 wide: "low") | high; everything within 5000 of its neighbour
    t = table([0.9, 0.02, 0.04, 0.9, 0.05, 0.9])
    r = V.valleys(t, np.zeros(6, np.int8))
    low = r["low"]
    assert low["is_high"].tolist() == [True, False, False, True, False, True]
    assert low["low_id"].tolist() == [1, 1, 1, 2, 2, 3]
    assert low["low_id_width"][1:3].tolist() == [6099.0, 6099.0] and low["low_id_width"][4] == 3000.0
    assert np.isnan(low["low_id_width"][[0, 3, 5]]).all()
    assert low["low_id_beta"][1] == (0.02 + 0.04) / 2 and low["low_id_beta"][4] == 0.05
    v = r["valleys"]
    # the DMV and the "low" region lie 3199 apart: one valley from row 1 to row 4; the high row between them is inside
    # its span but contributes nothing to the sums
    assert v["start"].tolist() == [float(t["start"][1])] and v["end"].tolist() == [float(t["end"][4])]
    assert v["width"].tolist() == [4 * 3000.0 + 3 * 99.0] and v["num_cpgs"].tolist() == [36]
    assert v["beta"][0] == np.mean([0.03, 0.05]) and v["coverage"].tolist() == [720.0] and v["pseudoweight"].tolist() == [300.0]
    assert r["row_in_valley"].tolist() == [False, True, True, True, True, False]
    # a gap wider than pmd_valley_min_width splits the run (follows.large.gap): two "low" regions, no DMV, no valley
    r = V.valleys(table([0.9, 0.02, 0.04, 0.9], gap=[0, 100, 6000, 100]), np.zeros(4, np.int8))
    assert r["low"]["low_id"].tolist() == [1, 1, 2, 3] and r["valleys"]["start"].size == 0 and not r["row_in_valley"].any()
    # a DMV whose CpGs are too sparse is dropped (:125): 6099 / (4 - 1) > 500
    r = V.valleys(table([0.9, 0.02, 0.04, 0.9], ncpg=[12, 2, 2, 12]), np.zeros(4, np.int8))
    assert r["valleys"]["start"].size == 0
    # conditions do not see each other: the second condition's first row opens a run of its own
    t = table([0.9, 0.02, 0.04, 0.03, 0.02, 0.9], cond=[1, 1, 1, 2, 2, 2])
    r = V.valleys(t, np.zeros(6, np.int8))
    assert r["valleys"]["condition"].tolist() == [1, 2] and r["valleys"]["start"].tolist() == [float(t["start"][1]), 1000.0]
    # an UMR above valley_max_beta (possible when valley_max_beta < umr_max_beta) overlaps no low region: it joins
    # new_valleys with NA width / beta (:114-115) and the merged valley's mean(beta) becomes NA
    t = table([0.9, 0.02, 0.03, 0.9, 0.08, 0.9])
    r = V.valleys(t, np.array([0, 0, 0, 0, 2, 0], np.int8), valley_max_beta=0.05)
    assert r["valleys"]["start"].tolist() == [float(t["start"][1])] and r["valleys"]["end"].tolist() == [float(t["end"][4])]
    assert np.isnan(r["valleys"]["beta"][0]) and r["valleys"]["num_cpgs"].tolist() == [36]
    # the same UMR lying inside a low region adds nothing
    r2 = V.valleys(t, np.array([0, 0, 2, 0, 0, 0], np.int8), valley_max_beta=0.05)
    assert r2["valleys"]["end"].tolist() == [float(t["end"][2])] and r2["valleys"]["beta"][0] == 0.025
    # closed intervals: a row that only touches the valley's last base is replaced, one base further it is not
    t = table([0.02, 0.04, 0.9], gap=[0, 100, 0])          # gap 0: the high row starts on the valley's end
    assert t["start"][2] == t["end"][1]
    assert V.valleys(t, np.zeros(3, np.int8))["row_in_valley"].tolist() == [True, True, True]
    t = table([0.02, 0.04, 0.9], gap=[0, 100, 1])
    assert V.valleys(t, np.zeros(3, np.int8))["row_in_valley"].tolist() == [True, True, False]


@pytest.fixture(scope="module")
def emu():
    deps = [SRC] + [os.path.join(HERE, "..", "methylasso_b200", "csrc", f) for f in ("valley_core.cuh", "x87.cuh", "hd.h")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", LIB, SRC])
    L = C.CDLL(LIB)
    L.emu_valleys.restype = C.c_int64
    L.emu_pmd_split.restype = C.c_int64
    L.emu_x87_sum.restype = C.c_double
    L.ref_x87_sum.restype = C.c_double

    def run(t, category, valley_max_beta=0.1, pmd_valley_min_width=5000.0, max_distance=500.0, min_num_cpgs=4):
        n = t["start"].size
        cols = [np.ascontiguousarray(t[k], dt) for k, dt in (("job", np.int32), ("estimate_id", np.int32), ("num_cpgs", np.int32),
                                                              ("start", np.uint32), ("end", np.uint32), ("beta", np.float64),
                                                              ("coverage", np.float64), ("pseudoweight", np.float64))]
        cols.append(np.ascontiguousarray(category, np.int8))
        m = max(n, 1)
        out = {k: np.empty(m, dt) for k, dt in VCOLS}
        rows = dict(row_in_valley=np.empty(m, np.int8), low_id_width=np.empty(m), low_id_beta=np.empty(m))
        row0 = np.empty(m, np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        k = L.emu_valleys(C.c_int64(n), *[p(a) for a in cols], C.c_double(valley_max_beta), C.c_double(pmd_valley_min_width),
                          C.c_double(max_distance), C.c_int32(min_num_cpgs), *[p(out[c]) for c, _ in VCOLS],
                          *[p(rows[c]) for c in ("row_in_valley", "low_id_width", "low_id_beta")], p(row0))
        res = {c: a[:k] for c, a in out.items()}
        res["row0"] = row0[:k]
        res.update({c: a[:n] for c, a in rows.items()})
        return res
    run.lib = L
    return run


def check_against_oracle(got, t, category, **kw):
    """got: columns over all jobs (emulation or CUDA); the oracle works per chromosome."""
    from oracle import valleys as V
    n_valleys = 0
    for j in np.unique(t["job"]):
        m = t["job"] == j
        want = V.valleys({k: v[m] for k, v in t.items()}, np.asarray(category)[m], **kw)
        g = got["job"] == j
        w = want["valleys"]
        assert g.sum() == w["start"].size, (j, int(g.sum()), w["start"].size)
        assert np.array_equal(got["condition"][g], w["condition"])
        for k in ("start", "end", "width", "beta", "coverage", "pseudoweight", "num_cpgs"):
            assert np.array_equal(got[k][g].astype(np.float64), w[k].astype(np.float64), equal_nan=True), (j, k)
        assert np.array_equal(got["row_in_valley"][m] != 0, want["row_in_valley"]), j
        assert np.array_equal(got["low_id_width"][m], want["low"]["low_id_width"], equal_nan=True), j
        assert np.array_equal(got["low_id_beta"][m], want["low"]["low_id_beta"], equal_nan=True), j
        n_valleys += w["start"].size
    return n_valleys


def random_table(rng, n_jobs=3, adjacent=False):
    beta, width, gap, ncpg, cond, job, pw, cov = [], [], [], [], [], [], [], []
    for j in range(n_jobs):
        for c in range(1, int(rng.integers(1, 4))):
            n = int(rng.integers(1, 120))
            low = rng.random(n) < 0.55
            beta += list(np.where(low, rng.random(n) * 0.12, 0.2 + rng.random(n) * 0.8))
            width += list(rng.integers(1, 4000, n).astype(float) if not adjacent else rng.integers(1, 6, n).astype(float))
            # gap 0 / 1: closed intervals that touch (positions one base apart) or just miss each other
            gap += list(np.where(rng.random(n) < 0.1, rng.integers(4000, 9000, n), rng.integers(0, 3, n) if adjacent
                                 else rng.integers(0, 1500, n)).astype(float))
            ncpg += list(rng.integers(1, 60, n))
            cond += [c] * n
            job += [j] * n
            pw += list(rng.random(n) * 1e4)
            cov += list(rng.integers(5, 3000, n).astype(float))
    return table(beta, width, gap, ncpg, cond, job, pw, cov)


def test_emulated_device_logic_matches_oracle(emu, oracle):
    from oracle import lmr_umr as LU
    rng = np.random.default_rng(7)
    total = nan_beta = 0
    for trial in range(120):
        t = random_table(rng, adjacent=trial % 4 == 3)
        n = t["start"].size
        if trial % 3 == 0:      # random UMR flags, also on high rows (valley_max_beta < umr_max_beta)
            category = np.where(rng.random(n) < 0.2, 2, 0).astype(np.int8)
        else:
            category = np.zeros(n, np.int8)
            category[(t["beta"] < 0.1) & (rng.random(n) < 0.5)] = 2
        kw = [{}, dict(valley_max_beta=0.05, pmd_valley_min_width=2500.0), dict(pmd_valley_min_width=800.0, max_distance=200.0,
                                                                                   min_num_cpgs=20)][trial % 3]
        got = emu(t, category, **kw)
        total += check_against_oracle(got, t, category, **kw)
        nan_beta += int(np.isnan(got["beta"]).sum())
    assert total > 300 and nan_beta > 5
    # no rows at all; a table without any low row
    assert emu(table([]), np.zeros(0, np.int8))["start"].size == 0
    assert emu(table([0.9, 0.8]), np.zeros(2, np.int8))["start"].size == 0


def test_emulated_long_double_sum(emu):
    """sum() of doubles in R accumulates in long double (summary.c rsum): x87.cuh's accumulator against the host's."""
    rng = np.random.default_rng(3)
    L = emu.lib
    for trial in range(4000):
        n = int(rng.integers(1, 40))
        x = np.ascontiguousarray(rng.random(n) * 10.0 ** rng.integers(-3, 7, n))
        p = x.ctypes.data_as(C.c_void_p)
        assert L.emu_x87_sum(n, p) == L.ref_x87_sum(n, p), trial


@pytest.mark.gpu
@pytest.mark.parametrize("reps", [(2,), (2, 2)])
def test_gpu_valleys_match_oracle(engine, oracle, reps):
    from methylasso_b200._native import MlParams
    tables = [synth.chromosome_table(n, 700 + i, reps) for i, n in enumerate((60_000, 9_000, 25_000))]
    ptr, cols = synth.concat_jobs(tables)
    res = engine.detect_batch(ptr, cols, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        call = res.call_table(0.01, 500.0)
        counts = []
        for kw, lmr in (({}, {}), (dict(pmd_valley_min_width=1200.0), {}),
                        (dict(valley_max_beta=0.04, pmd_valley_min_width=800.0, min_num_cpgs=3), dict(umr_large_density=0.01))):
            category = res.call_lmr_umr(0.01, 500.0, **dict(lmr, **({"min_num_cpgs": kw["min_num_cpgs"]} if "min_num_cpgs" in kw else {})))["category"]
            got = res.call_valleys(0.01, 500.0, **kw, **lmr)
            assert got["row_in_valley"].size == call["start"].size
            counts.append(check_against_oracle(got, call, category, max_distance=500.0, **kw))
        assert counts[1] > 10 and counts[2] > 10, counts
    finally:
        res.free()


# ---- R/call.R:131-164: valleys into the call table, regions, pieces of the PMD candidate segments ------------------------
MCOLS = (("job", np.int32), ("condition", np.int32), ("segment_id", np.int32), ("start", np.uint32), ("end", np.uint32),
         ("category", np.int8), ("src_row", np.int32), ("src_valley", np.int32), ("follows_large_gap", np.int8))
PCOLS = (("job", np.int32), ("estimate_id", np.int32), ("segment_id", np.int32), ("start", np.uint32), ("end", np.uint32),
         ("fused_std", np.float64), ("pseudoweight", np.float64))


def test_pmd_split_oracle_hand_example(oracle):
    from oracle import pmd_split as PS, valleys as V
    # rows: high | UMR (low, narrow) | high | low low (a DMV, 6099 wide) | high high; one PMD candidate segment over
    # rows 0-2, a second one from row 3 on
    t = table([0.9, 0.03, 0.9, 0.02, 0.04, 0.9, 0.8], gap=[0, 100, 100, 6000, 100, 100, 100])
    category = np.array([0, 2, 0, 0, 0, 0, 1], np.int8)
    val = V.valleys(t, category)
    assert val["valleys"]["start"].tolist() == [float(t["start"][3])] and val["row_in_valley"].tolist() == [0, 0, 0, 1, 1, 0, 0]
    M = PS.merged_table(t, category, val)
    assert M["src_row"].tolist() == [0, 1, 2, -1, 5, 6] and M["src_valley"].tolist() == [-1, -1, -1, 0, -1, -1]
    assert M["category"].tolist() == [0, 2, 0, 3, 0, 1] and M["segment_id"].tolist() == [1, 2, 3, 4, 5, 6]
    assert M["follows_large_gap"].tolist() == [False, False, False, True, False, False]
    reg = PS.regions(M)
    # complement | isolate (the UMR) | complement | isolate (the valley, also behind a large gap) | complement
    assert reg["isolate"].tolist() == [False, True, False, True, False]
    assert reg["start"].tolist() == [float(t["start"][i]) for i in (0, 1, 2, 3, 5)]
    assert reg["end"].tolist() == [float(t["end"][i]) for i in (0, 1, 2, 4, 6)]
    std_call = dict(estimate_id=np.array([1, 1]), start=np.array([t["start"][0], t["start"][3]], float),
                    end=np.array([t["start"][3] - 1.0, float(t["start"][6])]), fused_std=np.array([0.2, 0.3]),
                    pseudoweight=np.array([7.0, 9.0]))
    sp = PS.split_call(reg, std_call)
    assert sp["start"].tolist() == [float(t["start"][i]) for i in (0, 1, 2, 3, 5)]
    assert sp["end"].tolist() == [float(t["end"][0]), float(t["end"][1]), float(t["end"][2]), float(t["end"][4]), float(t["start"][6])]
    assert sp["fused_std"].tolist() == [0.2, 0.0, 0.2, 0.0, 0.3] and sp["pseudoweight"].tolist() == [7.0, 7.0, 7.0, 9.0, 9.0]
    assert sp["segment_id"].tolist() == [1, 2, 3, 4, 5]
    # split.pmds = F: only the valley isolates
    reg = PS.regions(M, split_pmds=False)
    assert reg["isolate"].tolist() == [False, True, False]
    assert PS.split_call(reg, std_call)["fused_std"].tolist() == [0.2, 0.0, 0.3]


def random_std_call(rng, t):
    """PMD candidate segments as the helper forms them: a segment starts on a call-table row's start and ends one base
    before the next segment's start (the last one: on the last row's start)."""
    out = dict(job=[], estimate_id=[], start=[], end=[], fused_std=[], pseudoweight=[])
    key = t["job"].astype(np.int64) * 100 + t["estimate_id"]
    for k in np.unique(key):
        rows = np.nonzero(key == k)[0]
        starts = t["start"][rows].astype(np.int64)
        pick = np.unique(np.concatenate([[0], np.nonzero(rng.random(rows.size) < 0.25)[0]]))
        pick = pick[np.concatenate([[True], np.diff(starts[pick]) > 0])]       # distinct starts
        for a, b in zip(pick, np.append(pick[1:], -1)):
            out["job"].append(int(t["job"][rows[0]]))
            out["estimate_id"].append(int(t["estimate_id"][rows[0]]))
            out["start"].append(starts[a])
            out["end"].append(starts[b] - 1 if b >= 0 else starts[-1])
            out["fused_std"].append(float(rng.random()))
            out["pseudoweight"].append(float(rng.random() * 100))
    return {k: np.asarray(v) for k, v in out.items()}


def check_split_against_oracle(merged, pieces, t, category, val, std_call, split_pmds=True, pmd_valley_min_width=5000.0):
    """merged / pieces: device (or emulated) tables over all jobs; val: the valley columns they were built from."""
    from oracle import pmd_split as PS
    n_pieces = 0
    for j in np.unique(t["job"]):
        m = t["job"] == j
        rows = np.nonzero(m)[0]
        g = val["job"] == j
        v0 = int(np.nonzero(g)[0][0]) if g.any() else 0
        vj = dict(valleys={k: np.asarray(val[k][g], np.float64 if k not in ("condition", "num_cpgs") else np.int64)
                           for k in ("condition", "start", "end")}, row_in_valley=val["row_in_valley"][m] != 0)
        M = PS.merged_table({k: v[m] for k, v in t.items()}, np.asarray(category)[m], vj, pmd_valley_min_width)
        mm = merged["job"] == j
        assert mm.sum() == M["condition"].size, j
        for k in ("condition", "start", "end", "category", "segment_id"):
            assert np.array_equal(np.asarray(merged[k][mm], np.float64), np.asarray(M[k], np.float64)), (j, k)
        assert np.array_equal(merged["follows_large_gap"][mm] != 0, M["follows_large_gap"]), j
        assert np.array_equal(merged["src_row"][mm], np.where(M["src_row"] >= 0, M["src_row"] + rows[0], -1)), j
        assert np.array_equal(merged["src_valley"][mm], np.where(M["src_valley"] >= 0, M["src_valley"] + v0, -1)), j
        x = std_call["job"] == j
        want = PS.split_call(PS.regions(M, split_pmds), {k: v[x] for k, v in std_call.items()})
        pp = pieces["job"] == j
        assert pp.sum() == want["start"].size, (j, int(pp.sum()), want["start"].size)
        for k in ("estimate_id", "segment_id", "start", "end", "fused_std", "pseudoweight"):
            assert np.array_equal(np.asarray(pieces[k][pp], np.float64), np.asarray(want[k], np.float64)), (j, k)
        n_pieces += want["start"].size
    return n_pieces


def split_trial(emu, rng, t, trial):
    """One emulated run of R/call.R:131-168 on table t against the restatement; returns (pieces, valley rows)."""
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = t["start"].size
    category = np.zeros(n, np.int8)
    category[(t["beta"] < 0.1) & (rng.random(n) < 0.4)] = 2
    category[(t["beta"] >= 0.1) & (rng.random(n) < 0.1)] = 1
    kw = [{}, dict(pmd_valley_min_width=2500.0), dict(valley_max_beta=0.05, pmd_valley_min_width=1500.0)][trial % 3]
    pvw = kw.get("pmd_valley_min_width", 5000.0)
    val = emu(t, category, **kw)
    x = random_std_call(rng, t)
    cols = [np.ascontiguousarray(t[k], dt) for k, dt in (("job", np.int32), ("estimate_id", np.int32), ("start", np.uint32), ("end", np.uint32))]
    cols += [np.ascontiguousarray(category, np.int8), np.ascontiguousarray(val["row_in_valley"], np.int8)]
    nv = val["start"].size
    vcols = [np.ascontiguousarray(val[k], dt) for k, dt in (("job", np.int32), ("condition", np.int32), ("start", np.uint32),
                                                             ("end", np.uint32), ("row0", np.int32))]
    nx = x["start"].size
    xcols = [np.ascontiguousarray(x[k], dt) for k, dt in (("job", np.int32), ("estimate_id", np.int32), ("start", np.uint32),
                                                           ("end", np.uint32), ("fused_std", np.float64), ("pseudoweight", np.float64))]
    split_pmds = trial % 4 != 1
    merged = {k: np.empty(n + nv + 1, dt) for k, dt in MCOLS}
    cap = nx + 2 * (n + nv) + 8
    pieces = {k: np.empty(cap, dt) for k, dt in PCOLS}
    npc = C.c_int64(0)
    nm = emu.lib.emu_pmd_split(C.c_int64(n), *[p(a) for a in cols], C.c_int32(nv), *[p(a) for a in vcols], C.c_int32(nx),
                               *[p(a) for a in xcols], C.c_double(pvw), C.c_int(1 if split_pmds else 0),
                               *[p(merged[k]) for k in ("job", "condition", "segment_id", "start", "end", "category", "src_row",
                                                        "src_valley", "follows_large_gap")],
                               C.c_int64(cap), C.byref(npc), *[p(pieces[k]) for k, _ in PCOLS])
    assert npc.value >= 0
    merged = {k: v[:nm] for k, v in merged.items()}
    pieces = {k: v[:npc.value] for k, v in pieces.items()}
    return (check_split_against_oracle(merged, pieces, t, category, val, x, split_pmds, pvw), int((merged["src_valley"] >= 0).sum()))


def test_emulated_pmd_split_matches_oracle(emu, oracle):
    rng = np.random.default_rng(11)
    total = n_valley_rows = 0
    for trial in range(90):
        a, b = split_trial(emu, rng, random_table(rng, adjacent=trial % 5 == 4), trial)
        total += a
        n_valley_rows += b
    assert total > 2000 and n_valley_rows > 100


@pytest.mark.gpu
@pytest.mark.parametrize("reps", [(2,), (2, 2)])
def test_gpu_pmd_split_matches_oracle(engine, oracle, reps):
    from methylasso_b200._native import MlParams
    tables = [synth.chromosome_table(n, 800 + i, reps) for i, n in enumerate((60_000, 9_000, 25_000))]
    ptr, cols = synth.concat_jobs(tables)
    res = engine.detect_batch(ptr, cols, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        call = res.call_table(0.01, 500.0)
        category = res.call_lmr_umr(0.01, 500.0)["category"]
        counts = []
        for kw, split_pmds, lam in (({}, True, 1000.0), (dict(pmd_valley_min_width=1200.0), True, 100.0),
                                    (dict(pmd_valley_min_width=1200.0), False, 100.0)):
            val = res.call_valleys(0.01, 500.0, **kw)
            x = res.call_pmd_segments(0.01, 500.0, lam)
            merged, pieces = res.call_pmd_split(0.01, 500.0, lam, split_pmds, **kw)
            counts.append(check_split_against_oracle(merged, pieces, call, category, val, x, split_pmds,
                                                     kw.get("pmd_valley_min_width", 5000.0)))
            assert (merged["src_valley"] >= 0).sum() == val["start"].size
        assert min(counts) > 50, counts
    finally:
        res.free()


# ---- R/call.R:169-199: per-position gather along split_call, PMD calls, merge, statistics --------------------------------
def test_pmd_calls_oracle_hand_example(oracle):
    from oracle import pmd_calls as PC
    # 60 positions 200 apart with noisy methylation (a PMD-like stretch), then 30 positions of high methylation
    pos = np.concatenate([1000 + 200 * np.arange(60), 13000 + 200 * np.arange(30)]).astype(np.int32)
    nm = np.concatenate([np.tile([2, 9, 5, 8], 15), np.full(30, 9)]).astype(np.int32)
    pooled = dict(estimate_id=np.ones(90, np.int32), pos=pos, Nmeth=nm, coverage=np.full(90, 10, np.int32))
    split = dict(estimate_id=np.array([1, 1, 1]), start=np.array([1000.0, 4001.0, 13001.0]), end=np.array([4000.0, 13000.0, 18800.0]),
                 fused_std=np.array([0.3, 0.2, 0.01]))
    r = PC.pmd_calls(pooled, split)
    # piece 1: positions 1000..4000 (16), 3003 wide: not admissible.  Piece 2: 4000 (start - 1 <= pos: position 4000 counts
    # twice) .. 13000 (46), 9003 wide, noisy: a PMD.  Piece 3: 13000 (again) .. 18800 (30), wide enough but not noisy
    assert r["category"].tolist() == [0, 1, 0]
    assert r["start"].tolist() == [1000.0, 4000.0, 13000.0] and r["end"].tolist() == [4002.0, 13002.0, 18802.0]
    assert r["num_cpgs"].tolist() == [16, 46, 30] and r["width"].tolist() == [3003.0, 9003.0, 5803.0]
    v2 = nm[15:61] / 10.0
    assert r["beta"][1] == pytest.approx(v2.mean(), abs=1e-15) and r["std"][1] == pytest.approx(v2.std(ddof=1), abs=1e-15)
    assert r["std"][2] == 0.0 and r["beta"][2] == 0.9
    # a lower width threshold makes piece 1 admissible too: pieces 1 and 2 are both PMD and merge
    r = PC.pmd_calls(pooled, split, pmd_valley_min_width=2500.0)
    assert r["category"].tolist() == [1, 0] and r["start"].tolist() == [1000.0, 13000.0] and r["end"].tolist() == [13002.0, 18802.0]
    assert r["num_cpgs"].tolist() == [62, 30] and r["fused_std"][0] == 0.25
    v = np.concatenate([nm[:16], nm[15:61]]) / 10.0
    assert r["beta"][0] == pytest.approx(v.mean(), abs=1e-15) and r["std"][0] == pytest.approx(v.std(ddof=1), abs=1e-15)
    # merge.pmds = F keeps them apart
    assert PC.pmd_calls(pooled, split, pmd_valley_min_width=2500.0, merge_pmds=False)["category"].tolist() == [1, 1, 0]
    # a piece without positions has no row; the two "other" rows around it lie more than 5000 apart: not merged
    split["start"][1], split["end"][1] = 4002.0, 4100.0
    r = PC.pmd_calls(pooled, split)
    assert r["start"].tolist() == [1000.0, 13000.0] and r["category"].tolist() == [0, 0]


def close(a, b, tol):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return a.shape == b.shape and np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b)))


@pytest.mark.gpu
@pytest.mark.parametrize("reps", [(2,), (2, 2)])
def test_gpu_pmd_calls_match_oracle(engine, oracle, reps):
    from methylasso_b200._native import MlParams
    from oracle import pmd_calls as PC
    from test_calltable import pooled_from_table
    tables = [synth.chromosome_table(n, 900 + i, reps) for i, n in enumerate((60_000, 9_000, 25_000))]
    ptr, cols = synth.concat_jobs(tables)
    res = engine.detect_batch(ptr, cols, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        n_pmd = []
        for kw in (dict(), dict(pmd_lambda2=100.0, pmd_valley_min_width=1500.0, pmd_std_threshold=0.1),
                   dict(pmd_lambda2=100.0, pmd_valley_min_width=1500.0, pmd_std_threshold=0.1, merge_pmds=False, split_pmds=False)):
            split_kw = {k: v for k, v in kw.items() if k in ("pmd_valley_min_width",)}
            _, pieces = res.call_pmd_split(0.01, 500.0, kw.get("pmd_lambda2", 1000.0), kw.get("split_pmds", True), **split_kw)
            got = res.call_pmds(0.01, 500.0, **kw)
            okw = {k: v for k, v in kw.items() if k in ("pmd_valley_min_width", "pmd_std_threshold", "merge_pmds")}
            total = 0
            for j, t in enumerate(tables):
                pj = pieces["job"] == j
                want = PC.pmd_calls(pooled_from_table(t), {k: pieces[k][pj].astype(np.float64) for k in ("estimate_id", "start", "end", "fused_std")}, **okw)
                g = got["job"] == j
                assert g.sum() == want["start"].size, (j, int(g.sum()), want["start"].size)
                for k in ("condition", "category", "start", "end", "width", "num_cpgs"):
                    assert np.array_equal(got[k][g].astype(np.float64), want[k].astype(np.float64)), (j, k)
                for k in ("fused_std", "beta", "std"):
                    assert close(got[k][g], want[k], 1e-12), (j, k)
                for c in np.unique(got["condition"][g]):       # regions numbered per condition
                    m = g & (got["condition"] == c)
                    assert np.array_equal(got["segment_id"][m], np.arange(1, m.sum() + 1))
                total += int((want["category"] == 1).sum())
            n_pmd.append(total)
        assert n_pmd[1] > 3 and n_pmd[2] > 3, n_pmd
    finally:
        res.free()


# ---- R/call.R:201-215: seg_call joined with the PMD table, PMD.status ------------------------------------------------------
def test_pmd_status_oracle_hand_example(oracle):
    from oracle import pmd_status as ST
    # seven admissible segments of one condition + a valley (adm NA); PMD table: other | PMD | other
    M = dict(condition=np.ones(8, np.int64), start=np.arange(8) * 1000.0 + 1000, end=np.arange(8) * 1000.0 + 1900,
             category=np.array([0, 2, 0, 1, 2, 0, 3, 2]), src_row=np.array([0, 1, 2, 3, 4, 5, -1, 6]))
    pmd = dict(condition=np.ones(3, np.int64), start=np.array([1000.0, 3500.0, 7000.0]), end=np.array([3400.0, 6950.0, 8900.0]),
               category=np.array([0, 1, 0]))
    r = ST.annotate(M, np.ones(7, np.int64), pmd)
    # row 2 ([3000, 3900]) overlaps "other" and "PMD": two rows; the others one each
    assert r["merged"].tolist() == [0, 1, 2, 2, 3, 4, 5, 6, 7]
    # the LMR (row 3) lies in the PMD: category dropped; UMR row 1 sits between other / other: other; UMR row 4 between the
    # ex-LMR (PMD) and row 5 (PMD): PMD; the valley is not admissible (NA): it is neither updated nor seen as a neighbour,
    # so the last UMR (row 7) has no successor: NA
    assert r["category"].tolist() == [0, 2, 0, 0, 0, 2, 0, 3, 2]
    assert r["status"].tolist() == [0, 0, 0, 1, 1, 1, 1, 0, -1]
    assert r["adm"].tolist() == [1, 1, 1, 1, 1, 1, 1, -1, 1]
    # neighbours that differ: "other"
    pmd2 = dict(condition=np.ones(2, np.int64), start=np.array([1000.0, 2950.0]), end=np.array([1900.0, 8900.0]), category=np.array([0, 1]))
    r = ST.annotate(M, np.ones(7, np.int64), pmd2)
    assert r["status"][1] == 0 and r["merged"].tolist() == list(range(8))       # row 1: no region -> NA, then other (0 vs 1)
    # an inadmissible neighbour is skipped: row 1's neighbours become row 0 and row 3
    adm = np.ones(7, np.int64); adm[2] = 0
    r = ST.annotate(M, adm, pmd)
    assert r["status"][1] == 0 and r["status"][5] == 1


def test_emulated_pmd_status_matches_oracle(emu, oracle):
    from oracle import pmd_status as ST
    rng = np.random.default_rng(5)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    two = na = changed = 0
    for trial in range(150):
        # merged table: segments of 1-3 conditions of 2 jobs, widths / gaps small enough to touch now and then
        rows = []
        for j in range(2):
            for c in range(1, int(rng.integers(2, 4))):
                pos = 1000
                for _ in range(int(rng.integers(1, 60))):
                    w = int(rng.integers(1, 800))
                    rows.append((j, c, pos, pos + w))
                    pos += w + int(rng.integers(0, 3) if rng.random() < 0.5 else rng.integers(3, 3000))
        nm = len(rows)
        job, cond, start, end = (np.asarray([r[k] for r in rows]) for k in range(4))
        cat = rng.choice([0, 1, 2, 3], nm, p=[0.45, 0.2, 0.25, 0.1]).astype(np.int8)
        src_row = np.where(cat == 3, -1, np.cumsum(cat != 3) - 1).astype(np.int32)
        row_adm = (rng.random(int((cat != 3).sum()) + 1) < 0.8).astype(np.int8)
        # PMD table: consecutive regions per (job, condition) covering most of the range, some holes
        g = []
        for j in np.unique(job):
            for c in np.unique(cond[job == j]):
                m = (job == j) & (cond == c)
                lo, hi = int(start[m].min()), int(end[m].max())
                cuts = np.unique(np.concatenate([[lo], rng.integers(lo, hi + 1, int(rng.integers(0, 8))), [hi + 1]]))
                for a, b in zip(cuts[:-1], cuts[1:]):
                    if rng.random() < 0.85:
                        g.append((j, c, a + int(rng.integers(0, 2)), b - 1 + int(rng.integers(0, 3)), int(rng.random() < 0.5)))
        gj, gc, gs, ge, gcat = (np.asarray([r[k] for r in g] or [0]) [:len(g)] for k in range(5))
        f = dict(merged=np.empty(2 * nm, np.int32), status=np.empty(2 * nm, np.int8), category=np.empty(2 * nm, np.int8),
                 adm=np.empty(2 * nm, np.int8))
        cols = [np.ascontiguousarray(a, dt) for a, dt in ((job, np.int32), (cond, np.int32), (start, np.uint32), (end, np.uint32),
                                                          (cat, np.int8), (src_row, np.int32), (row_adm, np.int8))]
        gcols = [np.ascontiguousarray(a, dt) for a, dt in ((gj, np.int32), (gc, np.int32), (gs, np.uint32), (ge, np.uint32), (gcat, np.int8))]
        nf = emu.lib.emu_pmd_status(C.c_int32(nm), *[p(a) for a in cols], C.c_int32(len(g)), *[p(a) for a in gcols],
                                    *[p(f[k]) for k in ("merged", "status", "category", "adm")])
        at = 0
        for j in np.unique(job):
            m = job == j
            base = int(np.nonzero(m)[0][0])
            gm = gj == j if len(g) else np.zeros(0, bool)
            want = ST.annotate(dict(condition=cond[m], start=start[m].astype(float), end=end[m].astype(float), category=cat[m],
                                    src_row=src_row[m]), row_adm,
                               dict(condition=gc[gm], start=gs[gm].astype(float), end=ge[gm].astype(float), category=gcat[gm]))
            k = want["merged"].size
            assert np.array_equal(f["merged"][at:at + k], want["merged"] + base), (trial, j)
            for key in ("status", "category", "adm"):
                assert np.array_equal(f[key][at:at + k], want[key]), (trial, j, key)
            two += int(k - m.sum())
            na += int((want["status"] < 0).sum())
            at += k
        assert at == nf
    assert two > 50 and na > 50


@pytest.mark.gpu
@pytest.mark.parametrize("reps", [(2,), (2, 2)])
def test_gpu_pmd_status_matches_oracle(engine, oracle, reps):
    from methylasso_b200._native import MlParams
    from oracle import pmd_status as ST
    tables = [synth.chromosome_table(n, 950 + i, reps) for i, n in enumerate((60_000, 9_000, 25_000))]
    ptr, cols = synth.concat_jobs(tables)
    res = engine.detect_batch(ptr, cols, MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0))
    try:
        adm = res.call_lmr_umr(0.01, 500.0)["is_admissible"]
        n_pmd_rows = []
        for kw in (dict(), dict(pmd_lambda2=100.0, pmd_valley_min_width=1500.0, pmd_std_threshold=0.1)):
            split_kw = {k: v for k, v in kw.items() if k == "pmd_valley_min_width"}
            merged, _ = res.call_pmd_split(0.01, 500.0, kw.get("pmd_lambda2", 1000.0), True, **split_kw)
            pmd = res.call_pmds(0.01, 500.0, **kw)
            got = res.call_pmd_status(0.01, 500.0, **kw)
            at = total = 0
            for j in range(len(tables)):
                m = merged["job"] == j
                base = int(np.nonzero(m)[0][0])
                g = pmd["job"] == j
                M = {k: merged[k][m].astype(np.int64) for k in ("condition", "category", "src_row")}
                M["start"], M["end"] = merged["start"][m].astype(float), merged["end"][m].astype(float)
                want = ST.annotate(M, adm, dict(condition=pmd["condition"][g], start=pmd["start"][g].astype(float),
                                                end=pmd["end"][g].astype(float), category=pmd["category"][g]))
                k = want["merged"].size
                assert np.array_equal(got["merged_row"][at:at + k], want["merged"] + base), j
                assert np.array_equal(got["pmd_status"][at:at + k], want["status"]), j
                assert np.array_equal(got["category"][at:at + k], want["category"]), j
                for key in ("job", "condition", "start", "end", "src_row", "src_valley"):
                    assert np.array_equal(got[key][at:at + k], merged[key][got["merged_row"][at:at + k]]), (j, key)
                total += int((want["status"] == 1).sum())
                at += k
            assert at == got["job"].size
            n_pmd_rows.append(total)
        assert n_pmd_rows[1] > 20, n_pmd_rows
    finally:
        res.free()


# ---- the whole rule engine, from raw rows: api.segment_methylation against the chain of restatements ----------------------
@pytest.mark.gpu
def test_gpu_segment_methylation_end_to_end(engine, oracle):
    from methylasso_b200 import api
    from oracle import rule_engine as RE
    from test_calltable import pooled_from_table
    tables = [synth.chromosome_table(n, 970 + i, (2,)) for i, n in enumerate((50_000, 20_000))]
    data = {k: np.concatenate([t[k] for t in tables]) for k in tables[0]}
    data["chr"] = np.concatenate([np.full(t["pos"].size, f"chr{i + 1}", dtype=object) for i, t in enumerate(tables)])
    for kw in (dict(), dict(pmd_lambda2=100, pmd_valley_min_width=1500, pmd_std_threshold=0.1)):
        ret = api.signal_detection_fast(data, lambda2=25, engine=engine, keep_resident=True, verbose=False)
        try:
            got = api.segment_methylation(data, ret, **kw)
        finally:
            ret["_resident"].free()
        n_rows = n_pmd = 0
        for i, t in enumerate(tables):
            o = oracle.detect(t, 0, lambda2=25.0)
            want = RE.segment_methylation_singlechr(pooled_from_table(t), o["estimates"], **kw)
            g = got["lmr_umr_valley"]
            m = g["chr"] == f"chr{i + 1}"
            w = want["lmr_umr_valley"]
            assert m.sum() == w["start"].size, (i, int(m.sum()), w["start"].size)
            for k in ("condition", "start", "end", "segment_id", "width", "coverage", "num_cpgs"):
                assert np.array_equal(np.asarray(g[k][m], np.float64), np.asarray(w[k], np.float64)), (i, k)
            assert np.array_equal(g["category"][m], np.array([None, "LMR", "UMR", "UMR-DMV"], dtype=object)[w["category"]]), i
            assert np.array_equal(g["PMD_status"][m], np.array([None, "other", "PMD"], dtype=object)[w["pmd_status"] + 1]), i
            for k in ("beta", "pseudoweight", "std"):
                a, b = np.asarray(g[k][m], float), np.asarray(w[k], float)
                assert np.array_equal(np.isnan(a), np.isnan(b)) and close(a[~np.isnan(a)], b[~np.isnan(b)], 1e-12), (i, k)
            p = got["pmd"]
            pm = p["chr"] == f"chr{i + 1}"
            wp = want["pmd"]
            assert pm.sum() == wp["start"].size, (i, int(pm.sum()), wp["start"].size)
            for k in ("condition", "start", "end", "width", "num_cpgs"):
                assert np.array_equal(np.asarray(p[k][pm], np.float64), np.asarray(wp[k], np.float64)), (i, k)
            for k in ("beta", "std", "fused_std"):
                assert close(p[k][pm], wp[k], 1e-12), (i, k)
            n_rows += int(m.sum())
            n_pmd += int(pm.sum())
        assert n_rows > 50
    assert n_pmd > 0


@pytest.mark.gpu
def test_gpu_segment_methylation_on_a_difference_fit(engine, oracle):
    """R/call.R:40-44: given the result of difference_detection_fast, segment_methylation "will return the segmentation of
    the reference dataset" - estimate_id 1 of the estimates (which the combine of R/fast_diff.R leaves untouched) and the
    rows of the reference condition.  Two conditions with two replicates each; condition 2 covers positions condition 1
    does not (zero pseudoweights in the reference condition's series)."""
    from methylasso_b200 import api
    from oracle import rule_engine as RE
    from test_calltable import pooled_from_table
    tables = []
    for i, n in enumerate((40_000, 15_000)):
        t = synth.chromosome_table(n, 1170 + i, (2, 2))
        upos = np.unique(t["pos"])
        gone = upos[np.random.default_rng(i).random(upos.size) < 0.04]          # positions only condition 2 covers
        keep = ~((t["condition"] == 1) & np.isin(t["pos"], gone))
        tables.append({k: v[keep] for k, v in t.items()})
    data = {k: np.concatenate([t[k] for t in tables]) for k in tables[0]}
    data["chr"] = np.concatenate([np.full(t["pos"].size, f"chr{i + 1}", dtype=object) for i, t in enumerate(tables)])
    ret = api.difference_detection_fast(data, 1, lambda2=25, engine=engine, keep_resident=True, verbose=False)
    try:
        assert ret["detection.type"] == "difference"
        got = api.segment_methylation(data, ret)
        dmr = api.call_differences(data, ret)            # the DMR caller still works on the same resident fit
        assert dmr is not None
    finally:
        ret["_resident"].free()
    n_rows = 0
    for i, t in enumerate(tables):
        o = oracle.detect(t, 0, lambda2=25.0)
        est = o["estimates"]
        ref = est["estimate_id"] == 1
        pooled = pooled_from_table(t)
        pm1 = pooled["estimate_id"] == 1
        want = RE.segment_methylation_singlechr({k: v[pm1] for k, v in pooled.items()}, {k: v[ref] for k, v in est.items()})
        g, w = got["lmr_umr_valley"], want["lmr_umr_valley"]
        m = g["chr"] == f"chr{i + 1}"
        assert m.sum() == w["start"].size, (i, int(m.sum()), w["start"].size)
        assert np.all(g["condition"][m] == 1)
        for k in ("start", "end", "segment_id", "width", "coverage", "num_cpgs"):
            assert np.array_equal(np.asarray(g[k][m], np.float64), np.asarray(w[k], np.float64)), (i, k)
        assert np.array_equal(g["category"][m], np.array([None, "LMR", "UMR", "UMR-DMV"], dtype=object)[w["category"]]), i
        assert np.array_equal(g["PMD_status"][m], np.array([None, "other", "PMD"], dtype=object)[w["pmd_status"] + 1]), i
        for k in ("beta", "std"):
            a, b = np.asarray(g[k][m], float), np.asarray(w[k], float)
            assert np.array_equal(np.isnan(a), np.isnan(b)) and close(a[~np.isnan(a)], b[~np.isnan(b)], 1e-12), (i, k)
        p, wp = got["pmd"], want["pmd"]
        pm = p["chr"] == f"chr{i + 1}"
        assert pm.sum() == wp["start"].size
        for k in ("start", "end", "width", "num_cpgs"):
            assert np.array_equal(np.asarray(p[k][pm], np.float64), np.asarray(wp[k], np.float64)), (i, k)
        n_rows += int(m.sum())
    assert n_rows > 30


@pytest.mark.gpu
def test_gpu_config1_rule_engine(engine, oracle):
    """configs[0] of BASELINE.json end to end: one chr22-sized sample (0.4 M CpGs, one condition), lambda2 = 25, the fit and
    the whole rule engine against the chain of restatements (all rows, drop = F)."""
    from methylasso_b200 import api
    from oracle import rule_engine as RE
    from test_calltable import pooled_from_table
    t = synth.chromosome_table(400_000, 20241, (1,))
    data = dict(t, chr=np.full(t["pos"].size, "chr22", dtype=object))
    ret = api.signal_detection_fast(data, lambda2=25, engine=engine, keep_resident=True, verbose=False)
    try:
        got = api.segment_methylation(data, ret, drop=False)
    finally:
        ret["_resident"].free()
    o = oracle.detect(t, 0, lambda2=25.0)
    want = RE.segment_methylation_singlechr(pooled_from_table(t), o["estimates"])["all"]
    g, w = got["lmr_umr_valley"], want["lmr_umr_valley"]
    assert g["start"].size == w["start"].size
    for k in ("condition", "start", "end", "segment_id", "width", "coverage", "num_cpgs"):
        assert np.array_equal(np.asarray(g[k], np.float64), np.asarray(w[k], np.float64)), k
    assert np.array_equal(g["category"], np.array([None, "LMR", "UMR", "UMR-DMV"], dtype=object)[w["category"]])
    assert np.array_equal(g["PMD_status"], np.array([None, "other", "PMD"], dtype=object)[w["pmd_status"] + 1])
    for k in ("beta", "pseudoweight", "std"):
        assert close(np.nan_to_num(np.asarray(g[k], float), nan=-1.0), np.nan_to_num(np.asarray(w[k], float), nan=-1.0), 1e-12), k
    p, wp = got["pmd"], want["pmd"]
    assert p["start"].size == wp["start"].size
    for k in ("condition", "start", "end", "width", "num_cpgs"):
        assert np.array_equal(np.asarray(p[k], np.float64), np.asarray(wp[k], np.float64)), k
    assert np.array_equal(p["category"] == "PMD", wp["category"] == 1)
    for k in ("beta", "std", "fused_std"):
        assert close(p[k], wp[k], 1e-12), k
    assert (w["category"] == 1).sum() > 50 and (w["category"] == 2).sum() > 200 and (wp["category"] == 1).sum() > 5
