"""CPU tests of the host half of the input validation: ml_batch_upload reduces dataset / condition / replicate to a
run table and evaluates the reference's predicates on them there (core/Observations.hpp:36-56); what the device is left
with is the order of the positions and the two checks on the counts.  Checked against a line-by-line restatement of
check_observations and against the C oracle (orc_check_observations, itself pinned to the reference's refusals in
tests/test_oracle_pin.py)."""
import ctypes as C

import numpy as np
import pytest

from methylasso_b200 import synth

STRUCTURAL = (3, 4, 5, 7, 8, 9, 10)


@pytest.fixture(scope="module")
def lib():
    from methylasso_b200 import build, _native
    build.build()
    return _native.load()


def scan(lib, t, threads=1):
    n = t["pos"].size
    cols = [np.ascontiguousarray(t[k], np.int32) for k in ("dataset", "condition", "replicate", "pos")]
    cap = max(n, 1)
    first, cond, rep = np.zeros(cap, np.int64), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    nr, code, row = C.c_int64(), C.c_int32(), C.c_int64()
    rc = lib.ml_debug_scan_runs(n, *[c.ctypes.data_as(C.c_void_p) for c in cols], threads, cap, C.byref(nr),
                                first.ctypes.data_as(C.c_void_p), cond.ctypes.data_as(C.c_void_p),
                                rep.ctypes.data_as(C.c_void_p), C.byref(code), C.byref(row))
    assert rc == 0
    k = nr.value
    return dict(first=first[:k], cond=cond[:k], rep=rep[:k], code=code.value, row=row.value)


def walk(t):
    """core/Observations.hpp:36-56, statement by statement: (code, row) of the first structural offence and the row
    of the first position-order offence before it (None: none)."""
    ds, cd, rp, pos = (np.asarray(t[k]).astype(np.uint32 if k != "pos" else np.int64) for k in ("dataset", "condition", "replicate", "pos"))
    if ds[0] != 1:
        return (3, 0), None
    if cd[0] != 1:
        return (4, 0), None
    if rp[0] != 1:
        return (5, 0), None
    last_d, last_c, last_r, last_pos = int(ds[0]), int(cd[0]), int(rp[0]), int(pos[0])
    order = None
    for i in range(1, ds.size):
        if last_d == ds[i]:
            if last_pos >= pos[i] and order is None:
                order = i
        else:
            last_d += 1
            if ds[i] != last_d:
                return (7, i), order
            if cd[i] == last_c:
                last_r += 1
                if rp[i] != last_r:
                    return (8, i), order
            else:
                last_c += 1
                if cd[i] != last_c:
                    return (9, i), order
                last_r = int(rp[i])
                if last_r != 1:
                    return (10, i), order
        last_pos = int(pos[i])
    return (0, -1), order


def corrupt(t, rng, at_head=False):
    t = {k: v.copy() for k, v in t.items()}
    n = t["pos"].size
    kind = rng.integers(0, 9)
    i = 0 if at_head else int(rng.integers(0, n))
    if kind == 0:
        t["dataset"][i:] += 1                      # a gap (or a first id that is not 1)
    elif kind == 1:
        t["dataset"][i] = rng.integers(0, 5)       # a stray code in the middle of a run
    elif kind == 2:
        t["condition"][i:] += rng.integers(1, 3)
    elif kind == 3:
        t["replicate"][i:] += 1
    elif kind == 4:
        t["replicate"][:] = 2
    elif kind == 5:
        t["pos"][i] = t["pos"][max(i - 1, 0)]      # a tie: not strictly increasing
    elif kind == 6:
        j = int(rng.integers(0, n))
        t["dataset"][min(i, j):max(i, j)] = 1
    elif kind == 7:
        t["condition"][i] += 1                     # inside a run: the reference never looks
    return t


def test_clean_tables_give_one_run_per_dataset(lib):
    for reps in ((1,), (3,), (2, 2), (1, 2, 1)):
        t = synth.chromosome_table(3000, 5, reps)
        for threads in (1, 4):
            r = scan(lib, t, threads)
            assert r["code"] == 0 and r["row"] == -1
            d = np.asarray(t["dataset"])
            want = np.flatnonzero(np.r_[True, d[1:] != d[:-1]])
            assert np.array_equal(r["first"], want)
            assert np.array_equal(r["cond"], t["condition"][want]) and np.array_equal(r["rep"], t["replicate"][want])


def test_corrupted_tables_match_the_reference_walk(lib, oracle):
    rng = np.random.default_rng(7)
    seen = set()
    for it in range(400):
        reps = [(1,), (3,), (2, 2), (1, 2, 1), (2, 1)][it % 5]
        t = corrupt(synth.chromosome_table(int(rng.integers(40, 400)), 100 + it, reps), rng, at_head=it % 8 == 0)
        (code, row), order = walk(t)
        r = scan(lib, t, 1 + it % 3)
        assert (r["code"], r["row"]) == (code, row), (it, r["code"], r["row"], code, row)
        # runs reported = those that lie before the offence
        d = np.asarray(t["dataset"])
        heads = np.flatnonzero(np.r_[True, d[1:] != d[:-1]])
        want = heads[heads < row] if code else heads
        assert np.array_equal(r["first"], want), it
        # and the whole verdict agrees with the C oracle: the first offence in row order wins
        o = oracle.check_observations(t)
        if code and (order is None or row < order):
            assert o == code, (it, o, code)
        elif order is not None:
            assert o == 6, (it, o)
        else:
            assert o in (0, 11, 12), (it, o)
        seen.add(o)
    assert {0, 3, 4, 6, 7, 8, 9, 10} <= seen, seen


def test_many_runs_and_empty(lib):
    n = 5000
    t = dict(dataset=np.arange(1, n + 1, dtype=np.int32), condition=np.ones(n, np.int32),
             replicate=np.arange(1, n + 1, dtype=np.int32), pos=np.arange(n, dtype=np.int32))
    r = scan(lib, t, 3)
    assert r["code"] == 0 and r["first"].size == n           # every row its own dataset: legal, if odd
    e = {k: np.zeros(0, np.int32) for k in ("dataset", "condition", "replicate", "pos")}
    r = scan(lib, e)
    assert r["code"] == 0 and r["first"].size == 0
