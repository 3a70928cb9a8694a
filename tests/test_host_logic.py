"""CPU tests of the Python host layer that need no GPU: chromosome splitting, job grouping."""
import os

import numpy as np

from methylasso_b200 import api, synth


def test_split_by_chr_keeps_first_appearance_order_and_rows():
    tables = [synth.chromosome_table(n, 10 + i, (2,)) for i, n in enumerate((300, 120, 500))]
    names = ["chr9", "chr1", "chrX"]   # data[, unique(chr)] keeps first-appearance order, not sorted order
    data = {k: np.concatenate([t[k] for t in tables]) for k in tables[0]}
    data["chr"] = np.concatenate([np.repeat(nm, t["pos"].size) for nm, t in zip(names, tables)])
    # interleave rows of different chromosomes to make the split do real work
    perm = np.random.default_rng(0).permutation(data["pos"].size)
    order = np.argsort(perm, kind="stable")
    shuffled = {k: v[perm] for k, v in data.items()}
    # restore (dataset, pos) order inside each chromosome as the R layer does (data is keyed that way)
    key = np.lexsort((shuffled["pos"], shuffled["dataset"]))
    shuffled = {k: v[key] for k, v in shuffled.items()}
    got_names, ptr, cols, _ = api.split_by_chr(shuffled)
    assert sorted(got_names.tolist()) == sorted(names)
    for nm, j in zip(got_names, range(3)):
        t = tables[names.index(nm)]
        sl = slice(ptr[j], ptr[j + 1])
        for k in ("dataset", "pos", "Nmeth", "coverage"):
            assert np.array_equal(cols[k][sl], t[k]), (nm, k)
    del order


def test_split_jobs_partitions_and_balances():
    tables = [synth.chromosome_table(n, 50 + i, (1,)) for i, n in enumerate((900, 100, 400, 350, 800, 50))]
    ptr, cols = synth.concat_jobs(tables)
    groups = api.split_jobs(ptr, cols, 3)
    seen = sorted(j for jobs, _, _ in groups for j in jobs)
    assert seen == list(range(6))
    sizes = [int(g[1][-1]) for g in groups]
    assert max(sizes) < 1.5 * (sum(sizes) / 3)
    for jobs, gptr, gcols in groups:
        for i, j in enumerate(jobs):
            assert np.array_equal(gcols["pos"][gptr[i]:gptr[i + 1]], tables[j]["pos"])


def test_segment_methylation_refuses_what_it_cannot_do():
    """R/call.R:229-244 mirror: needs a fit kept resident (no GPU needed to be told so)."""
    import pytest
    from methylasso_b200 import api
    with pytest.raises(api.MethyLassoError, match="signal_detection_fast or difference_detection_fast"):
        api.segment_methylation({}, {"detection.type": "something else"})
    for kind in ("signal", "difference"):
        with pytest.raises(api.MethyLassoError, match="keep_resident"):
            api.segment_methylation({}, {"detection.type": kind})


def test_codec_load_text_inflates_gzip(tmp_path):
    import gzip
    from methylasso_b200 import codec
    text = b"chr1\t100\t100\t50\t5\t5\nchr1\t102\t102\t25\t2\t6\n"
    plain, packed = tmp_path / "a.cov", tmp_path / "a.cov.gz"
    plain.write_bytes(text)
    packed.write_bytes(gzip.compress(text))
    assert codec.load_text(str(plain)) == text and codec.load_text(str(packed)) == text


def test_two_fma_quotient_of_counts_is_the_ieee_quotient():
    """counts_div.cuh: for an integer numerator up to 65536 and a coverage below 1024 the quotient formed from the
    correctly rounded reciprocal by two FMAs has the bits of the division - every pair, on the CPU with the C library's fma."""
    import ctypes as C
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    src, lib = os.path.join(here, "emu", "emu_counts_div.cpp"), os.path.join(here, "emu", "libemu_counts_div.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", lib, src])
    L = C.CDLL(lib)
    L.emu_counts_div_mismatches.restype = C.c_longlong
    assert L.emu_counts_div_mismatches(65536, 1024) == 0


def test_pipeline_turns_are_taken_in_index_order_and_an_abort_frees_the_waiters():
    """pipeline._Turns: the groups of an end-to-end call take the packing threads and the bus in the order of their index,
    whatever order their threads arrive in; a failing group does not leave the others waiting for ever."""
    import random
    import threading
    import time
    import pytest
    from methylasso_b200.pipeline import _Turns
    turns, order = _Turns(), []

    def worker(i):
        time.sleep(random.random() * 0.02)
        with turns.turn(i):
            order.append(i)
    ths = [threading.Thread(target=worker, args=(i,)) for i in random.sample(range(8), 8)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(5)
    assert order == list(range(8))
    turns.reset()
    seen = []

    def waiter():
        try:
            with turns.turn(3):
                seen.append("ran")
        except RuntimeError:
            seen.append("released")
    th = threading.Thread(target=waiter)
    th.start()
    time.sleep(0.02)
    turns.abort()
    th.join(5)
    assert seen == ["released"]
    with pytest.raises(RuntimeError):
        with turns.turn(0):
            pass
    turns.reset()
    with turns.turn(0):
        pass
