"""Input codec (MethyLasso.R:246-297, :361): the parsing logic compiled for the CPU against Python's float(), and the
CUDA path against the plain-Python restatement of process_replicate."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from methylasso_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "emu_codec.cpp")
    lib = os.path.join(HERE, "emu", "libemu_codec.so")
    deps = [src] + [os.path.join(HERE, "..", "methylasso_b200", "csrc", f) for f in ("codec_core.cuh", "hd.h")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", lib, src])
    return C.CDLL(lib)


def test_number_parser_is_exact_or_says_so(emu):
    def parse(s):
        v = C.c_double()
        return emu.emu_parse_double(s.encode(), len(s), C.byref(v)), v.value
    rng = np.random.default_rng(0)
    ok = 0
    for _ in range(60_000):
        k = int(rng.integers(0, 5))
        s = (str(int(rng.integers(0, 2**31))), repr(float(rng.random() * 100)), f"{rng.random() * 100:.{rng.integers(0, 14)}f}",
             f"{rng.random() * 10.0 ** float(rng.integers(-20, 20)):.{rng.integers(1, 16)}e}",
             repr(float(rng.integers(0, 1000) / max(1, rng.integers(1, 1000)))))[k]
        st, v = parse(s)
        assert st in (0, 1), s
        if st == 0:
            ok += 1
            assert v == float(s) and np.signbit(v) == np.signbit(float(s)), s
    assert ok > 30_000
    for s in ("", "NA", "nan", "1.2.3", "12a", ".", "-", "1e", "inf"):
        assert parse(s)[0] == 2, s
    assert parse("1e400")[0] == 1 and parse("12345678901234567")[0] == 1         # valid, but the host's job
    assert parse("66.6666666666667") == (0, 66.6666666666667)                    # 15 significant digits: a Bismark percentage


def test_line_parser(emu):
    def line(s, a, b, sep=0):
        pos, x, y, off, ln = C.c_longlong(), C.c_double(), C.c_double(), C.c_int(), C.c_int()
        st = emu.emu_parse_line(s.encode(), len(s), a, b, sep, C.byref(pos), C.byref(x), C.byref(y), C.byref(off), C.byref(ln))
        return st, pos.value, x.value, y.value, s[off.value:off.value + ln.value]
    assert line("chr1\t10468\t10468\t66.6666666666667\t2\t1", 4, 5) == (0, 10468, 2.0, 1.0, "chr1")
    assert line("chr1\t10468\t10468\t66.6666666666667\t2\t1\r", 3, 4) == (0, 10468, 66.6666666666667, 2.0, "chr1")
    assert line("chrX  77   5 0.25", 2, 3) == (0, 77, 5.0, 0.25, "chrX")           # runs of blanks are one separator
    assert line("1,5,3,4", 2, 3, ord(","))[:4] == (0, 5, 3.0, 4.0)
    assert line("chr1\t10\t3", 2, 3)[0] == 2                                      # a selected column is missing
    assert line("chr1\t1.5\t3\t4", 2, 3)[0] == 2                                   # position is not an integer
    assert line("chr1\tx\t3\t4", 2, 3)[0] == 2
    assert line("chr1\t10\t0.12345678901234567\t4", 2, 3)[0] == 1                  # 17 digits: host parse


def bismark_text(t, header=False, crlf=False, chrom="chr7"):
    """one replicate of a synthetic table as a Bismark .cov file: chr, start, end, percent, count_meth, count_unmeth"""
    nl = "\r\n" if crlf else "\n"
    # percentages with 15 significant digits as Bismark prints them; every 1000th line with all 17 (host parse)
    rows = [f"{chrom if i < t['pos'].size // 2 else 'chrUn_KI270742v1'}\t{p}\t{p}\t"
            f"{(repr if i % 1000 == 0 else (lambda x: format(x, '.15g')))(100.0 * int(m) / int(c))}\t{m}\t{c - m}"
            for i, (p, m, c) in enumerate(zip(t["pos"], t["Nmeth"], t["coverage"]))]
    return ((("chr\tstart\tend\tpct\tmeth\tunmeth" + nl) if header else "") + nl.join(rows) + nl).encode()


def one_replicate(n, seed):
    t = synth.chromosome_table(n, seed, (1,), min_cov=1)
    return {k: v for k, v in t.items()}


def test_oracle_reads_what_was_written():
    from oracle import codec as OC
    t = one_replicate(3000, 3)
    r = OC.read_replicate(bismark_text(t, header=True), mC=5, uC=6, min_coverage=5)
    keep = t["coverage"] >= 5
    assert np.array_equal(r["pos"], t["pos"][keep]) and np.array_equal(r["Nmeth"], t["Nmeth"][keep])
    assert np.array_equal(r["coverage"], t["coverage"][keep]) and set(r["chr"]) == {"chr7", "chrUn_KI270742v1"}
    # without a header line the reference's own detection eats the first data row ('chr7' is not numeric)
    r2 = OC.read_replicate(bismark_text(t, header=False), mC=5, uC=6, min_coverage=0)
    assert r2["pos"].size == t["pos"].size - 1


@pytest.mark.gpu
@pytest.mark.parametrize("header,crlf", [(True, False), (False, False), (True, True)])
def test_gpu_codec_matches_oracle(engine, header, crlf):
    from methylasso_b200 import codec
    from oracle import codec as OC
    t = one_replicate(120_000, 17)
    text = bismark_text(t, header=header, crlf=crlf)
    for kw in (dict(mC=5, uC=6), dict(cov=5, meth=4)):           # counts; "coverage" = meth count, level = percentage
        want = OC.read_replicate(text, min_coverage=5, **kw)
        got = codec.read_replicate(engine, text, min_coverage=5, **kw)
        assert want["pos"].size > 10_000
        assert np.array_equal(got["chr"], want["chr"])
        for k in ("pos", "coverage", "Nmeth", "beta"):
            assert np.array_equal(got[k], want[k], equal_nan=True), k


@pytest.mark.gpu
def test_gpu_codec_edge_cases(engine):
    from methylasso_b200 import codec
    from oracle import codec as OC
    # no trailing newline, blank lines, a 17-digit level (host parse), numeric chromosome names (no header detected)
    text = b"1 100 7 0.5\n1 105 9 0.12345678901234567\n\n2 7 30 1\n2 9 4 0.25"
    want = OC.read_replicate(text, cov=3, meth=4, min_coverage=5)
    got = codec.read_replicate(engine, text, cov=3, meth=4, min_coverage=5)
    assert got["pos"].tolist() == want["pos"].tolist() == [100, 105, 7]
    assert np.array_equal(got["chr"], want["chr"]) and np.array_equal(got["beta"], want["beta"])
    assert np.array_equal(got["Nmeth"], want["Nmeth"])             # round(7 * 0.5) = 4: half to even, as R rounds
    with pytest.raises(ValueError):
        codec.read_replicate(engine, b"chr1\t5\t3\n" * 3, cov=3, meth=4)          # a selected column is missing
    assert codec.parse_text(engine, b"", 2, 3)["pos"].size == 0
