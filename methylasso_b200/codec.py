"""Host side of the input codec: what MethyLasso.R's process_replicate (MethyLasso.R:246-297) and the coverage filter
(:361) do with one replicate file, with the text parsed on the GPU (csrc/codec.cu) instead of data.table::fread.

    table = read_replicate(engine, text, cov=5, meth=4)            # coverage / methylation-level columns (1-based)
    table = read_replicate(engine, text, mC=5, uC=6)               # methylated / unmethylated counts

The column numbers are the CLI's (1-based, as given to --cov / --meth or --mC / --uC).  Returns chr (names per
row), pos, coverage, Nmeth, beta - the columns process_replicate builds - after `coverage >= min_coverage`."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api


def load_text(path: str) -> bytes:
    """The bytes of a replicate file; a gzip file (magic 1f 8b, whatever its name) is inflated on the host, as fread does
    through R.utils (MethyLasso.R:246-262 reads `.gz` files the same way) - the parser itself takes plain text."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] == b"\x1f\x8b":
        import gzip
        return gzip.decompress(raw)
    return raw


def read_replicate_file(engine, path: str, **kw):
    """read_replicate on a file path (plain or gzip)."""
    return read_replicate(engine, load_text(path), **kw)


def _is_number(tok: bytes) -> bool:
    try:
        float(tok)
        return True
    except ValueError:
        return False


def has_header(text: bytes, sep=None) -> bool:
    """MethyLasso.R:248-251: the file has a header iff a field of its first line is not numeric.  (With chromosome
    names like 'chr1' that is always the case, so the reference always drops the first line of such files: mirrored,
    not corrected.)"""
    first = text.split(b"\n", 1)[0].rstrip(b"\r")
    return any(not _is_number(t) for t in (first.split(sep) if sep else first.split()))


def parse_text(engine, text: bytes, col_a: int, col_b: int, skip_lines: int = 0, sep: bytes | None = None):
    """Columns 1, 2, col_a + 1, col_b + 1 of every data line (col_a / col_b 0-based): dict with chr (object array of
    names), pos, a, b.  Fields the GPU could not convert in one rounded operation are parsed here with float() -
    the same correctly rounded value; blank lines are dropped; a malformed line raises."""
    lib, h = engine.lib, C.c_void_p()
    n, runs = C.c_int64(0), C.c_int64(0)
    engine._check(lib.ml_text_parse(engine.h, text, len(text), int(skip_lines), int(col_a), int(col_b),
                                    sep if sep else b"\x00", C.byref(h), C.byref(n), C.byref(runs)))
    try:
        n, runs = n.value, runs.value
        pos, a, b = np.empty(max(n, 1), np.int32), np.empty(max(n, 1)), np.empty(max(n, 1))
        status, start = np.empty(max(n, 1), np.int8), np.empty(max(n, 1), np.int64)
        rf, ro, rl = np.empty(max(runs, 1), np.int64), np.empty(max(runs, 1), np.int32), np.empty(max(runs, 1), np.int32)
        p = api._p
        engine._check(lib.ml_text_fetch(engine.h, h, p(pos), p(a), p(b), p(status), p(start), p(rf), p(ro), p(rl)))
    finally:
        lib.ml_text_free(h)
    pos, a, b, status, start = pos[:n], a[:n], b[:n], status[:n], start[:n]
    rf, ro, rl = rf[:runs], ro[:runs], rl[:runs]
    for r in np.nonzero(status == 1)[0]:                       # host parse of what the fast path does not cover
        line = text[start[r]:text.find(b"\n", start[r]) if text.find(b"\n", start[r]) >= 0 else len(text)].rstrip(b"\r")
        f = line.split(sep) if sep else line.split()
        pos[r], a[r], b[r] = int(float(f[1])), float(f[col_a]), float(f[col_b])
    bad = np.nonzero(status == 2)[0]
    if bad.size:
        raise ValueError(f"line {int(bad[0]) + skip_lines + 1} of the file is malformed")
    names = np.array([text[start[rf[k]] + ro[k]: start[rf[k]] + ro[k] + rl[k]].decode() for k in range(runs)], dtype=object)
    chrom = names[np.searchsorted(rf, np.arange(n), side="right") - 1] if n else np.empty(0, dtype=object)
    keep = status != 3
    return dict(chr=chrom[keep], pos=pos[keep], a=a[keep], b=b[keep])


def read_replicate(engine, text: bytes, cov: int | None = None, meth: int | None = None, mC: int | None = None,
                   uC: int | None = None, min_coverage: float = 5, sep: bytes | None = None):
    """process_replicate (MethyLasso.R:246-297) + the coverage filter (:361) for one replicate file."""
    if (cov is None) != (meth is None) or (mC is None) != (uC is None) or (cov is None) == (mC is None):
        raise ValueError("give either cov and meth or mC and uC (1-based column numbers)")
    skip = 1 if has_header(text, sep) else 0
    if cov is not None:
        t = parse_text(engine, text, cov - 1, meth - 1, skip, sep)
        coverage, level = t["a"], t["b"]
        if level.size and level.max() > 1:                      # :266-268 percentages
            nmeth, beta = np.round(coverage * level / 100), level / 100
        else:                                                   # :269-271 fractions
            nmeth, beta = np.round(coverage * level), level
    else:                                                       # :273-276
        t = parse_text(engine, text, mC - 1, uC - 1, skip, sep)
        coverage = np.round(t["a"] + t["b"])
        nmeth = t["a"]
        with np.errstate(divide="ignore", invalid="ignore"):
            beta = nmeth / coverage
    keep = coverage >= min_coverage                             # :361
    return dict(chr=t["chr"][keep], pos=t["pos"][keep], coverage=coverage[keep], Nmeth=nmeth[keep], beta=beta[keep])
