"""TEST INFRASTRUCTURE - plain-Python restatement of MethyLasso.R's process_replicate (MethyLasso.R:246-297) and the
coverage filter (:361) for one replicate file: str.split + float() (correctly rounded, as fread's parser is) and R's
round() (IEC 60559 half-to-even, numpy.round).  Only tests/ may import this module."""
import numpy as np


def read_replicate(text: bytes, cov=None, meth=None, mC=None, uC=None, min_coverage=5, sep=None):
    lines = [l.rstrip(b"\r") for l in text.split(b"\n")]
    lines = [l for l in lines if l.strip()]
    split = (lambda l: l.split(sep)) if sep else (lambda l: l.split())

    def num(tok):
        try:
            float(tok)
            return True
        except ValueError:
            return False
    if lines and any(not num(t) for t in split(lines[0])):     # :248-251 header detection
        lines = lines[1:]
    rows = [split(l) for l in lines]
    chrom = np.array([r[0].decode() for r in rows], dtype=object)
    pos = np.array([int(float(r[1])) for r in rows], np.int32)
    if cov is not None:
        coverage = np.array([float(r[cov - 1]) for r in rows])
        level = np.array([float(r[meth - 1]) for r in rows])
        if level.size and level.max() > 1:
            nmeth, beta = np.round(coverage * level / 100), level / 100
        else:
            nmeth, beta = np.round(coverage * level), level
    else:
        a = np.array([float(r[mC - 1]) for r in rows])
        b = np.array([float(r[uC - 1]) for r in rows])
        coverage, nmeth = np.round(a + b), a
        with np.errstate(divide="ignore", invalid="ignore"):
            beta = nmeth / coverage
    keep = coverage >= min_coverage
    return dict(chr=chrom[keep], pos=pos[keep], coverage=coverage[keep], Nmeth=nmeth[keep], beta=beta[keep])
