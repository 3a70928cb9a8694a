"""Generate tests/golden/valley_fixtures.npz: the valley (DMV) step of R/call.R:100-133 as restated by oracle/valleys.py,
on the LMR / UMR input table of rlayer_fixtures.npz (two conditions, 40 000 CpGs), for the reference's defaults and for
a narrower pmd_valley_min_width that yields more valleys on a chromosome of this size; and the two tables the whole rule
engine (R/call.R:31-221, oracle/rule_engine.py) returns for that input with drop = F (keys `re_<case>_<table>_<column>`).  R is not in the image: the
fixture pins the RESTATEMENT (regression net, and a run-independent target for the CUDA path), not R itself.

    python tests/golden/make_golden_valleys.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import lmr_umr as LU, oracle as O, rule_engine as RE, valleys as V  # noqa: E402
from test_calltable import pooled_from_table  # noqa: E402

COLS = ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage")
CASES = (("default", {}), ("narrow", dict(pmd_valley_min_width=1200.0)))
RE_CASES = (("default", {}), ("narrow", dict(pmd_lambda2=100.0, pmd_valley_min_width=1200.0, pmd_std_threshold=0.1)))


def compute(g):
    t = {k: g["lu_in_" + k] for k in COLS}
    o = O.detect(t, O.MODEL_FAST, lambda2=25.0)
    call = O.call_table(pooled_from_table(t), O.segment_methylation_helper(o["estimates"], 0.01), 500.0)
    category = LU.annotate(call)["category"]
    out = {}
    for name, kw in CASES:
        r = V.valleys(call, category, **kw)
        for k, v in r["valleys"].items():
            out[f"{name}_{k}"] = v
        out[f"{name}_row_in_valley"] = r["row_in_valley"]
        out[f"{name}_low_id_width"] = r["low"]["low_id_width"]
        out[f"{name}_low_id_beta"] = r["low"]["low_id_beta"]
    pooled = pooled_from_table(t)
    for name, kw in RE_CASES:
        r = RE.segment_methylation_singlechr(pooled, o["estimates"], **kw)["all"]
        for tab in ("lmr_umr_valley", "pmd"):
            for k, v in r[tab].items():
                out[f"re_{name}_{tab}_{k}"] = v
    return out


def main():
    O.build()
    out = compute(np.load(os.path.join(HERE, "rlayer_fixtures.npz")))
    np.savez_compressed(os.path.join(HERE, "valley_fixtures.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
