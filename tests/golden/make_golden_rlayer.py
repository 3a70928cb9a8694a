"""Generate tests/golden/rlayer_fixtures.npz: small fixed inputs and the outputs of the CPU restatements of the R-layer
steps that became native in round 1 (oracle/diffcall.py: R/call.R:254-365; oracle/lmr_umr.py: R/call.R:75-91;
oracle/codec.py: MethyLasso.R:246-297).  R is not in the image, so these fixtures pin the RESTATEMENTS (a regression
net for them and a second, run-independent target for the CUDA path), not R itself.

    python tests/golden/make_golden_rlayer.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from methylasso_b200 import synth  # noqa: E402
from oracle import codec as OC, diffcall as DC, lmr_umr as LU, oracle as O  # noqa: E402
from test_calltable import pooled_from_table  # noqa: E402

COLS = ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage")


def main():
    O.build()
    out = {}
    # DMR caller: 3 vs 2 replicates with missing positions
    rng = np.random.default_rng(11)
    t = synth.chromosome_table(5000, 901, (3, 2))
    keep = rng.random(t["pos"].size) > 0.2
    for d in np.unique(t["dataset"]):
        keep[np.nonzero(t["dataset"] == d)[0][0]] = True
    t = {k: v[keep] for k, v in t.items()}
    for k in COLS:
        out["dmr_in_" + k] = t[k]
    o = O.detect(t, O.MODEL_FAST, lambda2=25.0)
    est = O.fast_diff_combine(o["estimates"], o["n_params"], o["n_est"])
    g = DC.call_differences_singlechr(t, est, 2)
    tab, called = DC.call_differences([g])
    for k, v in tab.items():
        out["dmr_out_" + k] = v
    out["dmr_out_called"] = called
    # LMR / UMR annotation: two conditions
    t = synth.chromosome_table(40_000, 902, (2, 2))
    for k in COLS:
        out["lu_in_" + k] = t[k]
    o = O.detect(t, O.MODEL_FAST, lambda2=25.0)
    call = O.call_table(pooled_from_table(t), O.segment_methylation_helper(o["estimates"], 0.01), 500.0)
    for k, v in LU.annotate(call).items():
        out["lu_out_" + k] = v
    # codec: a Bismark-shaped file
    t = synth.chromosome_table(2000, 903, (1,), min_cov=1)
    rows = [f"chr{1 + (i > 900)}\t{p}\t{p}\t{format(100.0 * int(m) / int(c), '.15g')}\t{m}\t{c - m}"
            for i, (p, m, c) in enumerate(zip(t["pos"], t["Nmeth"], t["coverage"]))]
    text = ("chr\tstart\tend\tpct\tmeth\tunmeth\n" + "\n".join(rows) + "\n").encode()
    out["codec_text"] = np.frombuffer(text, np.uint8)
    for mode, kw in (("counts", dict(mC=5, uC=6)), ("level", dict(cov=5, meth=4))):
        r = OC.read_replicate(text, min_coverage=5, **kw)
        out[f"codec_{mode}_chr"] = r["chr"].astype("U")
        for k in ("pos", "coverage", "Nmeth", "beta"):
            out[f"codec_{mode}_{k}"] = r[k]
    np.savez_compressed(os.path.join(HERE, "rlayer_fixtures.npz"), **out)
    print({k: v.shape for k, v in list(out.items())[:5]}, len(out), "arrays")


if __name__ == "__main__":
    main()
