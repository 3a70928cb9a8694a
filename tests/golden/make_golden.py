"""Generate the committed golden fixtures (tests/golden/*.npz).

Run in the build container, where /root/reference is mounted:   python tests/golden/make_golden.py
The fixtures are produced by the oracle with its DP calls routed through the reference's OWN
tf_dp_weight (oracle/_ref/libtfdp_ref.so, compiled verbatim from
/root/reference/src/core/gfl/gfl_tf.c), so they pin (a) the reference DP kernel exactly and
(b) the line-by-line restatement of everything above it.  The reference ships no fixtures of its
own.

reference_fixtures.npz goes one step further: it is produced by the reference's OWN C++ path above
the DP as well (signal_detection_fast_R / signal_detection_R / difference_detection_R /
segment_methylation_helper / weighted_1d_flsa: the unmodified sources compiled in place against the
stand-in Rcpp / Eigen headers of oracle/stubs into oracle/_ref/libmethylasso_ref.so, see
oracle/Makefile target ref_full), without any of the restatement in the loop.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from methylasso_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    O.build(force=True)
    assert O.use_reference_dp(True), "oracle/_ref missing: run where /root/reference is mounted"
    rng = np.random.default_rng(7)
    # 1. raw DP vectors (reference kernel): random, zero weights, monotone, plateaus
    dp = {}
    cases = {
        "random": (rng.normal(size=257), rng.exponential(size=257), 0.7),
        "zero_w": (rng.random(300), np.where(rng.random(300) < 0.6, 0.0, rng.integers(1, 40, 300)).astype(float), 5.0),
        "monotone": (np.linspace(0, 1, 200) ** 2, np.ones(200), 0.05),
        "plateaus": (np.repeat(rng.random(12), 25), rng.integers(5, 60, 300).astype(float), 25.0),
        "known": (np.array([0., 0., 1., 1.]), np.ones(4), 0.5),
    }
    for k, (y, w, lam) in cases.items():
        beta, tm, tp = O.tf_dp_weight(y, w, lam, "reference")
        dp[f"{k}_y"], dp[f"{k}_w"], dp[f"{k}_lam"], dp[f"{k}_beta"] = y, w, np.array(lam), beta
    np.savez_compressed(os.path.join(HERE, "dp_vectors.npz"), **dp)

    # 2. whole-path fixtures on small seeded synthetic chromosomes
    def pack(prefix, t, r, out):
        for c in ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage"):
            out[f"{prefix}_in_{c}"] = t[c]
        out[f"{prefix}_mu"] = r["mu"]
        out[f"{prefix}_offset"] = r["offset"]
        out[f"{prefix}_meta"] = np.array([r["n_params"], r["n_est"], r["n_dsets"], int(r["converged"]),
                                          r["irls_steps"], r["dampings"], r["outer_iters"]])
        for c in ("estimate_id", "start", "beta", "betahat", "pseudoweight"):
            out[f"{prefix}_est_{c}"] = r["estimates"][c]
        seg = O.segment_methylation_helper(r["estimates"], 0.01)
        for c, v in seg.items():
            out[f"{prefix}_seg_{c}"] = v

    g = {}
    t = synth.chromosome_table(3000, 11, (1,))
    pack("fast_1x1", t, O.detect(t, O.MODEL_FAST, lambda2=25.0), g)
    t = synth.chromosome_table(2500, 12, (3,))
    pack("fast_1x3", t, O.detect(t, O.MODEL_FAST, lambda2=25.0), g)
    t = synth.chromosome_table(2500, 13, (2, 2))
    r = O.detect(t, O.MODEL_FAST, lambda2=25.0)
    pack("fast_2x2", t, r, g)
    d = O.fast_diff_combine(r["estimates"], r["n_params"], r["n_est"])
    for c in ("beta", "betahat", "pseudoweight"):
        g[f"fast_2x2_diff_{c}"] = d[c]
    t = synth.chromosome_table(2000, 14, (1,))
    pack("fast_iter3", t, O.detect(t, O.MODEL_FAST, lambda2=10.0, max_iter=3), g)
    t = synth.chromosome_table(1500, 15, (2,))
    pack("full_sig", t, O.detect(t, O.MODEL_FULL_SIGNAL, lambda2=25.0, max_iter=30, nouter=20, hyper=100.0,
                                 bf_per_kb=50.0, max_lambda2=1000.0), g)
    t = synth.chromosome_table(1500, 16, (2, 2))
    pack("full_diff", t, O.detect(t, O.MODEL_FULL_DIFFERENCE, lambda2=25.0, max_iter=30, nouter=20, hyper=100.0,
                                  bf_per_kb=50.0, max_lambda2=1000.0, ref=1), g)
    np.savez_compressed(os.path.join(HERE, "path_fixtures.npz"), **g)

    # 3. the same inputs (and two more designs) through the reference's own C++ path
    assert O.ref_full_lib() is not None, "oracle/_ref/libmethylasso_ref.so missing"
    ref_cases = {
        "fast_1x1": (synth.chromosome_table(3000, 11, (1,)), dict(model=0, lambda2=25.0)),
        "fast_1x3": (synth.chromosome_table(2500, 12, (3,)), dict(model=0, lambda2=25.0)),
        "fast_2x2": (synth.chromosome_table(2500, 13, (2, 2)), dict(model=0, lambda2=25.0)),
        "fast_iter3": (synth.chromosome_table(2000, 14, (1,)), dict(model=0, lambda2=10.0, max_iter=3)),
        "fast_lam1000": (synth.chromosome_table(4000, 17, (2,)), dict(model=0, lambda2=1000.0)),
        "full_sig": (synth.chromosome_table(1500, 15, (2,)), dict(model=1, lambda2=25.0, max_iter=30, nouter=20)),
        "full_sig_3c": (synth.chromosome_table(1200, 18, (1, 1, 1)), dict(model=1, lambda2=50.0, max_iter=30, nouter=20)),
        "full_diff": (synth.chromosome_table(1500, 16, (2, 2)), dict(model=2, lambda2=25.0, max_iter=30, nouter=20, ref=1)),
        "full_diff_ref2": (synth.chromosome_table(1200, 19, (1, 2)), dict(model=2, lambda2=50.0, max_iter=30, nouter=20, ref=2)),
    }
    h = {}
    for name, (t, kw) in ref_cases.items():
        r = O.reference_detect(t, **kw)
        for c in ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage"):
            h[f"{name}_in_{c}"] = t[c]
        h[f"{name}_kw"] = np.array([kw["model"], kw["lambda2"], kw.get("max_iter", 1), kw.get("nouter", 20), kw.get("ref", 1)], dtype=float)
        h[f"{name}_mu"], h[f"{name}_offset"], h[f"{name}_lambda2"] = r["mu"], r["offset"], r["lambda2"]
        h[f"{name}_converged"] = np.array(int(r["converged"]))
        for c, v in r["estimates"].items():
            h[f"{name}_est_{c}"] = v
        seg = O.reference_segment_helper(r["estimates"], 0.01)
        for c, v in seg.items():
            h[f"{name}_seg_{c}"] = v
    y, w = rng.random(500), rng.integers(0, 30, 500).astype(float)
    for lam2, lam1 in ((25.0, 0.0), (3.0, 0.05)):
        h[f"flsa_{lam2:g}_{lam1:g}"] = O.reference_weighted_1d_flsa(y, w, lam2, lam1)
    h["flsa_y"], h["flsa_w"] = y, w
    np.savez_compressed(os.path.join(HERE, "reference_fixtures.npz"), **h)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
