"""ctypes binding of the TEST-ONLY CPU oracle (oracle/liboracle.so, oracle/_ref/libtfdp_ref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under methylasso_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libtfdp_ref.so")

MODEL_FAST, MODEL_FULL_SIGNAL, MODEL_FULL_DIFFERENCE = 0, 1, 2


class OrcParams(C.Structure):
    _fields_ = [("lambda2", C.c_double), ("max_lambda2", C.c_double), ("lambda1", C.c_double),
                ("tol_val", C.c_double), ("bf_per_kb", C.c_double), ("hyper", C.c_double),
                ("max_iter", C.c_uint32), ("nouter", C.c_uint32), ("ref", C.c_uint32),
                ("model", C.c_int32), ("optimize_lambda2", C.c_int32), ("optimize_hyper", C.c_int32)]


class OrcResult(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_params", C.c_int64), ("n_est", C.c_int32),
                ("n_dsets", C.c_int32), ("converged", C.c_int32), ("irls_steps", C.c_int32),
                ("dampings", C.c_int32), ("outer_iters", C.c_int32), ("hyper", C.c_double),
                ("last_mlogp", C.c_double),
                ("mu", C.POINTER(C.c_double)), ("lambda2", C.POINTER(C.c_double)),
                ("offset", C.POINTER(C.c_double)), ("estimate_id", C.POINTER(C.c_int32)),
                ("start", C.POINTER(C.c_double)), ("beta", C.POINTER(C.c_double)),
                ("betahat", C.POINTER(C.c_double)), ("pseudoweight", C.POINTER(C.c_double))]


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "methylasso_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "-s", os.path.join(HERE, "liboracle.so")])
    if (force or not os.path.exists(REF_PATH)) and os.path.exists("/root/reference/src/core/gfl/gfl_tf.c"):
        subprocess.check_call(["make", "-C", HERE, "-s", "ref"])
    full = os.path.join(HERE, "_ref", "libmethylasso_ref.so")
    if (force or not os.path.exists(full)) and os.path.exists("/root/reference/src/signal_detection.cpp"):
        subprocess.check_call(["make", "-C", HERE, "-s", "ref_full"])


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp = C.c_void_p
        L.orc_set_dp.argtypes = [dp]
        L.orc_weighted_1d_flsa.argtypes = [C.c_int64, dp, dp, C.c_double, C.c_double, dp]
        L.orc_check_observations.argtypes = [C.c_int64] + [dp] * 6
        L.orc_detect.argtypes = [C.POINTER(OrcParams), C.c_int64] + [dp] * 6 + [C.POINTER(OrcResult)]
        L.orc_result_free.argtypes = [C.POINTER(OrcResult)]
        L.orc_segment_helper.argtypes = [C.c_int64] + [dp] * 5 + [C.c_double] + [dp] * 8
        L.orc_segment_helper.restype = C.c_int64
        L.orc_call_table.argtypes = [C.c_int64] + [dp] * 4 + [C.c_int64] + [dp] * 5 + [C.c_double] + [dp] * 10
        L.orc_call_table.restype = C.c_int64
        L.orc_wilcox_test.argtypes = [C.c_int64, dp, C.c_int64, dp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_p_adjust_fdr.argtypes = [C.c_int64, dp, dp]
        L.orc_p_adjust_fdr.restype = None
        L.orc_fast_diff_combine.argtypes = [C.c_int64, C.c_int32] + [dp] * 6
        L.orc_strerror.restype = C.c_char_p
        L.orc_tf_dp_weight.argtypes = [C.c_int] + [dp, dp, C.c_double] + [dp] * 6
        _lib = L
    return _lib


def ref_lib():
    """The reference's own tf_dp_weight compiled verbatim (None if oracle/_ref is absent)."""
    global _ref
    if _ref is None and os.path.exists(REF_PATH):
        R = C.CDLL(REF_PATH)
        R.tf_dp_weight.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 6
        _ref = R
    return _ref


def use_reference_dp(on=True):
    """Route the oracle's DP calls through the reference's verbatim kernel (if available)."""
    R = ref_lib()
    if on and R is not None:
        lib().orc_set_dp(C.cast(R.tf_dp_weight, C.c_void_p))
        return True
    lib().orc_set_dp(None)
    return False


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().orc_strerror(rc).decode())


def tf_dp_weight(y, w, lam, which="port"):
    """Raw DP (no clamp); which = 'port' (restatement) or 'reference' (verbatim gfl_tf.c)."""
    y, w = _f64(y).copy(), _f64(w).copy()
    n = y.size
    beta = np.zeros(n)
    x, a, b = np.zeros(2 * n + 2), np.zeros(2 * n + 2), np.zeros(2 * n + 2)
    tm, tp = np.zeros(n + 1), np.zeros(n + 1)
    fn = lib().orc_tf_dp_weight if which == "port" else ref_lib().tf_dp_weight
    fn(n, _p(y), _p(w), float(lam), _p(beta), _p(x), _p(a), _p(b), _p(tm), _p(tp))
    return beta, tm[:max(n - 1, 0)], tp[:max(n - 1, 0)]


def weighted_1d_flsa(y, wt, lambda2, lambda1=0.0):
    y, wt = _f64(y), _f64(wt)
    if y.size != wt.size:
        raise OracleError(2, "Input vectors must be of same size!")
    beta = np.zeros(y.size)
    _check(lib().orc_weighted_1d_flsa(y.size, _p(y), _p(wt), float(lambda2), float(lambda1), _p(beta)))
    return beta


def check_observations(t):
    return lib().orc_check_observations(
        t["pos"].size, _p(_i32(t["dataset"])), _p(_i32(t["condition"])), _p(_i32(t["replicate"])),
        _p(_i32(t["pos"])), _p(_i32(t["Nmeth"])), _p(_i32(t["coverage"])))


def detect(t, model=MODEL_FAST, lambda2=25.0, tol_val=0.01, max_iter=1, nouter=1, max_lambda2=None,
           lambda1=0.0, bf_per_kb=50.0, hyper=100.0, ref=1):
    """methylation_detection on one chromosome table (dict of int32 columns)."""
    p = OrcParams(lambda2, lambda2 + 1 if max_lambda2 is None else max_lambda2, lambda1, tol_val,
                  bf_per_kb, hyper, max_iter, nouter, ref, model, 0, 0)
    cols = [_i32(t[k]) for k in ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage")]
    res = OrcResult()
    rc = lib().orc_detect(C.byref(p), cols[0].size, *[_p(c) for c in cols], C.byref(res))
    _check(rc)
    try:
        n, P, E, D = res.n_rows, res.n_params, res.n_est, res.n_dsets

        def arr(ptr, k, dt=np.float64):
            return np.ctypeslib.as_array(ptr, shape=(k,)).astype(dt, copy=True)
        out = dict(mu=arr(res.mu, n), lambda2=arr(res.lambda2, E), offset=arr(res.offset, D),
                   converged=bool(res.converged), hyper=res.hyper, irls_steps=res.irls_steps,
                   dampings=res.dampings, outer_iters=res.outer_iters, last_mlogp=res.last_mlogp,
                   n_params=P, n_est=E, n_dsets=D,
                   estimates=dict(estimate_id=arr(res.estimate_id, E * P, np.int32),
                                  start=arr(res.start, E * P), beta=arr(res.beta, E * P),
                                  betahat=arr(res.betahat, E * P),
                                  pseudoweight=arr(res.pseudoweight, E * P)))
    finally:
        lib().orc_result_free(C.byref(res))
    return out


def segment_methylation_helper(est, tol_val=0.01):
    idx, start = _i32(est["estimate_id"]), _i32(est["start"])
    beta, bh, pw = _f64(est["beta"]), _f64(est["betahat"]), _f64(est["pseudoweight"])
    n = idx.size
    o = dict(estimate_id=np.zeros(n, np.int32), segment_id=np.zeros(n, np.int32),
             start=np.zeros(n, np.uint32), end=np.zeros(n, np.uint32), beta=np.zeros(n),
             betahat=np.zeros(n), pseudoweight=np.zeros(n), stdev=np.zeros(n))
    S = lib().orc_segment_helper(n, _p(idx), _p(start), _p(beta), _p(bh), _p(pw), float(tol_val),
                                 *[_p(o[k]) for k in ("estimate_id", "segment_id", "start", "end",
                                                      "beta", "betahat", "pseudoweight", "stdev")])
    if S < 0:
        _check(int(-S))
    return {k: v[:S].copy() for k, v in o.items()}


CALL_COLUMNS = (("estimate_id", np.int32), ("segment_id", np.int32), ("start", np.uint32), ("end", np.uint32),
                ("width", np.float64), ("beta", np.float64), ("coverage", np.float64), ("pseudoweight", np.float64),
                ("num_cpgs", np.int32), ("std", np.float64))


def call_table(pooled, seg, max_distance=500.0):
    """R/call.R:43-68.  pooled: dict(estimate_id, pos, Nmeth, coverage) per position and condition, sorted by
    (estimate_id, pos); seg: the helper's table (estimate_id, segment_id, start, end, pseudoweight)."""
    e, p, m, c = (_i32(pooled[k]) for k in ("estimate_id", "pos", "Nmeth", "coverage"))
    se, si = _i32(seg["estimate_id"]), _i32(seg["segment_id"])
    ss, sn = np.ascontiguousarray(seg["start"], np.uint32), np.ascontiguousarray(seg["end"], np.uint32)
    sp = _f64(seg["pseudoweight"])
    o = {k: np.zeros(e.size, dt) for k, dt in CALL_COLUMNS}
    n = lib().orc_call_table(e.size, _p(e), _p(p), _p(m), _p(c), se.size, _p(se), _p(si), _p(ss), _p(sn), _p(sp),
                             float(max_distance), *[_p(o[k]) for k, _ in CALL_COLUMNS])
    if n < 0:
        _check(int(-n))
    return {k: v[:n].copy() for k, v in o.items()}


def wilcox_test(x, y):
    """(W, p) of R's wilcox.test(x, y, conf.int = F) (R/call.R:315); p is nan when undefined."""
    x, y = _f64(x), _f64(y)
    w, p = C.c_double(), C.c_double()
    lib().orc_wilcox_test(x.size, _p(x), y.size, _p(y), C.byref(w), C.byref(p))
    return w.value, p.value


def p_adjust_fdr(p):
    """R's p.adjust(p, method = 'fdr') (R/call.R:360)."""
    p = _f64(p)
    q = np.empty_like(p)
    lib().orc_p_adjust_fdr(p.size, _p(p), _p(q))
    return q


def fast_diff_combine(est, n_params, n_est):
    beta, bh, pw = _f64(est["beta"]), _f64(est["betahat"]), _f64(est["pseudoweight"])
    ob, obh, opw = np.zeros_like(beta), np.zeros_like(bh), np.zeros_like(pw)
    _check(lib().orc_fast_diff_combine(n_params, n_est, _p(beta), _p(bh), _p(pw), _p(ob), _p(obh), _p(opw)))
    return dict(estimate_id=est["estimate_id"].copy(), start=est["start"].copy(), beta=ob, betahat=obh,
                pseudoweight=opw)


# ---------------------------------------------------------------------------------------------------
# The reference's OWN hot path (unmodified sources compiled against stub headers: oracle/Makefile,
# target ref_full; C entry points in oracle/ref_driver.cpp).  Used only to pin the restatement above.
REF_FULL_PATH = os.path.join(HERE, "_ref", "libmethylasso_ref.so")
_ref_full = None


class ReferenceError_(RuntimeError):
    pass


def ref_full_lib():
    """oracle/_ref/libmethylasso_ref.so (None if absent and the reference tree is not there to build it)."""
    global _ref_full
    if _ref_full is None:
        if not os.path.exists(REF_FULL_PATH) and os.path.exists("/root/reference/src/signal_detection.cpp"):
            subprocess.check_call(["make", "-C", HERE, "-s", "ref_full"])
        if not os.path.exists(REF_FULL_PATH):
            return None
        R = C.CDLL(REF_FULL_PATH)
        vp = C.c_void_p
        R.ref_last_error.restype = C.c_char_p
        R.ref_detect.argtypes = [C.c_int, C.c_long] + [vp] * 6 + [C.c_double, C.c_double, C.c_double, C.c_uint,
                                                                 C.c_double, C.c_double, C.c_uint, C.c_uint,
                                                                 C.c_int, C.c_int]
        R.ref_result_sizes.argtypes = [C.POINTER(C.c_long)] * 4
        R.ref_result_copy.argtypes = [vp] * 9
        R.ref_segment_helper.argtypes = [C.c_long] + [vp] * 5 + [C.c_double, C.POINTER(C.c_long)]
        R.ref_segments_copy.argtypes = [vp] * 8
        R.ref_weighted_1d_flsa.argtypes = [C.c_long, vp, vp, C.c_double, C.c_double, vp]
        _ref_full = R
    return _ref_full


def _ref_check(rc):
    if rc != 0:
        raise ReferenceError_(ref_full_lib().ref_last_error().decode())


def reference_detect(t, model=MODEL_FAST, lambda2=25.0, tol_val=0.01, max_iter=1, nouter=20, max_lambda2=1000.0,
                     bf_per_kb=50.0, hyper=100.0, ref=1):
    """The reference's signal_detection_fast_R / signal_detection_R / difference_detection_R
    (src/signal_detection.cpp:11,51; src/difference_detection.cpp:13) on one chromosome table."""
    R = ref_full_lib()
    cols = [_i32(t[k]) for k in ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage")]
    _ref_check(R.ref_detect(int(model), cols[0].size, *[_p(c) for c in cols], float(lambda2), float(max_lambda2),
                            float(hyper), int(nouter), float(bf_per_kb), float(tol_val), int(max_iter), int(ref), 0, 0))
    sz = [C.c_long() for _ in range(4)]
    _ref_check(R.ref_result_sizes(*[C.byref(s) for s in sz]))
    n, ne, nl, no = [s.value for s in sz]
    mu, lam, off = np.zeros(n), np.zeros(nl), np.zeros(no)
    conv = C.c_int()
    est = {k: np.zeros(ne) for k in ("estimate_id", "start", "beta", "betahat", "pseudoweight")}
    _ref_check(R.ref_result_copy(_p(mu), _p(lam), _p(off), C.cast(C.byref(conv), C.c_void_p),
                                 *[_p(est[k]) for k in ("estimate_id", "start", "beta", "betahat", "pseudoweight")]))
    est["estimate_id"] = est["estimate_id"].astype(np.int32)
    return dict(mu=mu, lambda2=lam, offset=off, converged=bool(conv.value), estimates=est)


def reference_segment_helper(est, tol_val=0.01):
    """The reference's segment_methylation_helper (src/methylation_helpers.cpp:6)."""
    R = ref_full_lib()
    idx, start = _i32(est["estimate_id"]), _i32(est["start"])
    beta, bh, pw = _f64(est["beta"]), _f64(est["betahat"]), _f64(est["pseudoweight"])
    S = C.c_long()
    _ref_check(R.ref_segment_helper(idx.size, _p(idx), _p(start), _p(beta), _p(bh), _p(pw), float(tol_val), C.byref(S)))
    keys = ("estimate_id", "segment_id", "start", "end", "beta", "betahat", "pseudoweight", "stdev")
    o = {k: np.zeros(S.value) for k in keys}
    _ref_check(R.ref_segments_copy(*[_p(o[k]) for k in keys]))
    return o


def reference_weighted_1d_flsa(y, wt, lambda2, lambda1=0.0):
    """The reference's weighted_1d_flsa (src/weighted_1d_flsa.cpp:7)."""
    R = ref_full_lib()
    y, wt = _f64(y), _f64(wt)
    out = np.zeros(y.size)
    _ref_check(R.ref_weighted_1d_flsa(y.size, _p(y), _p(wt), float(lambda2), float(lambda1), _p(out)))
    return out
