"""Host-side mirror of the reference's native interface, on top of the C ABI.

Same names, argument meaning, defaults and error behaviour as the five functions of the Rcpp
module `MethyLasso_cpp` (/root/reference/src/MethyLasso_cpp.cpp:13-28) and of the R wrappers that
loop over chromosomes (/root/reference/R/genome_wide.R:38-135, R/fast_diff.R:12-30).  A "data
frame" is a mapping of column name -> 1-D array (a pandas DataFrame works too); results are dicts
with the reference's list element names.

The genome-wide wrappers hand ALL chromosomes to the engine in one batched native call instead of
forking one R worker per chromosome (fork + CUDA do not mix).
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _native
from ._native import MlParams

MODEL_FAST, MODEL_FULL_SIGNAL, MODEL_FULL_DIFFERENCE = 0, 1, 2
_COLUMNS = ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage")


class MethyLassoError(RuntimeError):
    """An R-level error of the reference (Rcpp::stop / std::invalid_argument)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = int(code)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    # R hands Nmeth / coverage as doubles sometimes (MethyLasso.R:266-277); Rcpp coerces to int
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float64)


class _Pinned:
    """Owner of one page-locked allocation; numpy arrays made from it keep it alive through their `base`."""

    def __init__(self, lib, nbytes):
        self.lib, self.ptr = lib, C.c_void_p()
        rc = lib.ml_host_alloc(int(max(nbytes, 1)), C.byref(self.ptr))
        if rc:
            raise MethyLassoError(rc, lib.ml_last_error().decode())
        self.nbytes = int(max(nbytes, 1))
        self.__array_interface__ = {"data": (self.ptr.value, False), "shape": (self.nbytes,), "typestr": "|u1", "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                self.lib.ml_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(n, dtype, lib=None):
    """numpy array of n elements in page-locked host memory (downloads into it run at the PCIe rate)."""
    lib = lib or _native.load()
    dt = np.dtype(dtype)
    own = _Pinned(lib, int(n) * dt.itemsize)
    return np.asarray(own)[:int(n) * dt.itemsize].view(dt)


class Engine:
    """One GPU, one stream.  Creating it without a CUDA device raises (no CPU path)."""

    def __init__(self, device: int = -1):
        self.lib = _native.load()
        h = C.c_void_p()
        rc = self.lib.ml_engine_create(int(device), C.byref(h))
        if rc:
            raise MethyLassoError(rc, self.lib.ml_last_error().decode())
        self.h = h
        self.want_mu = True
        # what lives on this engine (batches, results, gather windows): released before the engine itself, whatever order
        # the caller - or the interpreter at exit - lets go of them in
        self._children = weakref.WeakSet()

    def _adopt(self, child):
        self._children.add(child)

    def pinned_buffer(self, name, n, dtype):
        """View of n elements of an engine-owned page-locked buffer called `name` (allocated on first use, grown
        by half when too small, released with the engine).  A later request for the same name reuses the memory."""
        dt = np.dtype(dtype)
        cache = self.__dict__.setdefault("_pinned", {})
        cur = cache.get(name)
        need = int(n) * dt.itemsize
        if cur is None or cur.nbytes < need:
            cur = np.asarray(_Pinned(self.lib, need + need // 2))
            cache[name] = cur
        return cur[:need].view(dt)

    def close(self):
        if getattr(self, "h", None):
            for child in list(getattr(self, "_children", ())):
                try:
                    (getattr(child, "free", None) or child.close)()
                except Exception:
                    pass
            self.__dict__.pop("_pinned", None)
            self.lib.ml_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise MethyLassoError(rc, self.lib.ml_last_error().decode())

    def set(self, key: str, value: int):
        self._check(self.lib.ml_engine_set(self.h, key.encode(), int(value)))
        if key == "want_mu":
            self.want_mu = bool(value)

    @property
    def launch_count(self) -> int:
        return int(self.lib.ml_engine_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.ml_engine_stream(self.h) or 0)

    def synchronize(self):
        self._check(self.lib.ml_engine_synchronize(self.h))

    def set_stream(self, cuda_stream: int):
        """Launch on a caller-owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(self.lib.ml_engine_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def profile(self):
        """{kernel name: (total ms, launches)} accumulated since the last profile_reset()."""
        out = {}
        for i in range(self.lib.ml_engine_profile_count(self.h)):
            name = C.create_string_buffer(128)
            ms, n = C.c_double(), C.c_longlong()
            self._check(self.lib.ml_engine_profile_get(self.h, i, name, 128, C.byref(ms), C.byref(n)))
            out[name.value.decode()] = (ms.value, n.value)
        return out

    def profile_work(self):
        """{kernel name: algorithmic bytes} for the kernels whose work is data-dependent (FLSA chunk kernels)."""
        out = {}
        for i in range(self.lib.ml_engine_profile_count(self.h)):
            name = C.create_string_buffer(128)
            ms, n, b = C.c_double(), C.c_longlong(), C.c_double()
            self._check(self.lib.ml_engine_profile_get(self.h, i, name, 128, C.byref(ms), C.byref(n)))
            self._check(self.lib.ml_engine_profile_work(self.h, i, C.byref(b)))
            if b.value >= 0:
                out[name.value.decode()] = b.value
        return out

    def profile_reset(self):
        self._check(self.lib.ml_engine_profile_reset(self.h))

    def profile_spans(self):
        """[(kernel name, start ms, end ms)] of the profiled launches since the last reset, on the process-wide clock
        (set("profile_timeline", 1) next to set("profile", 1))."""
        n = self.lib.ml_engine_profile_spans(self.h, 0, 0, None, 0, None, None)
        if n <= 0:
            return []
        names = C.create_string_buffer(n * 48)
        t0, t1 = (C.c_float * n)(), (C.c_float * n)()
        k = self.lib.ml_engine_profile_spans(self.h, 0, n, names, 48, t0, t1)
        raw = names.raw
        return [(raw[i * 48:(i + 1) * 48].split(b"\0", 1)[0].decode(), float(t0[i]), float(t1[i])) for i in range(k)]

    def segment_call_table(self, pooled, seg, max_distance=500.0):
        """R/call.R:43-68 from host tables: pooled = dict(estimate_id, pos, Nmeth, coverage) per position and
        condition sorted by (estimate_id, pos); seg = the table of segment_methylation_helper."""
        e, p, m, c = (np.ascontiguousarray(pooled[k], np.int32) for k in ("estimate_id", "pos", "Nmeth", "coverage"))
        se = np.ascontiguousarray(seg["estimate_id"], np.int32)
        si = np.ascontiguousarray(seg["segment_id"], np.int32)
        ss = np.ascontiguousarray(seg["start"], np.uint32)
        sn = np.ascontiguousarray(seg["end"], np.uint32)
        sp = _f64(seg["pseudoweight"])
        cols = [(k, dt) for k, dt in Result.CALL_COLUMNS if k != "job"]
        out = {k: np.empty(max(e.size, 1), dt) for k, dt in cols}
        n = C.c_int64(0)
        self._check(self.lib.ml_segment_call_table(self.h, e.size, _p(e), _p(p), _p(m), _p(c), se.size, _p(se), _p(si),
                                                   _p(ss), _p(sn), _p(sp), float(max_distance), C.byref(n),
                                                   *[_p(out[k]) for k, _ in cols]))
        return {k: v[:n.value].copy() for k, v in out.items()}

    # ---- region statistics (R/call.R:315-317, :360) -----------------------------------------------
    def wilcoxon_groups(self, x_ptr, x, y_ptr, y, want_stat=False):
        """wilcox.test(x_g, y_g, conf.int = F)$p.value for every group g (CSR offsets x_ptr / y_ptr); NaN where R
        gives NaN or fails (R/call.R:317 maps those to 1)."""
        x_ptr = np.ascontiguousarray(x_ptr, dtype=np.int64)
        y_ptr = np.ascontiguousarray(y_ptr, dtype=np.int64)
        if x_ptr.size != y_ptr.size or x_ptr.size < 1:
            raise MethyLassoError(2, "x_ptr and y_ptr must have n_groups + 1 entries")
        x, y = _f64(x), _f64(y)
        if x_ptr[-1] > x.size or y_ptr[-1] > y.size or x_ptr[0] < 0 or y_ptr[0] < 0:
            raise MethyLassoError(19, "offsets point outside the samples")
        ng = x_ptr.size - 1
        p = np.empty(ng)
        w = np.empty(ng) if want_stat else None
        self._check(self.lib.ml_wilcoxon_groups(self.h, ng, _p(x_ptr), _p(x), _p(y_ptr), _p(y), _p(p),
                                                _p(w) if want_stat else None))
        return (p, w) if want_stat else p

    def p_adjust_fdr(self, p):
        """p.adjust(p, method = 'fdr')."""
        p = _f64(p)
        q = np.empty_like(p)
        self._check(self.lib.ml_p_adjust_fdr(self.h, p.size, _p(p), _p(q)))
        return q

    # ---- weighted_1d_flsa ---------------------------------------------------------------------
    def weighted_1d_flsa(self, y, wt, lambda2, lambda1=0.0):
        y, wt = _f64(y), _f64(wt)
        if y.size != wt.size:  # src/weighted_1d_flsa.cpp:9
            raise MethyLassoError(2, "Input vectors must be of same size!")
        beta = np.empty_like(y)
        self._check(self.lib.ml_weighted_1d_flsa(self.h, y.size, _p(y), _p(wt), float(lambda2),
                                                 float(lambda1), _p(beta)))
        return beta

    def weighted_1d_flsa_batch(self, ptr, y, wt, lambda2, lambda1=0.0):
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        y, wt = _f64(y), _f64(wt)
        lam = _f64(np.broadcast_to(lambda2, (ptr.size - 1,)))
        beta = np.empty_like(y)
        self._check(self.lib.ml_weighted_1d_flsa_batch(self.h, ptr.size - 1, _p(ptr), _p(y), _p(wt), _p(lam),
                                                       float(lambda1), _p(beta)))
        return beta

    # ---- detection ------------------------------------------------------------------------------
    def upload(self, job_ptr, cols) -> "Batch":
        job_ptr = np.ascontiguousarray(job_ptr, dtype=np.int64)
        arrs = [_i32(cols[k]) for k in _COLUMNS]
        h = C.c_void_p()
        self._check(self.lib.ml_batch_upload(self.h, job_ptr.size - 1, _p(job_ptr), *[_p(a) for a in arrs],
                                             C.byref(h), None))
        return Batch(self, h, job_ptr.size - 1)

    def stage(self, job_ptr, cols) -> "Batch":
        """First half of `upload` (host work only: run table, counts packed for the bus); finish with `commit`.  The
        columns must stay alive and unchanged until then."""
        job_ptr = np.ascontiguousarray(job_ptr, dtype=np.int64)
        c = {k: _i32(cols[k]) for k in _COLUMNS}
        h = C.c_void_p()
        self._check(self.lib.ml_batch_stage(self.h, job_ptr.size - 1, _p(job_ptr), *[_p(c[k]) for k in _COLUMNS], C.byref(h)))
        b = Batch(self, h, job_ptr.size - 1)
        b._keep = (job_ptr, c)
        return b

    def commit(self, batch: "Batch") -> "Batch":
        self._check(self.lib.ml_batch_commit(self.h, batch.h))
        batch._keep = None
        return batch

    def upload_runs(self, job_ptr, dset_ptr, dset_row, dset_condition, dset_replicate, cols) -> "Batch":
        """ml_batch_upload_runs: the caller hands the dataset runs (first row, condition, replicate of every dataset
        of every job) instead of the three factor columns; cols needs pos, Nmeth, coverage only."""
        job_ptr = np.ascontiguousarray(job_ptr, dtype=np.int64)
        dset_ptr = np.ascontiguousarray(dset_ptr, dtype=np.int64)
        dset_row = np.ascontiguousarray(dset_row, dtype=np.int64)
        dc, dr = _i32(dset_condition), _i32(dset_replicate)
        arrs = [_i32(cols[k]) for k in ("pos", "Nmeth", "coverage")]
        h = C.c_void_p()
        self._check(self.lib.ml_batch_upload_runs(self.h, job_ptr.size - 1, _p(job_ptr), _p(dset_ptr), _p(dset_row), _p(dc),
                                                  _p(dr), *[_p(a) for a in arrs], C.byref(h)))
        return Batch(self, h, job_ptr.size - 1)

    def detect(self, batch: "Batch", params: MlParams) -> "Result":
        h = C.c_void_p()
        status = np.zeros(max(batch.njobs, 1), np.int32)
        self._check(self.lib.ml_batch_detect(self.h, batch.h, C.byref(params), C.byref(h), _p(status)))
        return Result(self, h, status[:batch.njobs].copy())

    def detect_batch(self, job_ptr, cols, params: MlParams) -> "Result":
        """upload + detect through the single host-buffer entry point (ml_detect_batch)."""
        job_ptr = np.ascontiguousarray(job_ptr, dtype=np.int64)
        arrs = [_i32(cols[k]) for k in _COLUMNS]
        nj = job_ptr.size - 1
        h = C.c_void_p()
        status = np.zeros(max(nj, 1), np.int32)
        self._check(self.lib.ml_detect_batch(self.h, nj, _p(job_ptr), *[_p(a) for a in arrs],
                                             C.byref(params), C.byref(h), _p(status)))
        return Result(self, h, status[:nj].copy())

    # ---- segment_methylation_helper -------------------------------------------------------------
    def segment_methylation_helper(self, df, tol_val=0.01):
        idx, start = _i32(df["estimate_id"]), _i32(df["start"])
        beta, bh, pw = _f64(df["beta"]), _f64(df["betahat"]), _f64(df["pseudoweight"])
        n = idx.size
        out = _seg_buffers(n, with_job=False)
        ns = C.c_int64(0)
        self._check(self.lib.ml_segment_methylation_helper(
            self.h, n, _p(idx), _p(start), _p(beta), _p(bh), _p(pw), float(tol_val), C.byref(ns),
            *[_p(out[k]) for k in _SEG_COLS]))
        return {k: v[:ns.value].copy() for k, v in out.items()}


_SEG_COLS = ("estimate_id", "segment_id", "start", "end", "beta", "betahat", "pseudoweight", "stdev")


def _seg_buffers(n, with_job):
    out = dict(estimate_id=np.empty(n, np.int32), segment_id=np.empty(n, np.int32),
               start=np.empty(n, np.uint32), end=np.empty(n, np.uint32), beta=np.empty(n),
               betahat=np.empty(n), pseudoweight=np.empty(n), stdev=np.empty(n))
    if with_job:
        out = dict(job=np.empty(n, np.int32), **out)
    return out


class Batch:
    def __init__(self, eng, h, njobs):
        self.eng, self.h, self.njobs = eng, h, njobs
        eng._adopt(self)

    def free(self):
        if self.h:
            self.eng.lib.ml_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Result:
    """Outputs of one detect call, resident on the GPU until copied."""

    def __init__(self, eng, h, status):
        self.eng, self.h, self.status = eng, h, status
        eng._adopt(self)
        self.njobs = status.size
        self.has_mu = bool(getattr(eng, "want_mu", True))    # one-step fits of an engine with want_mu = 0 store no mu

    def free(self):
        if self.h:
            self.eng.lib.ml_result_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def sizes(self, job):
        n, p = C.c_int64(), C.c_int64()
        e, d, s = C.c_int32(), C.c_int32(), C.c_int32()
        self.eng._check(self.eng.lib.ml_result_job_sizes(self.h, job, C.byref(n), C.byref(p), C.byref(e),
                                                         C.byref(d), C.byref(s)))
        return n.value, p.value, e.value, d.value, s.value

    def info(self, job):
        v = [C.c_int32() for _ in range(4)]
        hy, ml = C.c_double(), C.c_double()
        self.eng._check(self.eng.lib.ml_result_job_info(self.h, job, *[C.byref(x) for x in v], C.byref(hy),
                                                        C.byref(ml)))
        return dict(converged=bool(v[0].value), irls_steps=v[1].value, dampings=v[2].value,
                    outer_iters=v[3].value, hyper=hy.value, mlogp=ml.value)

    def job(self, job, want_mu=True):
        """The Rcpp List of one job (src/methylation_detection.hpp:47-53)."""
        n, P, E, D, status = self.sizes(job)
        if status:
            raise MethyLassoError(status, self.eng.lib.ml_strerror(status).decode())
        mu = np.empty(n) if (want_mu and self.has_mu) else None
        lam, off = np.empty(E), np.empty(D)
        est = dict(estimate_id=np.empty(E * P, np.int32), start=np.empty(E * P), beta=np.empty(E * P),
                   betahat=np.empty(E * P), pseudoweight=np.empty(E * P))
        self.eng._check(self.eng.lib.ml_result_copy_job(
            self.h_eng(), self.h, job, _p(mu), _p(lam), _p(off), _p(est["estimate_id"]), _p(est["start"]),
            _p(est["beta"]), _p(est["betahat"]), _p(est["pseudoweight"])))
        out = dict(mu=mu, lambda2=lam, offset=off, estimates=est, n_params=P, n_est=E, n_dsets=D)
        out.update(self.info(job))
        return out

    def h_eng(self):
        return self.eng.h

    def fast_diff_combine(self):
        self.eng._check(self.eng.lib.ml_result_fast_diff_combine(self.eng.h, self.h))

    def lambda_sweep(self, lambdas, tol_val=0.01, lambda1=0.0, want_beta=False):
        """ml_result_lambda_sweep: the one-step FAST fit of this result redone for every lambda2 of `lambdas` in one batched
        FLSA call on the resident pseudodata.  Returns the number of segments per value (and, want_beta, the coefficients
        as an array [len(lambdas), total parameters] in copy_all order)."""
        lam = np.ascontiguousarray(lambdas, np.float64)
        nseg = np.zeros(lam.size, np.int64)
        beta = None
        if want_beta:
            npar = int(sum(self.sizes(j)[1] * self.sizes(j)[2] for j in range(self.njobs) if self.sizes(j)[4] == 0))
            beta = np.empty((lam.size, npar))
        self.eng._check(self.eng.lib.ml_result_lambda_sweep(self.eng.h, self.h, lam.size, _p(lam), float(lambda1), float(tol_val),
                                                            _p(nseg), _p(beta)))
        return (nseg, beta) if want_beta else nseg

    def segments(self, tol_val=0.01, reuse_pinned=False):
        """segment_methylation_helper on the resident estimates of every valid job.  reuse_pinned=True downloads
        into engine-owned page-locked buffers and returns views of them (valid until the next such call)."""
        ns = C.c_int64(0)
        self.eng._check(self.eng.lib.ml_result_segments(self.eng.h, self.h, float(tol_val), C.byref(ns),
                                                        *([None] * 9)))
        out = _seg_buffers(max(ns.value, 1), with_job=True)
        if reuse_pinned:
            out = {k: self.eng.pinned_buffer("seg_" + k, v.size, v.dtype) for k, v in out.items()}
        self.eng._check(self.eng.lib.ml_result_segments(
            self.eng.h, self.h, float(tol_val), C.byref(ns), _p(out["job"]),
            *[_p(out[k]) for k in _SEG_COLS]))
        return {k: (v[:ns.value] if reuse_pinned else v[:ns.value].copy()) for k, v in out.items()}


    def segments_count(self, tol_val=0.01) -> int:
        """Run the segmentation on the device and leave the table there (fetch later with segments())."""
        ns = C.c_int64(0)
        self.eng._check(self.eng.lib.ml_result_segments(self.eng.h, self.h, float(tol_val), C.byref(ns),
                                                        *([None] * 9)))
        return ns.value

    def copy_all(self, mu=None, offset=None, estimate_id=None, start=None, beta=None, betahat=None,
                 pseudoweight=None):
        """All valid jobs concatenated in job order into caller buffers (one D2H per column)."""
        self.eng._check(self.eng.lib.ml_result_copy_all(self.eng.h, self.h, _p(mu), _p(offset), _p(estimate_id),
                                                        _p(start), _p(beta), _p(betahat), _p(pseudoweight)))

    def copy_all_async(self, mu=None, offset=None, estimate_id=None, start=None, beta=None, betahat=None,
                       pseudoweight=None):
        """copy_all that only enqueues the transfers (pinned buffers!); engine.synchronize() completes them."""
        self.eng._check(self.eng.lib.ml_result_copy_all_async(self.eng.h, self.h, _p(mu), _p(offset),
                                                              _p(estimate_id), _p(start), _p(beta), _p(betahat),
                                                              _p(pseudoweight)))

    def pmd_segments(self, tol_val=0.01, pmd_lambda2=1000.0, fetch=True):
        """R/call.R:93-95 on the resident segments (see ml_result_pmd_segments)."""
        ns = C.c_int64(0)
        self.eng._check(self.eng.lib.ml_result_pmd_segments(self.eng.h, self.h, float(tol_val), float(pmd_lambda2),
                                                            C.byref(ns), *([None] * 9)))
        if not fetch:
            return ns.value
        n = max(ns.value, 1)
        out = dict(job=np.empty(n, np.int32), estimate_id=np.empty(n, np.int32), segment_id=np.empty(n, np.int32),
                   start=np.empty(n, np.uint32), end=np.empty(n, np.uint32), fused_std=np.empty(n),
                   std=np.empty(n), pseudoweight=np.empty(n))
        self.eng._check(self.eng.lib.ml_result_pmd_segments(
            self.eng.h, self.h, float(tol_val), float(pmd_lambda2), C.byref(ns), None,
            *[_p(out[k]) for k in ("job", "estimate_id", "segment_id", "start", "end", "fused_std", "std",
                                   "pseudoweight")]))
        return {k: v[:ns.value].copy() for k, v in out.items()}

    CALL_COLUMNS = (("job", np.int32), ("estimate_id", np.int32), ("segment_id", np.int32), ("start", np.uint32),
                    ("end", np.uint32), ("width", np.float64), ("beta", np.float64), ("coverage", np.float64),
                    ("pseudoweight", np.float64), ("num_cpgs", np.int32), ("std", np.float64))

    DIFF_COLUMNS = (("job", np.int32), ("condition", np.int32), ("estimate_id", np.int32), ("segment_id", np.int32),
                    ("start", np.uint32), ("end", np.uint32), ("width", np.float64), ("beta", np.float64),
                    ("std", np.float64), ("coverage", np.float64), ("num_cpgs", np.int32), ("beta_ref", np.float64),
                    ("std_ref", np.float64), ("coverage_ref", np.float64), ("num_cpgs_ref", np.int32),
                    ("pseudoweight", np.float64), ("diff", np.float64), ("cohen_d", np.float64), ("pval", np.float64),
                    ("coverage_score", np.float64))

    def call_differences(self, batch, ref_condition, min_diff=0.1, min_delta_diff=0.1, tol_val=0.01,
                         max_distance=500.0, fetch=True, reuse_pinned=False):
        """R/call.R:254-333 (call_differences_singlechr) for every job of a difference_detection_fast result
        (fast_diff_combine applied); `batch` is the resident input the result was computed from."""
        n = C.c_int64(0)
        args = (self.h_eng(), batch.h, self.h, int(ref_condition), float(min_diff), float(min_delta_diff),
                float(tol_val), float(max_distance))
        self.eng._check(self.eng.lib.ml_result_call_differences(*args, C.byref(n), *([None] * 20)))
        if not fetch:
            return n.value
        if reuse_pinned:
            # engine-owned page-locked buffers, grown as needed and REUSED by the next call with reuse_pinned=True
            # (page-locking costs far more than it saves when done per call): the returned arrays are views of them
            out = {k: self.eng.pinned_buffer("diff_" + k, max(n.value, 1), dt) for k, dt in self.DIFF_COLUMNS}
        else:
            out = {k: np.empty(max(n.value, 1), dt) for k, dt in self.DIFF_COLUMNS}
        self.eng._check(self.eng.lib.ml_result_call_differences(*args, C.byref(n), *[_p(out[k]) for k, _ in self.DIFF_COLUMNS]))
        return {k: v[:n.value] for k, v in out.items()}

    def call_table(self, tol_val=0.01, max_distance=500.0, fetch=True, reuse_pinned=False):
        """R/call.R:43-68 on the resident result (see ml_result_call_table): one row per gap-split segment."""
        n = C.c_int64(0)
        self.eng._check(self.eng.lib.ml_result_call_table(self.eng.h, self.h, float(tol_val), float(max_distance),
                                                          C.byref(n), *([None] * 11)))
        if not fetch:
            return n.value
        if reuse_pinned:
            out = {k: self.eng.pinned_buffer("call_" + k, max(n.value, 1), dt) for k, dt in self.CALL_COLUMNS}
        else:
            out = {k: np.empty(max(n.value, 1), dt) for k, dt in self.CALL_COLUMNS}
        self.eng._check(self.eng.lib.ml_result_call_table(self.eng.h, self.h, float(tol_val), float(max_distance),
                                                          C.byref(n), *[_p(out[k]) for k, _ in self.CALL_COLUMNS]))
        return {k: (v[:n.value] if reuse_pinned else v[:n.value].copy()) for k, v in out.items()}

    def call_lmr_umr(self, tol_val=0.01, max_distance=500.0, **thresholds):
        """R/call.R:75-91 on the resident call table: is_admissible, is_V, is_flank, is_sep (1 TRUE, 0 FALSE, -1 NA),
        flank_dist, flank_beta (NaN = NA) and category (0 NA, 1 LMR, 2 UMR), one entry per row of call_table()."""
        from ._native import MlLmrParams
        p = MlLmrParams(max_distance=float(max_distance), **thresholds)
        n = self.call_table(tol_val, max_distance, fetch=False)
        i8 = lambda: np.empty(max(n, 1), np.int8)
        out = dict(is_admissible=i8(), is_V=i8(), is_flank=i8(), flank_dist=np.empty(max(n, 1)),
                   flank_beta=np.empty(max(n, 1)), is_sep=i8(), category=i8())
        nn = C.c_int64(0)
        self.eng._check(self.eng.lib.ml_result_call_lmr_umr(
            self.eng.h, self.h, float(tol_val), float(max_distance), C.byref(p), C.byref(nn),
            *[_p(out[k]) for k in ("is_admissible", "is_V", "is_flank", "flank_dist", "flank_beta", "is_sep", "category")]))
        return {k: v[:nn.value] for k, v in out.items()}

    VALLEY_COLUMNS = (("job", np.int32), ("condition", np.int32), ("start", np.uint32), ("end", np.uint32),
                      ("width", np.float64), ("beta", np.float64), ("coverage", np.float64),
                      ("pseudoweight", np.float64), ("num_cpgs", np.int32))

    def call_valleys(self, tol_val=0.01, max_distance=500.0, valley_max_beta=0.1, pmd_valley_min_width=5000.0,
                     min_num_cpgs=4, **lmr_thresholds):
        """R/call.R:100-133 on the resident call table: the "UMR-DMV" valleys (one row each, ordered by job,
        condition, start; category is "UMR-DMV" for every row) and, per row of call_table(), row_in_valley (the row
        overlaps a valley and is replaced by it, :132-133) and the low.id.width / low.id.beta columns of :106."""
        from ._native import MlLmrParams, MlValleyParams
        lp = MlLmrParams(max_distance=float(max_distance), min_num_cpgs=int(min_num_cpgs), **lmr_thresholds)
        vp = MlValleyParams(valley_max_beta=float(valley_max_beta), pmd_valley_min_width=float(pmd_valley_min_width),
                            max_distance=float(max_distance), min_num_cpgs=int(min_num_cpgs))
        nv, nr = C.c_int64(0), C.c_int64(0)
        args = (self.eng.h, self.h, float(tol_val), float(max_distance), C.byref(lp), C.byref(vp))
        self.eng._check(self.eng.lib.ml_result_call_valleys(*args, C.byref(nv), *([None] * 9), C.byref(nr), *([None] * 3)))
        out = {k: np.empty(max(nv.value, 1), dt) for k, dt in self.VALLEY_COLUMNS}
        rows = dict(row_in_valley=np.empty(max(nr.value, 1), np.int8), low_id_width=np.empty(max(nr.value, 1)),
                    low_id_beta=np.empty(max(nr.value, 1)))
        self.eng._check(self.eng.lib.ml_result_call_valleys(
            *args, C.byref(nv), *[_p(out[k]) for k, _ in self.VALLEY_COLUMNS], C.byref(nr),
            *[_p(rows[k]) for k in ("row_in_valley", "low_id_width", "low_id_beta")]))
        res = {k: v[:nv.value] for k, v in out.items()}
        res.update({k: v[:nr.value] for k, v in rows.items()})
        return res

    MERGED_COLUMNS = (("job", np.int32), ("condition", np.int32), ("segment_id", np.int32), ("start", np.uint32),
                      ("end", np.uint32), ("category", np.int8), ("src_row", np.int32), ("src_valley", np.int32),
                      ("follows_large_gap", np.int8))
    PIECE_COLUMNS = (("job", np.int32), ("estimate_id", np.int32), ("segment_id", np.int32), ("start", np.uint32),
                     ("end", np.uint32), ("fused_std", np.float64), ("pseudoweight", np.float64))

    def call_pmd_split(self, tol_val=0.01, max_distance=500.0, pmd_lambda2=1000.0, split_pmds=True, valley_max_beta=0.1,
                       pmd_valley_min_width=5000.0, min_num_cpgs=4, **lmr_thresholds):
        """R/call.R:131-164 on the resident tables: (seg_call after :140, split_call after :168) as two dicts of
        columns (see ml_result_call_pmd_split)."""
        from ._native import MlLmrParams, MlValleyParams
        lp = MlLmrParams(max_distance=float(max_distance), min_num_cpgs=int(min_num_cpgs), **lmr_thresholds)
        vp = MlValleyParams(valley_max_beta=float(valley_max_beta), pmd_valley_min_width=float(pmd_valley_min_width),
                            max_distance=float(max_distance), min_num_cpgs=int(min_num_cpgs))
        nm, npc = C.c_int64(0), C.c_int64(0)
        args = (self.eng.h, self.h, float(tol_val), float(max_distance), float(pmd_lambda2), C.byref(lp), C.byref(vp),
                1 if split_pmds else 0)
        self.eng._check(self.eng.lib.ml_result_call_pmd_split(*args, C.byref(nm), *([None] * 9), C.byref(npc), *([None] * 7)))
        merged = {k: np.empty(max(nm.value, 1), dt) for k, dt in self.MERGED_COLUMNS}
        pieces = {k: np.empty(max(npc.value, 1), dt) for k, dt in self.PIECE_COLUMNS}
        self.eng._check(self.eng.lib.ml_result_call_pmd_split(
            *args, C.byref(nm), *[_p(merged[k]) for k, _ in self.MERGED_COLUMNS], C.byref(npc),
            *[_p(pieces[k]) for k, _ in self.PIECE_COLUMNS]))
        return ({k: v[:nm.value] for k, v in merged.items()}, {k: v[:npc.value] for k, v in pieces.items()})

    PMD_COLUMNS = (("job", np.int32), ("condition", np.int32), ("category", np.int8), ("segment_id", np.int32),
                   ("start", np.uint32), ("end", np.uint32), ("width", np.float64), ("num_cpgs", np.int32),
                   ("fused_std", np.float64), ("beta", np.float64), ("std", np.float64))

    def call_pmds(self, tol_val=0.01, max_distance=500.0, pmd_lambda2=1000.0, pmd_std_threshold=0.15, pmd_max_beta=0.75,
                  split_pmds=True, merge_pmds=True, valley_max_beta=0.1, pmd_valley_min_width=5000.0, min_num_cpgs=4,
                  **lmr_thresholds):
        """R/call.R:169-199 on the resident tables: pmd_call after :198, one row per merged region; category 1 = "PMD",
        0 = "other" (see ml_result_call_pmds)."""
        from ._native import MlLmrParams, MlPmdParams, MlValleyParams
        lp = MlLmrParams(max_distance=float(max_distance), min_num_cpgs=int(min_num_cpgs), **lmr_thresholds)
        vp = MlValleyParams(valley_max_beta=float(valley_max_beta), pmd_valley_min_width=float(pmd_valley_min_width),
                            max_distance=float(max_distance), min_num_cpgs=int(min_num_cpgs))
        pp = MlPmdParams(pmd_lambda2=float(pmd_lambda2), pmd_std_threshold=float(pmd_std_threshold),
                         pmd_max_beta=float(pmd_max_beta), split_pmds=1 if split_pmds else 0, merge_pmds=1 if merge_pmds else 0)
        n = C.c_int64(0)
        args = (self.eng.h, self.h, float(tol_val), float(max_distance), C.byref(lp), C.byref(vp), C.byref(pp))
        self.eng._check(self.eng.lib.ml_result_call_pmds(*args, C.byref(n), *([None] * 11)))
        out = {k: np.empty(max(n.value, 1), dt) for k, dt in self.PMD_COLUMNS}
        self.eng._check(self.eng.lib.ml_result_call_pmds(*args, C.byref(n), *[_p(out[k]) for k, _ in self.PMD_COLUMNS]))
        return {k: v[:n.value] for k, v in out.items()}

    STATUS_COLUMNS = (("job", np.int32), ("condition", np.int32), ("start", np.uint32), ("end", np.uint32),
                      ("category", np.int8), ("pmd_status", np.int8), ("merged_row", np.int32), ("src_row", np.int32),
                      ("src_valley", np.int32))

    def call_pmd_status(self, tol_val=0.01, max_distance=500.0, pmd_lambda2=1000.0, pmd_std_threshold=0.15,
                        pmd_max_beta=0.75, split_pmds=True, merge_pmds=True, valley_max_beta=0.1,
                        pmd_valley_min_width=5000.0, min_num_cpgs=4, **lmr_thresholds):
        """R/call.R:201-215 on the resident tables: seg_call after :212 (see ml_result_call_pmd_status); the rows with
        category != 0 are the reference's `lmr_umr_valley` table (drop = T)."""
        from ._native import MlLmrParams, MlPmdParams, MlValleyParams
        lp = MlLmrParams(max_distance=float(max_distance), min_num_cpgs=int(min_num_cpgs), **lmr_thresholds)
        vp = MlValleyParams(valley_max_beta=float(valley_max_beta), pmd_valley_min_width=float(pmd_valley_min_width),
                            max_distance=float(max_distance), min_num_cpgs=int(min_num_cpgs))
        pp = MlPmdParams(pmd_lambda2=float(pmd_lambda2), pmd_std_threshold=float(pmd_std_threshold),
                         pmd_max_beta=float(pmd_max_beta), split_pmds=1 if split_pmds else 0, merge_pmds=1 if merge_pmds else 0)
        n = C.c_int64(0)
        args = (self.eng.h, self.h, float(tol_val), float(max_distance), C.byref(lp), C.byref(vp), C.byref(pp))
        self.eng._check(self.eng.lib.ml_result_call_pmd_status(*args, C.byref(n), *([None] * 9)))
        out = {k: np.empty(max(n.value, 1), dt) for k, dt in self.STATUS_COLUMNS}
        self.eng._check(self.eng.lib.ml_result_call_pmd_status(*args, C.byref(n), *[_p(out[k]) for k, _ in self.STATUS_COLUMNS]))
        return {k: v[:n.value] for k, v in out.items()}

    REGION_SEG_COLUMNS = (("job", np.int32), ("condition", np.int32), ("start", np.uint32), ("end", np.uint32),
                          ("width", np.float64), ("segment_id", np.int32), ("beta", np.float64), ("coverage", np.float64),
                          ("pseudoweight", np.float64), ("num_cpgs", np.int32), ("std", np.float64), ("category", np.int8),
                          ("pmd_status", np.int8))
    REGION_PMD_COLUMNS = (("job", np.int32), ("condition", np.int32), ("start", np.uint32), ("end", np.uint32),
                          ("width", np.float64), ("segment_id", np.int32), ("beta", np.float64), ("std", np.float64),
                          ("category", np.int8), ("num_cpgs", np.int32), ("fused_std", np.float64))

    def segment_methylation_tables(self, tol_val=0.01, max_distance=500.0, pmd_lambda2=1000.0, pmd_std_threshold=0.15,
                                   pmd_max_beta=0.75, split_pmds=True, merge_pmds=True, valley_max_beta=0.1,
                                   pmd_valley_min_width=5000.0, min_num_cpgs=4, drop=True, reuse_pinned=False,
                                   **lmr_thresholds):
        """ml_result_segment_methylation: the whole rule engine of R/call.R:31-221 on the resident fit and the two
        tables the reference returns (`lmr_umr_valley`, `pmd`), assembled on the device.  Returns (seg, pmd): dicts of
        columns with integer codes (category 0 NA / 1 LMR / 2 UMR / 3 UMR-DMV, pmd_status -1 NA / 0 other / 1 PMD; pmd
        category 1 PMD / 0 other); `job` indexes the jobs of the batch."""
        from ._native import MlLmrParams, MlPmdParams, MlValleyParams
        lp = MlLmrParams(max_distance=float(max_distance), min_num_cpgs=int(min_num_cpgs), **lmr_thresholds)
        vp = MlValleyParams(valley_max_beta=float(valley_max_beta), pmd_valley_min_width=float(pmd_valley_min_width),
                            max_distance=float(max_distance), min_num_cpgs=int(min_num_cpgs))
        pp = MlPmdParams(pmd_lambda2=float(pmd_lambda2), pmd_std_threshold=float(pmd_std_threshold),
                         pmd_max_beta=float(pmd_max_beta), split_pmds=1 if split_pmds else 0, merge_pmds=1 if merge_pmds else 0)
        ns, npm = C.c_int64(0), C.c_int64(0)
        args = (self.eng.h, self.h, float(tol_val), float(max_distance), C.byref(lp), C.byref(vp), C.byref(pp), 1 if drop else 0)
        self.eng._check(self.eng.lib.ml_result_segment_methylation(*args, C.byref(ns), *([None] * 13), C.byref(npm), *([None] * 11)))
        if reuse_pinned:
            seg = {k: self.eng.pinned_buffer("rseg_" + k, max(ns.value, 1), dt) for k, dt in self.REGION_SEG_COLUMNS}
            pmd = {k: self.eng.pinned_buffer("rpmd_" + k, max(npm.value, 1), dt) for k, dt in self.REGION_PMD_COLUMNS}
        else:
            seg = {k: np.empty(max(ns.value, 1), dt) for k, dt in self.REGION_SEG_COLUMNS}
            pmd = {k: np.empty(max(npm.value, 1), dt) for k, dt in self.REGION_PMD_COLUMNS}
        self.eng._check(self.eng.lib.ml_result_segment_methylation(
            *args, C.byref(ns), *[_p(seg[k]) for k, _ in self.REGION_SEG_COLUMNS], C.byref(npm),
            *[_p(pmd[k]) for k, _ in self.REGION_PMD_COLUMNS]))
        return ({k: v[:ns.value] for k, v in seg.items()}, {k: v[:npm.value] for k, v in pmd.items()})

    def call_pmd_segments(self, tol_val=0.01, max_distance=500.0, pmd_lambda2=1000.0, fetch=True, want_fused=False):
        """R/call.R:93-95 on the call table (std and width as the reference computes them)."""
        ns = C.c_int64(0)
        args = (self.eng.h, self.h, float(tol_val), float(max_distance), float(pmd_lambda2))
        self.eng._check(self.eng.lib.ml_result_call_pmd_segments(*args, C.byref(ns), *([None] * 9)))
        if not fetch:
            return ns.value
        n = max(ns.value, 1)
        out = dict(job=np.empty(n, np.int32), estimate_id=np.empty(n, np.int32), segment_id=np.empty(n, np.int32),
                   start=np.empty(n, np.uint32), end=np.empty(n, np.uint32), fused_std=np.empty(n),
                   std=np.empty(n), pseudoweight=np.empty(n))
        fused = np.empty(max(self.call_table(tol_val, max_distance, fetch=False), 1)) if want_fused else None
        self.eng._check(self.eng.lib.ml_result_call_pmd_segments(
            *args, C.byref(ns), _p(fused) if want_fused else None,
            *[_p(out[k]) for k in ("job", "estimate_id", "segment_id", "start", "end", "fused_std", "std",
                                   "pseudoweight")]))
        res = {k: v[:ns.value].copy() for k, v in out.items()}
        if want_fused:
            res["row_fused_std"] = fused[:self.call_table(tol_val, max_distance, fetch=False)]
        return res


class EnginePool:
    """Several engines (= CUDA streams) on ONE GPU, each driven by its own host thread.

    The C ABI calls are synchronous, so one engine alternates between PCIe transfers and kernels.
    Splitting the chromosomes of a call into groups and running the groups on different engines lets
    the H2D copy of one group, the kernels of another and the D2H copy of a third overlap (PCIe is
    full duplex and the copy engines are independent of the SMs).  ctypes releases the GIL during the
    native calls, so plain Python threads are enough.  Jobs never interact, so results do not depend
    on the grouping (see tests/test_gpu_detect.py::test_placement_independence)."""

    def __init__(self, n_engines=3, device=-1):
        import queue
        import threading
        self.engines = [Engine(device) for _ in range(n_engines)]
        # One transfer at a time per direction: if every group queued its H2D copies at once the copy
        # engine would interleave them and all groups would finish uploading together (no overlap
        # with compute); taking turns lets group 1 compute while group 2 uploads.
        self.h2d_turn = threading.Lock()
        self.d2h_turn = threading.Lock()
        # the host passes over a group's input columns (run scan, packing of the counts) use several threads each:
        # one group at a time, so that the first group is ready for the bus as early as possible
        self.pack_turn = threading.Lock()
        # one long-lived worker thread per engine (a fresh thread per call makes the CUDA runtime set up
        # per-thread state again and again; measured as a 100 ms hiccup every ~64 threads)
        self._tasks = [queue.Queue() for _ in range(n_engines)]
        self._done = queue.Queue()
        self._threads = [threading.Thread(target=self._worker, args=(k,), daemon=True) for k in range(n_engines)]
        for t in self._threads:
            t.start()

    def _worker(self, k):
        while True:
            task = self._tasks[k].get()
            if task is None:
                return
            fn, items, idx, out = task
            err = None
            try:
                for i in idx:
                    out[i] = fn(self.engines[k], items[i])
            except Exception as e:  # surfaced to the caller by map()
                err = e
            self._done.put(err)

    def close(self):
        for q in self._tasks:
            q.put(None)
        for t in self._threads:
            t.join()
        for e in self.engines:
            e.close()

    def map(self, fn, items):
        """Run fn(engine, item) for every item, items[i] on engine i % n, each engine on its own thread."""
        n = len(self.engines)
        out = [None] * len(items)
        used = min(n, len(items))
        for k in range(used):
            self._tasks[k].put((fn, items, range(k, len(items), n), out))
        errs = [self._done.get() for _ in range(used)]
        for e in errs:
            if e is not None:
                raise e
        return out


def split_jobs(job_ptr, cols, n_groups, edge_share=None):
    """Split a batch into n_groups contiguous-in-memory sub-batches, balanced longest-first by rows.
    edge_share=s (0 < s < 0.5, n_groups >= 3): the FIRST and the LAST group only get the smallest jobs, about a
    fraction s of the rows each, and the groups in between are balanced.  In an upload | compute | download
    pipeline (EnginePool) the first results are then ready to download early and little is left to compute and
    download after the last upload.
    Returns [(job_indices, sub_ptr, sub_cols)], with sub_cols freshly concatenated arrays."""
    from .shard import lpt_assign
    job_ptr = np.asarray(job_ptr, dtype=np.int64)
    sizes = np.diff(job_ptr)
    if not edge_share or n_groups < 3 or sizes.size < n_groups:
        group_of = lpt_assign(sizes, n_groups)
    else:
        order = [int(j) for j in np.argsort(sizes, kind="stable")]
        budget = float(edge_share) * float(sizes.sum())
        group_of = np.full(sizes.size, -1, np.int64)
        for edge in (0, n_groups - 1):
            got = 0.0
            while order and (got == 0.0 or got + sizes[order[0]] <= budget):
                j = order.pop(0)
                group_of[j] = edge
                got += float(sizes[j])
        rest = np.array(order, np.int64)
        group_of[rest] = 1 + lpt_assign(sizes[rest], n_groups - 2)
    out = []
    for g in range(n_groups):
        jobs = [int(j) for j in np.nonzero(group_of == g)[0]]
        if not jobs:
            continue
        ptr = np.zeros(len(jobs) + 1, np.int64)
        ptr[1:] = np.cumsum(sizes[jobs])
        sub = {k: np.concatenate([np.asarray(cols[k])[job_ptr[j]:job_ptr[j + 1]] for j in jobs]) for k in _COLUMNS}
        out.append((jobs, ptr, sub))
    return out


# ---------------------------------------------------------------------------------------------
# module-level mirror of the Rcpp module (one shared engine on the current device)
# ---------------------------------------------------------------------------------------------
_engine = None


def default_engine() -> Engine:
    global _engine
    if _engine is None:
        _engine = Engine()
    return _engine


def _columns(data):
    missing = [k for k in _COLUMNS if k not in data]
    if missing:  # src/signal_detection.cpp:13-15
        raise MethyLassoError(19, "data should contain columns dataset, condition, replicate, pos, "
                                  "Nmeth and coverage!")
    return {k: data[k] for k in _COLUMNS}


def _single(data, params, engine=None):
    eng = engine or default_engine()
    cols = _columns(data)
    n = len(cols["pos"])
    res = eng.detect_batch(np.array([0, n], np.int64), cols, params)
    try:
        return res.job(0)
    finally:
        res.free()


def signal_detection_fast_singlechr(data, lambda2=50, tol_val=0.01, max_iter=1, engine=None):
    """src/signal_detection.cpp:11-47 (module defaults src/MethyLasso_cpp.cpp:18-19)."""
    p = MlParams(lambda2, lambda2 + 1, 0.0, tol_val, 50.0, 1.0, max_iter, 1, 1, MODEL_FAST, 0, 0)
    r = _single(data, p, engine)
    return {"mu": r["mu"], "lambda2": r["lambda2"], "converged": r["converged"], "offset": r["offset"],
            "estimates": r["estimates"], "detection.type": "signal", "model": "normal",
            "_info": {k: r[k] for k in ("irls_steps", "dampings", "outer_iters", "mlogp")}}


def signal_detection_singlechr(data, optimize_lambda2=False, lambda2=50, max_lambda2=1000,
                               optimize_hyper=False, hyper=100, nouter=20, bf_per_kb=50, tol_val=0.01,
                               max_iter=30, engine=None):
    """src/signal_detection.cpp:49-80 (module defaults src/MethyLasso_cpp.cpp:13-16)."""
    p = MlParams(lambda2, max_lambda2, 0.0, tol_val, bf_per_kb, hyper, max_iter, nouter, 1,
                 MODEL_FULL_SIGNAL, int(bool(optimize_lambda2)), int(bool(optimize_hyper)))
    r = _single(data, p, engine)
    return {"mu": r["mu"], "lambda2": r["lambda2"], "hyper": np.array([r["hyper"]]),
            "converged": r["converged"], "offset": r["offset"], "estimates": r["estimates"],
            "detection.type": "signal", "model": "beta-binomial",
            "_info": {k: r[k] for k in ("irls_steps", "dampings", "outer_iters", "mlogp")}}


def difference_detection_singlechr(data, ref, optimize_lambda2=False, lambda2=50, max_lambda2=1000,
                                   optimize_hyper=False, hyper=100, nouter=20, bf_per_kb=50, tol_val=0.01,
                                   max_iter=30, engine=None):
    """src/difference_detection.cpp:13-45 (module defaults src/MethyLasso_cpp.cpp:23-26)."""
    p = MlParams(lambda2, max_lambda2, 0.0, tol_val, bf_per_kb, hyper, max_iter, nouter, int(ref),
                 MODEL_FULL_DIFFERENCE, int(bool(optimize_lambda2)), int(bool(optimize_hyper)))
    r = _single(data, p, engine)
    return {"mu": r["mu"], "lambda2": r["lambda2"], "hyper": np.array([r["hyper"]]),
            "converged": r["converged"], "offset": r["offset"], "estimates": r["estimates"],
            "ref.cond": int(ref), "detection.type": "difference", "model": "beta-binomial",
            "_info": {k: r[k] for k in ("irls_steps", "dampings", "outer_iters", "mlogp")}}


def segment_methylation_helper(df, tol_val=0.01, engine=None):
    """src/methylation_helpers.cpp:6-72 (module line src/MethyLasso_cpp.cpp:21)."""
    return (engine or default_engine()).segment_methylation_helper(df, tol_val)


def weighted_1d_flsa(y, wt, lambda2, lambda1=0, engine=None):
    """src/weighted_1d_flsa.cpp:7-21 (module line src/MethyLasso_cpp.cpp:28)."""
    return (engine or default_engine()).weighted_1d_flsa(y, wt, lambda2, lambda1)


def difference_detection_fast_singlechr(data, ref, lambda2=25, tol_val=0.01, max_iter=1, engine=None):
    """R/fast_diff.R:12-30: FAST fit of all conditions, then differences to condition 1 on the GPU."""
    eng = engine or default_engine()
    cols = _columns(data)
    p = MlParams(lambda2, lambda2 + 1, 0.0, tol_val, 50.0, 1.0, max_iter, 1, 1, MODEL_FAST, 0, 0)
    res = eng.detect_batch(np.array([0, len(cols["pos"])], np.int64), cols, p)
    try:
        res.sizes(0)
        if res.status[0] == 0:
            res.fast_diff_combine()
        r = res.job(0)
    finally:
        res.free()
    return {"mu": r["mu"], "lambda2": r["lambda2"], "converged": r["converged"], "offset": r["offset"],
            "estimates": r["estimates"], "detection.type": "difference", "model": "normal",
            "ref.cond": ref}


# ---- genome-wide wrappers (R/genome_wide.R): one batched native call over all chromosomes -------
def split_by_chr(data):
    """Rows grouped by chromosome in order of first appearance (data[,unique(chr)])."""
    chr_col = np.asarray(data["chr"])
    _, first, inv = np.unique(chr_col, return_index=True, return_inverse=True)
    order = np.argsort(first)                    # unique() keeps first-appearance order in R
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    job_of_row = rank[inv]
    perm = np.argsort(job_of_row, kind="stable")
    counts = np.bincount(job_of_row, minlength=order.size)
    ptr = np.zeros(order.size + 1, np.int64)
    np.cumsum(counts, out=ptr[1:])
    cols = {k: np.asarray(data[k])[perm] for k in _COLUMNS}
    names = chr_col[first[order]]
    return names, ptr, cols, perm


class ResidentFit:
    """Input batch and result of a genome-wide fit kept in HBM for the calls that follow it (call_differences)."""

    def __init__(self, eng, batch, res, names):
        self.eng, self.batch, self.res, self.names = eng, batch, res, names

    def free(self):
        if self.res is not None:
            self.res.free()
            self.batch.free()
            self.res = self.batch = None


def _genome_wide(data, params, detection_type, model, engine=None, combine_diff=False, ref=None,
                 verbose=False, keep_resident=False):
    eng = engine or default_engine()
    names, ptr, cols, perm = split_by_chr(data)
    batch = None
    if keep_resident:
        batch = eng.upload(ptr, cols)
        res = eng.detect(batch, params)
    else:
        res = eng.detect_batch(ptr, cols, params)
    try:
        if combine_diff:
            res.fast_diff_combine()
        ok = [j for j in range(res.njobs) if res.status[j] == 0]
        missing = [str(names[j]) for j in range(res.njobs) if res.status[j] != 0]
        if missing and verbose:  # foreach(.errorhandling="remove") + warning (R/genome_wide.R:51-52)
            print("Warning: the following chromosomes did not work:", " ".join(missing))
        jobs = [res.job(j) for j in ok]
    finally:
        if not keep_resident:
            res.free()
    # combine_ret (R/genome_wide.R:6-15)
    est = {k: np.concatenate([r["estimates"][k] for r in jobs]) if jobs else np.empty(0)
           for k in ("estimate_id", "start", "beta", "betahat", "pseudoweight")}
    est["chr"] = np.concatenate([np.repeat(names[j], jobs[i]["estimates"]["beta"].size)
                                 for i, j in enumerate(ok)]) if jobs else np.empty(0, dtype=object)
    ret = {"mu": np.concatenate([r["mu"] for r in jobs]) if jobs else np.empty(0),
           "lambda2": np.concatenate([r["lambda2"] for r in jobs]) if jobs else np.empty(0),
           "converged": np.array([r["converged"] for r in jobs], bool),
           "offset": np.concatenate([r["offset"] for r in jobs]) if jobs else np.empty(0),
           "estimates": est, "chr": np.array([names[j] for j in ok]),
           "detection.type": detection_type, "model": model}
    if model == "beta-binomial":
        ret["hyper"] = np.array([r["hyper"] for r in jobs])
    if ref is not None:
        ret["ref.cond"] = ref
    if keep_resident:
        ret["_resident"] = ResidentFit(eng, batch, res, names)
    return ret


def signal_detection_fast(data, lambda2=25, tol_val=0.01, max_iter=1, ncores=1, verbose=True, engine=None,
                          keep_resident=False):
    """R/genome_wide.R:64-78.  `ncores` is accepted and ignored: chromosomes are batched on the GPU.
    keep_resident=True leaves the input and the fit in HBM (ret["_resident"]) for segment_methylation."""
    p = MlParams(lambda2, lambda2 + 1, 0.0, tol_val, 50.0, 1.0, max_iter, 1, 1, MODEL_FAST, 0, 0)
    return _genome_wide(data, p, "signal", "normal", engine, verbose=verbose, keep_resident=keep_resident)


def segment_methylation(data, ret, ncores=1, min_num_cpgs=4, pmd_lambda2=1000, pmd_std_threshold=0.15,
                        umr_std_threshold=0.1, umr_max_beta=0.1, lmr_max_beta=0.5, pmd_max_beta=0.75, valley_max_beta=0.1,
                        pmd_valley_min_width=5000, max_distance=500, min_width=30, flank_width=300, flank_dist_max=5000,
                        flank_num_cpgs=10, flank_beta_min=0.5, umr_large_width=500, umr_large_density=0.03,
                        split_pmds=True, merge_pmds=True, drop=True, tol_val=0.01):
    """R/call.R:229-244 (segment_methylation) around :31-221 (segment_methylation_singlechr), for the result of
    signal_detection_fast(..., keep_resident=True) - or of difference_detection_fast(..., keep_resident=True), whose
    reference condition is then segmented (:40-44): the whole rule engine runs on the GPU on the resident fit, all
    chromosomes at once.  Returns dict(lmr_umr_valley, pmd) of column dicts with the reference's column order
    (:241-243; `num.cpgs` is `num_cpgs`, `PMD.status` is `PMD_status`: "PMD" / "other" / None); `ov`, the per-position
    join, is not returned.  `data` is only consulted through the resident batch; `ncores` is ignored."""
    if ret.get("detection.type") not in ("signal", "difference"):
        raise MethyLassoError(19, "segment_methylation needs the result of signal_detection_fast or difference_detection_fast")
    fit = ret.get("_resident")
    if fit is None or fit.res is None:
        raise MethyLassoError(19, "segment_methylation needs signal_detection_fast / difference_detection_fast(..., "
                                  "keep_resident=True): the rule engine works on the fit resident in HBM")
    # A difference fit: "will return the segmentation of the reference dataset" (R/call.R:40-44: estimate_id 1 of the
    # estimates, the rows of ret$ref.cond of the data).  The engine does the same on the resident fit: after the combine
    # the call table and the rule engine see the first estimator only, whose pooled counts the combine leaves untouched.
    res = fit.res
    lmr = dict(min_width=float(min_width), flank_width=float(flank_width), flank_dist_max=float(flank_dist_max),
               lmr_max_beta=float(lmr_max_beta), flank_beta_min=float(flank_beta_min), umr_max_beta=float(umr_max_beta),
               umr_std_threshold=float(umr_std_threshold), umr_large_width=float(umr_large_width),
               umr_large_density=float(umr_large_density), flank_num_cpgs=int(flank_num_cpgs))
    common = dict(valley_max_beta=valley_max_beta, pmd_valley_min_width=pmd_valley_min_width, min_num_cpgs=min_num_cpgs, **lmr)
    pkw = dict(pmd_lambda2=pmd_lambda2, pmd_std_threshold=pmd_std_threshold, pmd_max_beta=pmd_max_beta,
               split_pmds=split_pmds, merge_pmds=merge_pmds)
    seg_t, pmd_c = res.segment_methylation_tables(tol_val, max_distance, drop=drop, **pkw, **common)
    return region_tables(seg_t, pmd_c, np.asarray(fit.names))


STATUS_NAMES = np.array([None, "other", "PMD"], dtype=object)          # pmd_status + 1
CATEGORY_NAMES = np.array([None, "LMR", "UMR", "UMR-DMV"], dtype=object)


def region_tables(seg_t, pmd_c, names):
    """The two tables of ml_result_segment_methylation with the reference's columns and labels (R/call.R:241-243)."""
    seg = dict(condition=seg_t["condition"], chr=names[seg_t["job"]] if seg_t["job"].size else np.empty(0, dtype=object),
               start=seg_t["start"], end=seg_t["end"], width=seg_t["width"], segment_id=seg_t["segment_id"], beta=seg_t["beta"],
               coverage=seg_t["coverage"], pseudoweight=seg_t["pseudoweight"], num_cpgs=seg_t["num_cpgs"].astype(np.int64),
               std=seg_t["std"], category=CATEGORY_NAMES[seg_t["category"]], PMD_status=STATUS_NAMES[seg_t["pmd_status"] + 1])
    pmd_t = dict(condition=pmd_c["condition"], chr=names[pmd_c["job"]] if pmd_c["job"].size else np.empty(0, dtype=object),
                 start=pmd_c["start"], end=pmd_c["end"], width=pmd_c["width"], segment_id=pmd_c["segment_id"], beta=pmd_c["beta"],
                 std=pmd_c["std"], category=np.where(pmd_c["category"] == 1, "PMD", "other").astype(object),
                 num_cpgs=pmd_c["num_cpgs"], fused_std=pmd_c["fused_std"])
    return dict(lmr_umr_valley=seg, pmd=pmd_t)

def signal_detection(data, optimize_lambda2=False, lambda2=25, max_lambda2=1000, optimize_hyper=False,
                     hyper=100, nouter=20, bf_per_kb=50, tol_val=0.01, max_iter=30, ncores=1, engine=None):
    """R/genome_wide.R:38-54."""
    p = MlParams(lambda2, max_lambda2, 0.0, tol_val, bf_per_kb, hyper, max_iter, nouter, 1,
                 MODEL_FULL_SIGNAL, int(bool(optimize_lambda2)), int(bool(optimize_hyper)))
    return _genome_wide(data, p, "signal", "beta-binomial", engine, verbose=True)


def difference_detection(data, ref, optimize_lambda2=False, lambda2=25, max_lambda2=1000,
                         optimize_hyper=False, hyper=100, nouter=20, bf_per_kb=50, tol_val=0.01,
                         max_iter=30, ncores=1, engine=None):
    """R/genome_wide.R:91-107."""
    p = MlParams(lambda2, max_lambda2, 0.0, tol_val, bf_per_kb, hyper, max_iter, nouter, int(ref),
                 MODEL_FULL_DIFFERENCE, int(bool(optimize_lambda2)), int(bool(optimize_hyper)))
    return _genome_wide(data, p, "difference", "beta-binomial", engine, ref=int(ref), verbose=True)


def difference_detection_fast(data, ref, lambda2=25, tol_val=0.01, max_iter=1, ncores=1, verbose=True,
                              engine=None, keep_resident=False):
    """R/genome_wide.R:121-135 + R/fast_diff.R:12-30.  keep_resident=True leaves the input and the fit in HBM
    (ret["_resident"]) for call_differences; release them with ret["_resident"].free()."""
    p = MlParams(lambda2, lambda2 + 1, 0.0, tol_val, 50.0, 1.0, max_iter, 1, 1, MODEL_FAST, 0, 0)
    return _genome_wide(data, p, "difference", "normal", engine, combine_diff=True, ref=ref, verbose=verbose,
                        keep_resident=keep_resident)


def call_differences(data, ret, min_diff=0.1, pval_cutoff=0.05, fdr_cutoff=1, cov_score=70, min_num_cpgs=4,
                     tol_val=0.01, ncores=1, verbose=True, min_delta_diff=0.1, max_distance=500.0, return_all=False):
    """R/call.R:348-365: DMR calls from the result of difference_detection_fast(..., keep_resident=True).
    Per chromosome call_differences_singlechr (:254-333) runs on the GPU for all chromosomes at once; then the FDR
    over the genome (:360, on the GPU) and the threshold filter (:363).  `ret["ref.cond"]` is the 1-based code of the
    reference condition.  Returns the table of :361-363 as a dict of arrays (all rows with return_all=True, plus
    the boolean column `called`)."""
    if ret.get("detection.type") != "difference":           # R/call.R:256
        raise MethyLassoError(19, "Please call call_differences on the result of difference_detection")
    fit = ret.get("_resident")
    if fit is None or fit.res is None:
        raise MethyLassoError(19, "call_differences needs difference_detection_fast(..., keep_resident=True): the "
                                  "DMR caller works on the fit and the input rows resident in HBM")
    t = fit.res.call_differences(fit.batch, int(ret["ref.cond"]), min_diff, min_delta_diff, tol_val, max_distance)
    t = dict(t)
    t["chr"] = np.asarray(fit.names)[t["job"]] if t["job"].size else np.empty(0, dtype=object)
    t["ref"] = np.full(t["job"].size, ret["ref.cond"])
    t["fdr"] = fit.eng.p_adjust_fdr(t["pval"]) if t["pval"].size else np.empty(0)
    with np.errstate(invalid="ignore"):
        called = (~np.isnan(t["diff"]) & (np.abs(t["diff"]) >= min_diff) & (t["coverage_score"] >= cov_score)
                  & (t["pval"] <= pval_cutoff) & (t["fdr"] <= fdr_cutoff)
                  & (np.minimum(t["num_cpgs"], t["num_cpgs_ref"]) >= min_num_cpgs))
    if return_all:
        t["called"] = called
        return t
    return {k: v[called] for k, v in t.items()}
