"""Genome-wide calls from HOST buffers through several engines of one GPU.

What `MethyLasso.R` does for one condition (Part 1, MethyLasso.R:559-592) is
    ret <- signal_detection_fast(data, lambda2 = 25)       # R/genome_wide.R:64-78: one native call per chromosome
    seg <- segment_methylation(data, ret, ...)             # R/call.R:229-244: the rule engine, per chromosome
and for two conditions (Part 2, :453-469)
    ret <- difference_detection_fast(data, ref, ...)       # R/genome_wide.R:121-135, R/fast_diff.R:12-30
    dmr <- call_differences(data, ret, ...)                # R/call.R:348-365
Here the chromosomes of such a call are split into groups; every group has an engine (a CUDA stream and a memory
pool) and a host thread, so that the upload of one group, the kernels of another and the download of a third overlap
(`api.EnginePool`).  Jobs never interact, so the result does not depend on the grouping.

The class below is the one code path behind `bench.py`'s end-to-end numbers AND `tests/test_gpu_pipeline.py`, which
checks its outputs against single-engine calls and against the CPU restatement of the reference.
"""
from __future__ import annotations

import contextlib
import os
import threading
import time

import numpy as np

from . import api
from ._native import MlParams

_INPUT = ("dataset", "condition", "replicate", "pos", "Nmeth", "coverage")


def _pin(arr, lib):
    out = api.pinned_empty(arr.size, arr.dtype, lib)
    out[:] = arr
    return out


class _Turns:
    """The groups of a call take a resource (the packing threads, the bus) in the order of their index: the engine of group k
    is served before those of the groups in front of it (its stream has the higher priority), because after the last upload
    everybody waits for the last group's chain of kernels."""

    def __init__(self):
        self.cv = threading.Condition()
        self.next = 0

    def reset(self):
        with self.cv:
            self.next = 0
            self.cv.notify_all()

    def abort(self):
        with self.cv:
            self.next = -1
            self.cv.notify_all()

    @contextlib.contextmanager
    def turn(self, i):
        with self.cv:
            while self.next != i and self.next >= 0:
                self.cv.wait()
            if self.next < 0:
                raise RuntimeError("another group of this call failed")
        try:
            yield
        finally:
            with self.cv:
                if self.next >= 0:
                    self.next = i + 1
                self.cv.notify_all()


class GenomePipeline:
    """The jobs (chromosomes) of one data set, split into groups for `n_engines` engines of one GPU.

    job_ptr / cols: the batch as `synth.concat_jobs` or `api.split_by_chr` give it.  pin=True copies every group's
    input columns into page-locked memory once (the host buffers the calls below read from)."""

    def __init__(self, job_ptr, cols, n_engines=6, device=-1, lambda2=25.0, tol_val=0.01, edge_share=None, pin=True,
                 host_threads=None):
        self.pool = api.EnginePool(max(1, int(n_engines)), device)
        self.lib = self.pool.engines[0].lib
        self._pack, self._bus = _Turns(), _Turns()
        if len(self.pool.engines) > 1 and os.environ.get("ML_PIPE_PRIORITY", "1") != "0":
            for k, e in enumerate(self.pool.engines):      # group k enters after group k - 1: served first
                e.set("stream_priority", k)
        self.params = MlParams(float(lambda2), float(lambda2) + 1, 0.0, float(tol_val), 50.0, 1.0, 1, 1, 1, api.MODEL_FAST, 0, 0)
        self.tol_val = float(tol_val)
        self.pin = bool(pin)
        self._pinned = []               # per group: column name -> page-locked buffer (kept across rebind)
        self._edge_share = edge_share
        self.rebind(job_ptr, cols)
        # packing threads: all CPUs of this process but four (the engines' own host threads and the driver's need some; measured
        # on a 16-CPU box: 12 threads 25.4 ms per genome, 16 threads 28 ms with outliers, 8 threads 28 ms)
        ht = int(host_threads or os.environ.get("ML_HOST_THREADS") or max(4, min(32, len(os.sched_getaffinity(0)) - 4)))
        for e in self.pool.engines:
            e.set("host_threads", ht)
        self.host_threads = ht
        self.timeline = []
        self._t0 = 0.0

    def rebind(self, job_ptr, cols, reuse_pinned=True):
        """Another data set of the same kind through the same engines (a cohort, sample after sample): the groups are cut
        anew and copied into the page-locked buffers of the previous ones where they fit (reuse_pinned=False: into
        buffers of their own, so that `bound()` / `bind()` can switch between prepared data sets without copying)."""
        if not reuse_pinned:
            self._pinned = []
        self.njobs = int(np.asarray(job_ptr).size - 1)
        self.groups = []
        self._bufs = None
        parts = api.split_jobs(job_ptr, cols, len(self.pool.engines), edge_share=self._edge_share)
        for g, (jobs, gptr, gcols) in enumerate(parts):
            gcols = {k: np.ascontiguousarray(gcols[k], np.int32) for k in _INPUT}
            if self.pin:
                if g >= len(self._pinned):
                    self._pinned.append({})
                held = self._pinned[g]
                for k, v in gcols.items():
                    if k not in held or held[k].size < v.size:
                        held[k] = api.pinned_empty(int(v.size * 1.1) + 64, v.dtype, self.lib)
                    held[k][:v.size] = v
                    gcols[k] = held[k][:v.size]
            self.groups.append((np.asarray(jobs, np.int32), np.ascontiguousarray(gptr, np.int64), gcols))

    def bound(self):
        return self.njobs, self.groups

    def bind(self, prepared):
        self.njobs, self.groups = prepared
        self._bufs = None

    # ---- bytes that cross PCIe per call ---------------------------------------------------------------------------
    def h2d_bytes(self):
        """pos (4 bytes per row) + Nmeth / coverage as packed by ml_batch_stage (2 x 1 bytes per row when every count of
        the group is below 256, 2 x 2 below 65536, else 2 x 4): the three factor columns are reduced to a run table on
        the host and never travel."""
        total = 0
        for _, gptr, gcols in self.groups:
            top = max(int(gcols["Nmeth"].max(initial=0)), int(gcols["coverage"].max(initial=0)))
            low = min(int(gcols["Nmeth"].min(initial=0)), int(gcols["coverage"].min(initial=0)))
            per = 1 if (top < 256 and low >= 0) else 2 if (top < 65536 and low >= 0) else 4
            total += (4 + 2 * per) * int(gptr[-1])
        return int(total)

    def _upload(self, eng, gptr, gcols, t):
        """ml_batch_stage (host passes, one group at a time on all the packing threads) then ml_batch_commit (the
        transfers, one group at a time on the bus); the packing of a group overlaps the transfer of the one before.  The
        groups take their turns in index order."""
        i = self._group_index[id(gptr)]
        try:
            with self._pack.turn(i):
                t.append(time.perf_counter())
                bt = eng.stage(gptr, gcols)
            with self._bus.turn(i):
                t.append(time.perf_counter())
                eng.commit(bt)
        except BaseException:
            self._pack.abort()
            self._bus.abort()
            raise
        t.append(time.perf_counter())
        return bt

    def close(self):
        self.pool.close()

    def warm_pools(self, percent=25):
        for e in self.pool.engines:
            e.set("pool_headroom_percent", int(percent))

    def _stamp(self, marks):
        self.timeline.append([round((t - self._t0) * 1e3, 2) for t in marks])

    def _run(self, fn, items=None):
        self.timeline = []
        self._group_index = {id(g[1]): i for i, g in enumerate(self.groups)}
        self._pack.reset()
        self._bus.reset()
        self._t0 = time.perf_counter()
        return self.pool.map(fn, self.groups if items is None else items)

    # ---- Part 1 of the CLI: region tables only --------------------------------------------------------------------
    def segment_methylation(self, max_distance=500.0, drop=True, **rule_kw):
        """signal_detection_fast + segment_methylation for all chromosomes: host columns in, the two region tables
        out (`lmr_umr_valley`, `pmd` as `api.Result.segment_methylation_tables` returns them, `job` = index of the
        chromosome in the batch, rows ordered by chromosome).  Per group: ml_batch_upload, ml_batch_detect (mu not
        stored: nothing in the reference's R layer reads it), ml_result_segment_methylation."""
        pool = self.pool

        def run_group(eng, grp):
            jobs, gptr, gcols = grp
            t = [time.perf_counter()]
            eng.set("want_mu", 0)
            bt = self._upload(eng, gptr, gcols, t)
            res = eng.detect(bt, self.params)
            t.append(time.perf_counter())
            seg, pmd = res.segment_methylation_tables(self.tol_val, max_distance, drop=drop, reuse_pinned=True, **rule_kw)
            status = res.status.copy()
            res.free()
            bt.free()
            t.append(time.perf_counter())
            self._stamp(t)
            return jobs, seg, pmd, status

        outs = self._run(run_group)
        return self._merge_regions(outs)

    def _merge_regions(self, outs):
        segs, pmds = [], []
        status = np.zeros(self.njobs, np.int32)
        for jobs, seg, pmd, st in outs:
            status[jobs] = st
            segs.append({k: (jobs[v] if k == "job" else v) for k, v in seg.items()})
            pmds.append({k: (jobs[v] if k == "job" else v) for k, v in pmd.items()})
        out = []
        for parts in (segs, pmds):
            cat = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
            order = np.argsort(cat["job"], kind="stable")
            out.append({k: v[order] for k, v in cat.items()})
        return out[0], out[1], status

    # ---- the reference's own return values: per-CpG tables too ------------------------------------------------------
    def signal_detection_fast(self, max_distance=500.0, pmd_lambda2=1000.0, want_mu=True, buffers=None):
        """signal_detection_fast with everything the reference's list holds (mu, offset, the estimates table) downloaded
        into page-locked buffers, plus the segment table, the call table and the PMD candidate segments of every
        chromosome (R/call.R:53-95).  Returns one dict per group: jobs, mu, offset, estimates, segments, call_table,
        pmd_segments.  `buffers` (from `output_buffers`) are reused from call to call."""
        pool = self.pool
        bufs = buffers or self.output_buffers(want_mu)

        def run_group(eng, item):
            grp, o = item
            jobs, gptr, gcols = grp
            t = [time.perf_counter()]
            eng.set("want_mu", 1 if want_mu else 0)
            bt = self._upload(eng, gptr, gcols, t)
            res = eng.detect(bt, self.params)
            bt.free()
            t.append(time.perf_counter())
            # D2H of mu / estimates is enqueued on the copy stream and overlaps the segmentation kernels
            res.copy_all_async(o["mu"] if want_mu else None, o["offset"], o["estimate_id"], o["start"], o["beta"], o["betahat"],
                               o["pseudoweight"])
            res.segments_count(self.tol_val)                      # resident; the tables are fetched after the big copies
            res.call_table(self.tol_val, max_distance, fetch=False)
            res.call_pmd_segments(self.tol_val, max_distance, pmd_lambda2, fetch=False)
            t.append(time.perf_counter())
            eng.synchronize()
            seg = res.segments(self.tol_val, reuse_pinned=True)    # the engine's own page-locked buffers
            call = res.call_table(self.tol_val, max_distance, reuse_pinned=True)
            pmd = res.call_pmd_segments(self.tol_val, max_distance, pmd_lambda2)
            status = res.status.copy()
            res.free()
            t.append(time.perf_counter())
            self._stamp(t)
            est = {k: o[k] for k in ("estimate_id", "start", "beta", "betahat", "pseudoweight")}
            return dict(jobs=jobs, mu=o["mu"] if want_mu else None, offset=o["offset"], estimates=est, segments=seg,
                        call_table=call, pmd_segments=pmd, status=status)

        return self._run(run_group, list(zip(self.groups, bufs)))

    def output_buffers(self, want_mu=True):
        """Page-locked output buffers of signal_detection_fast, sized by one untimed fit per group."""
        if getattr(self, "_bufs", None) is not None and self._bufs_mu == want_mu:
            return self._bufs

        def size_group(eng, grp):
            jobs, gptr, gcols = grp
            res = eng.detect_batch(gptr, gcols, self.params)
            sizes = [res.sizes(j) for j in range(res.njobs)]
            res.free()
            ep = int(sum(s[1] * s[2] for s in sizes if s[4] == 0))
            nd = int(sum(s[3] for s in sizes if s[4] == 0))
            nr = int(sum(s[0] for s in sizes if s[4] == 0))
            o = dict(offset=api.pinned_empty(max(nd, 1), np.float64, self.lib),
                     estimate_id=api.pinned_empty(max(ep, 1), np.int32, self.lib),
                     **{k: api.pinned_empty(max(ep, 1), np.float64, self.lib) for k in ("start", "beta", "betahat", "pseudoweight")})
            o["mu"] = api.pinned_empty(max(nr, 1), np.float64, self.lib) if want_mu else None
            o["n"] = (nr, ep, nd)
            return o
        self._bufs = self.pool.map(size_group, self.groups)
        self._bufs_mu = want_mu
        return self._bufs

    @staticmethod
    def d2h_bytes_regions(seg, pmd):
        return int(sum(v.nbytes for v in seg.values()) + sum(v.nbytes for v in pmd.values()))

    @staticmethod
    def d2h_bytes_full(outs):
        total = 0
        for o in outs:
            nr, ep, nd = 0, 0, 0
            total += (o["mu"].nbytes if o["mu"] is not None else 0) + o["offset"].nbytes
            total += sum(v.nbytes for v in o["estimates"].values())
            total += sum(v.nbytes for k in ("segments", "call_table", "pmd_segments") for v in o[k].values())
        return int(total)
