#!/usr/bin/env python
"""bench.py — CpGs/s through IRLS + weighted 1-D FLSA (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cpgs M]

Workload (config.workload): BASELINE.json configs[1] — a full human-genome-shaped synthetic WGBS
sample set (~28 M CpGs over 24 chromosomes, one condition, 3 replicates, coverage ~ Poisson(20)
filtered at >= 5), UMR/LMR segmentation at lambda2 = 25 followed by the PMD std fusion at
lambda2 = 1000.  One STEP = the native calls the reference's R layer makes for that config:
    signal_detection_fast (24 x signal_detection_fast_singlechr, MethyLasso.R:559)
    segment_methylation_helper on the estimates            (R/call.R:53)
    the per-segment call table: gap split, shrink, mean/sd (R/call.R:55-72; data.table code in the reference,
                                                            native here so that the chain stays on the device)
    weighted_1d_flsa(std, width, 1000) per chromosome      (R/call.R:93)
    segment_methylation_helper on the fused table          (R/call.R:95)
`value` times the step with inputs already resident in HBM (validation, indexing, IRLS, FLSA,
segmentation all inside); `e2e` times the same step through the host-buffer C ABI from pinned host
memory, including the H2D of the six input columns and the D2H of mu, the estimates table and both
segment tables.  With N > 1 every rank processes its own genome (independent units, no data-path
collective; weak scaling) and the ranks all-gather their segment-table sizes over NCCL at the end.

Timing hygiene: `value` is one contiguous loop of K steps between two CUDA events, no per-kernel instrumentation
(the `kernels` table comes from a second pass over the same K steps).  Clocks / throttle reasons are read through NVML
right before and right after that loop and, polled every 5 ms, during one more untimed repetition of the step: any
driver query issued while the steps run can stall them (measured).  The boxes are shared and an outside stall of
0.1-1.5 s hits about one step in a hundred; a loop in which one step took more than 1.3 x the loop's median step
(undisturbed loops stay within 1.15 x) is listed under `disturbed_attempts` and the loop is timed again (at most twice).

--impl reference times the reference's CPU path for the same workload on the host cores: the oracle's
plain-C restatement of methylation_detection with the reference's own tf_dp_weight (oracle/_ref,
compiled verbatim) as the DP, one chromosome per thread (mirrors foreach %dopar% / -t).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from methylasso_b200 import synth  # noqa: E402

LAMBDA2, PMD_LAMBDA2, TOL, MAX_DISTANCE = 25.0, 1000.0, 0.01, 500.0
DISTURBED = 1.3     # a step slower than this multiple of its loop's median step marks the loop as disturbed from outside


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons, read through NVML (pynvml: nvmlDeviceGetClockInfo,
    nvmlDeviceGetCurrentClocksEventReasons - the fields of the recipe's nvidia-smi clocks line).
    Anything that queries the driver while the steps run disturbs them on these boxes: a separate `nvidia-smi -lms 100`
    process gave steps of 338 and 452 ms among 15.3 ms ones, an NVML thread polling every 10 ms steps of 34 ms, and
    even one query between two steps was now and then followed by a step of 82 or 661 ms; without any query 60 steps
    in a row stayed within 15.2-15.9 ms (one of 19.6).  So the timed loop is bracketed, not interleaved: one sample
    right before its first step and one right after its last kernel has been enqueued and timed (the GPU has been
    busy for the whole loop microseconds earlier), and then ONE MORE, UNTIMED repetition of the same step runs with
    a thread polling NVML every 5 ms - the clocks under exactly this load."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []          # (sm_mhz, reasons bitmask)
        self.query_s = 0.0
        self.max_mhz = None
        self.nv = self.h = self.get_reasons = None

    def start(self):
        if os.environ.get("ML_BENCH_NO_SAMPLER"):      # debugging aid
            return
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.idx
            self.h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nv = nv
        except Exception:
            self.nv = None

    def mark(self):
        """Start of the timed region: samples taken before are dropped."""
        self.samples.clear()
        self.query_s = 0.0

    def sample(self):
        if self.nv is None:
            return
        t0 = time.perf_counter()
        try:
            self.samples.append((float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)), int(self.get_reasons(self.h))))
        except Exception:
            pass
        self.query_s += time.perf_counter() - t0

    def sample_during(self, fn, period_s=0.005):
        """run fn() once (untimed) while a thread polls NVML"""
        if self.nv is None:
            fn()
            return
        stop = threading.Event()

        def poll():
            while not stop.is_set():
                self.sample()
                time.sleep(period_s)
        th = threading.Thread(target=poll, daemon=True)
        th.start()
        try:
            fn()
        finally:
            stop.set()
            th.join(timeout=2)

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"], "samples": 0}
        mask = 0
        for _, r in self.samples:
            mask |= r
        return {"sm_mhz": float(np.median([s[0] for s in self.samples])) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(n for b, n in self.REASONS.items() if mask & b),
                "samples": len(self.samples), "query_ms_total": round(self.query_s * 1e3, 3),
                "how": "NVML: one sample right before and one right after the timed loop, then polled every 5 ms "
                       "during one more, untimed repetition of the same step"}


def bind_to_gpu_cpus(gpu_index):
    """One process per GPU: run this rank's threads on the CPUs NVML reports as local to its GPU, so that the pinned
    host buffers of the e2e step (allocated afterwards: first touch) sit in the memory of the GPU's own NUMA node and
    the ranks do not all cross the socket interconnect for their 3 GB per step.  Best effort; returns what was done."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu_index
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": len(cpus), "first": min(cpus), "last": max(cpus)}
    except Exception as e:           # no NVML, no permission: run unbound
        return {"error": str(e)[:80]}
    return None


def make_workload(total_cpgs, seed_offset=0):
    tables = synth.genome_tables(2, total_cpgs=total_cpgs, reps_per_condition=(3,),
                                 seed=synth.BASE_SEED + 2 + 1000 * seed_offset)
    ptr, cols = synth.concat_jobs(tables)
    return tables, ptr, cols


# ------------------------------------------------------------------------------------------------
# reference arm: oracle port + the reference's verbatim DP, all host threads
# ------------------------------------------------------------------------------------------------
def reference_step(tables, O, threads):
    """One pass of the same native-call sequence on the CPU; returns unique CpGs processed."""
    out = [None] * len(tables)

    def work(ids):
        for j in ids:
            t = tables[j]
            r = O.detect(t, O.MODEL_FAST, lambda2=LAMBDA2, tol_val=TOL, max_iter=1)
            est = r["estimates"]
            seg = O.segment_methylation_helper(est, TOL)
            # R/call.R:43-72: call table from the pooled per-position data (one condition: the estimates table
            # holds exactly that: pooled coverage = pseudoweight, pooled Nmeth = betahat * pseudoweight)
            pooled = dict(estimate_id=est["estimate_id"], pos=est["start"].astype(np.int32),
                          Nmeth=np.rint(est["betahat"] * est["pseudoweight"]).astype(np.int32),
                          coverage=est["pseudoweight"].astype(np.int32))
            call = O.call_table(pooled, seg, MAX_DISTANCE)
            fused = O.weighted_1d_flsa(call["std"], call["width"], PMD_LAMBDA2)              # R/call.R:93
            pmd = O.segment_methylation_helper(dict(estimate_id=call["estimate_id"], start=call["start"].astype(np.int32),
                                                    beta=fused, betahat=call["std"],
                                                    pseudoweight=call["pseudoweight"]), TOL)  # R/call.R:95
            out[j] = (r["n_params"], seg["start"].size, pmd["start"].size)
    order = sorted(range(len(tables)), key=lambda j: -tables[j]["pos"].size)  # longest first
    buckets = [[] for _ in range(threads)]
    load = [0] * threads
    for j in order:
        k = int(np.argmin(load))
        buckets[k].append(j)
        load[k] += tables[j]["pos"].size
    ths = [threading.Thread(target=work, args=(b,)) for b in buckets if b]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    return sum(o[0] for o in out)


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    kind_dp = "reference tf_dp_weight (oracle/_ref)" if O.use_reference_dp(True) else "oracle DP restatement"
    threads = os.cpu_count() or 1
    tables, ptr, cols = make_workload(args.cpgs)
    U = None
    for _ in range(args.warmup):
        U = reference_step(tables, O, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        U = reference_step(tables, O, threads)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = U / dt
    # For information: the reference's OWN C++ path (unmodified sources compiled against the stand-in Rcpp/Eigen
    # headers, oracle/_ref/libmethylasso_ref.so) on a bounded sample, one thread.  It is much slower than the
    # restatement (std::map binner, eager stand-in Eigen), so the restatement stays the conservative baseline.
    ref_cxx = None
    if O.ref_full_lib() is not None:
        small = synth.chromosome_table(100_000, 4242, (3,))
        t1 = time.perf_counter()
        rr = O.reference_detect(small, model=0, lambda2=LAMBDA2, tol_val=TOL, max_iter=1)
        O.reference_segment_helper(rr["estimates"], TOL)
        d1 = time.perf_counter() - t1
        t1 = time.perf_counter()
        O.detect(small, O.MODEL_FAST, lambda2=LAMBDA2, tol_val=TOL, max_iter=1)
        d2 = time.perf_counter() - t1
        u = int(np.unique(small["pos"]).size)
        ref_cxx = {"value": u / d1, "unit": "CpG/s", "cores": 1, "port_same_sample": u / d2,
                   "sample": "one 100k-CpG x 3-replicate chromosome, FAST detect + segment helper, 1 thread; "
                             "reference sources + stand-in Rcpp/Eigen headers (oracle/stubs)"}
    line = {
        "impl": "reference", "metric": "CpGs/sec through IRLS+weighted 1D FLSA", "value": value, "unit": "CpG/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, tables, ptr, U),
        "cpu_baseline": {"value": value, "unit": "CpG/s", "cores": threads, "kind": "port",
                         "sample": f"whole workload per step; oracle C restatement of methylation_detection, DP = {kind_dp}, "
                                   f"one chromosome per thread over {threads} threads"},
        "e2e": {"value": value, "unit": "CpG/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ref_cxx": ref_cxx,
    }
    print(json.dumps(line), flush=True)
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(line) + "\n")


def workload_config(args, tables, ptr, U=None):
    return {"workload": "configs[1]: full human-genome synthetic, one condition, 3 replicates, UMR/LMR (lambda2=25) "
                        "+ PMD std fusion (lambda2=1000)",
            "cpgs_per_genome": int(U) if U is not None else int(args.cpgs),
            "rows_per_genome": int(ptr[-1]), "chromosomes": len(tables), "replicates": 3, "lambda2": LAMBDA2,
            "pmd_lambda2": PMD_LAMBDA2, "tol_val": TOL, "model": "FAST (normal/identity, max_iter=1)",
            "l2_policy": "inputs (2.0 GB per genome) and parameter arrays are larger than the 126 MB L2",
            "pmd_input": "std and width of the native call table (R/call.R:55-72), max_distance = 500"}


# ------------------------------------------------------------------------------------------------
# algorithmic bytes per launch: N rows, EP = E*P parameters, D datasets
# ------------------------------------------------------------------------------------------------
def kernel_bytes(N, EP, D, words):
    return {
        "k_validate": 24 * N,
        "k_dset_bounds": 0,
        "k_bitmap_mark": 4 * N + 8 * words,
        "k_bitmap_tile_count": 4 * words,
        "k_bitmap_word_prefix": 8 * words,
        "k_rank_unique": 16 * N + 4 * D * EP + 8 * EP,
        "k_pseudodata": 4 * D * EP + 8 * N + 32 * EP,
        "k_offset_partial": 8 * N,
        "flsa_warm_kernel": 16 * EP,   # y, w of the warm-up stretch (warm == chunk: one extra read of each)
        "flsa_main_kernel": 32 * EP,   # y, w read; tm, tp written
        "flsa_bwd_reduce_kernel": 16 * EP,
        "flsa_bwd_apply_kernel": 32 * EP,
        "k_center": 16 * EP,
        "k_eval_rows": 22 * N,          # two launches per step: the initial evaluation reads nmeth, cov (8 N: every parameter is
                                        # still zero, no index, no gather, no mu), the one after the update 36 N; average
        "k_tv_partial": 4 * EP,         # likewise: the initial launch writes zeros, the second reads beta (8 EP)
        "k_seg_run_heads": 12 * EP,
        "k_seg_emit": 28 * EP,          # beta, betahat, pw, start once (two launches: the second one is on the call table)
        "k_call_count": 16 * EP,        # pos, pw
        "k_call_parts": 16 * EP,        # pos, pw
        "k_call_emit_part": 24 * EP,    # pos, betahat, pw once (the second and third sweeps hit L1 / L2)
        "k_scan_tile_sums": 4 * EP,
        "k_scan_apply": 12 * EP,
    }


# ------------------------------------------------------------------------------------------------
# auxiliary line: configs[2], two-condition DMR calling (3 vs 3 replicates) - not the contract line
# ------------------------------------------------------------------------------------------------
def run_dmr(args, local):
    """difference_detection_fast (MethyLasso.R:453) + call_differences (:463) on a 3-vs-3 synthetic genome: resident
    step (detect, fast-diff combine, DMR caller, FDR) timed with CUDA events, per-kernel times, and the CPU oracle
    (C restatement + Python restatement of the R lines) on one small chromosome beside it."""
    import torch
    from methylasso_b200 import api
    from methylasso_b200._native import MlParams
    torch.cuda.set_device(local)
    eng = api.Engine(local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    tables = synth.genome_tables(3, total_cpgs=args.cpgs, reps_per_condition=(3, 3))
    ptr, cols = synth.concat_jobs(tables)
    N = int(ptr[-1])
    params = MlParams(LAMBDA2, LAMBDA2 + 1, 0.0, TOL, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
    batch = eng.upload(ptr, cols)

    def step():
        res = eng.detect(batch, params)
        res.fast_diff_combine()
        t = res.call_differences(batch, 2, 0.1, 0.1, TOL, MAX_DISTANCE, reuse_pinned=True)
        fdr = eng.p_adjust_fdr(t["pval"])
        return res, t, fdr

    res, t, fdr = step()
    U = int(sum(res.sizes(j)[1] for j in range(res.njobs)))
    res.free()
    eng.set("profile", 1)
    for _ in range(max(3, args.warmup)):
        r, t, fdr = step()
        r.free()
    eng.set("pool_headroom_percent", 25)     # no cuMemMap inside a timed step
    eng.profile_reset()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    step_wall = []
    for _ in range(args.steps):
        tw0 = time.perf_counter()
        r, t, fdr = step()
        r.free()
        step_wall.append(round((time.perf_counter() - tw0) * 1e3, 2))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    prof = eng.profile()
    kern = sorted(((k, v[0] / args.steps, v[1] / args.steps) for k, v in prof.items()), key=lambda x: -x[1])
    called = (~np.isnan(t["diff"]) & (np.abs(t["diff"]) >= 0.1) & (t["coverage_score"] >= 70) & (t["pval"] <= 0.05)
              & (np.minimum(t["num_cpgs"], t["num_cpgs_ref"]) >= 4))
    line = {"metric": "CpGs/sec through difference_detection_fast + call_differences (aux, configs[2])",
            "value": U / (ms * 1e-3), "unit": "CpG/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": ms,
            "config": {"workload": "configs[2]: two-condition DMR calling, 3 vs 3 replicates genome-wide",
                       "cpgs_per_genome": U, "rows_per_genome": N, "chromosomes": len(tables)},
            "regions": int(t["start"].size), "dmrs_called": int(called.sum()), "host_wall_ms_each_step": step_wall,
            "gpu_launches": int(eng.launch_count - l0),
            "kernels": [{"kernel": k, "ms_per_step": round(a, 4), "launches_per_step": b} for k, a, b in kern[:24]]}
    # ---- e2e: host buffers in, region table out (what r/R/genome_wide_gpu.R::dmr_calling_gpu does) ------------------
    # chromosomes in G groups, one engine + host thread per group: pinned H2D of the six... four moved input columns,
    # detect, fast-diff combine, call_differences with its table downloaded; then one FDR over all regions.
    if not args.no_e2e:
        G = max(1, args.e2e_groups)
        pool = api.EnginePool(G, local)
        groups = []
        for jobs, gptr, gcols in api.split_jobs(ptr, cols, G):
            pin_in = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in gcols.items()}
            groups.append((jobs, gptr, pin_in))

        def run_group(eng_g, grp):
            jobs, gptr, pin_in = grp
            with pool.h2d_turn:
                bt = eng_g.upload(gptr, pin_in)
            rr = eng_g.detect(bt, params)
            rr.fast_diff_combine()
            tt = rr.call_differences(bt, 2, 0.1, 0.1, TOL, MAX_DISTANCE, reuse_pinned=True)   # engine-owned buffers
            rr.free()
            bt.free()
            return tt

        def step_e2e():
            outs = pool.map(run_group, groups)
            pv = np.concatenate([o["pval"] for o in outs])
            return outs, pool.engines[0].p_adjust_fdr(pv)
        for _ in range(3):
            outs, q = step_e2e()
        for e_g in pool.engines:
            e_g.set("pool_headroom_percent", 25)
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        x0.record(stream)
        e_wall = []
        for _ in range(args.steps):
            tw0 = time.perf_counter()
            outs, q = step_e2e()
            e_wall.append(round((time.perf_counter() - tw0) * 1e3, 1))
        torch.cuda.synchronize()
        x1.record(stream)
        torch.cuda.synchronize()
        ems = x0.elapsed_time(x1) / args.steps
        d2h = int(sum(v.nbytes for o in outs for v in o.values()) + q.nbytes)
        line["e2e"] = {"value": U / (ems * 1e-3), "unit": "CpG/s", "ms_per_step": ems, "engines": G,
                       "h2d_bytes_per_step": 24 * N, "d2h_bytes_per_step": d2h, "host_wall_ms_each_step": e_wall,
                       "regions": int(sum(o["start"].size for o in outs)),
                       "api": "per chromosome group: ml_batch_upload + ml_batch_detect + ml_result_fast_diff_combine + "
                              "ml_result_call_differences (table downloaded); then ml_p_adjust_fdr over all regions"}
        pool.close()
    if not args.no_cpu_baseline:
        from oracle import oracle as O, diffcall as DC
        O.build()
        # conservative CPU arm for this config: the FIT ALONE (oracle C port + the reference's verbatim DP) of four
        # chromosomes, one per host thread - without any of the region calling the GPU line includes
        samp = tables[:4]
        Us = [0] * len(samp)

        def fit(i):
            o_ = O.detect(samp[i], O.MODEL_FAST, lambda2=LAMBDA2, tol_val=TOL, max_iter=1)
            O.fast_diff_combine(o_["estimates"], o_["n_params"], o_["n_est"])
            Us[i] = int(o_["n_params"])
        tc0 = time.perf_counter()
        ths = [threading.Thread(target=fit, args=(i,)) for i in range(len(samp))]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        tc1 = time.perf_counter()
        line["cpu_fit_only"] = {"value": sum(Us) / (tc1 - tc0), "unit": "CpG/s", "cores": len(samp), "kind": "port",
                                "sample": "difference_detection_fast of the 4 largest chromosomes (%d CpGs, 3 vs 3), one per "
                                          "thread, %.1f s; no region calling" % (sum(Us), tc1 - tc0)}
        small = synth.chromosome_table(100_000, 77, (3, 3))
        t0 = time.perf_counter()
        o = O.detect(small, O.MODEL_FAST, lambda2=LAMBDA2, tol_val=TOL, max_iter=1)
        est = O.fast_diff_combine(o["estimates"], o["n_params"], o["n_est"])
        t1 = time.perf_counter()
        DC.call_differences_singlechr(small, est, 2)
        t2 = time.perf_counter()
        line["cpu_baseline"] = {"value": o["n_params"] / (t2 - t0), "unit": "CpG/s", "cores": 1, "kind": "port",
                                "sample": "one 100 k-CpG chromosome, 3 vs 3; detect %.2f s (C port), call_differences "
                                          "%.2f s (Python restatement of the R lines: not a timing baseline)" % (t1 - t0, t2 - t1)}
    print(json.dumps(line), flush=True)
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(line) + "\n")
    batch.free()
    eng.close()


# ------------------------------------------------------------------------------------------------
# auxiliary line: ONE genome sharded over the ranks, longest chromosome first (north star, point 3)
# ------------------------------------------------------------------------------------------------
def run_strong(args, rank, world, local):
    """Strong scaling of configs[1]: the 24 chromosomes of one genome are assigned to the ranks longest-first
    (methylasso_b200.shard.lpt_assign), every rank runs its shard resident (detect, segments, call table, PMD fusion)
    with no data-path collective, and the per-shard segment and PMD tables are all-gathered over NCCL at the end
    (shard.all_gather_tables) - inside the timed region.  Rank 0 checks the gathered tables against its own
    single-GPU run of the whole genome (bit-equal: a job gives the same bits wherever it is placed)."""
    import torch
    import torch.distributed as dist
    from methylasso_b200 import api, shard
    from methylasso_b200._native import MlParams
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = api.Engine(local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    params = MlParams(LAMBDA2, LAMBDA2 + 1, 0.0, TOL, 50.0, 1.0, 1, 1, 1, 0, 0, 0)
    tables, ptr, cols = make_workload(args.cpgs, seed_offset=0)          # the SAME genome on every rank
    sizes = np.diff(ptr)
    rank_of = shard.lpt_assign([shard.job_cost(n) for n in sizes], world)
    mine = shard.my_jobs(rank_of, rank)
    mptr = np.zeros(len(mine) + 1, np.int64)
    mptr[1:] = np.cumsum(sizes[mine])
    mcols = {k: np.ascontiguousarray(np.concatenate([cols[k][ptr[j]:ptr[j + 1]] for j in mine])) for k in cols}
    batch = eng.upload(mptr, mcols)
    job_ids = np.asarray(mine, np.int32)

    t_parts = [0.0, 0.0]           # host wall of the shard's own work / of the exchange, accumulated over the timed steps

    def step():
        t0 = time.perf_counter()
        res = eng.detect(batch, params)
        seg = res.segments(TOL)
        pmd = res.call_pmd_segments(TOL, MAX_DISTANCE, PMD_LAMBDA2)
        res.free()
        for t in (seg, pmd):
            t["job"] = job_ids[t["job"]]                                  # shard-local job index -> chromosome
        t1 = time.perf_counter()
        # the exchange: every rank ends up with every shard's tables in its HBM; rank 0 also copies them to its host
        gs = shard.all_gather_packed(seg, dist, device=f"cuda:{local}") if world > 1 else None
        gp = shard.all_gather_packed(pmd, dist, device=f"cuda:{local}") if world > 1 else None
        if world > 1 and rank == 0:
            gs = ([p.cpu() for p in gs[0]],) + gs[1:]
            gp = ([p.cpu() for p in gp[0]],) + gp[1:]
        t_parts[0] += t1 - t0
        t_parts[1] += time.perf_counter() - t1
        return seg, pmd, gs, gp

    def tables_of(seg, pmd, gs, gp):      # untimed: the gathered matrices as tables in chromosome order
        if world > 1:
            seg, pmd = shard.unpack_gathered(*gs), shard.unpack_gathered(*gp)
        return shard.reorder_by_job(seg), shard.reorder_by_job(pmd)

    for _ in range(max(3, args.warmup)):
        out = step()
    t_parts[0] = t_parts[1] = 0.0
    eng.set("pool_headroom_percent", 25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    wall = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = step()
        wall.append(round((time.perf_counter() - t0) * 1e3, 2))
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        seg, pmd = tables_of(*out)
        same = None
        if world > 1:                      # the whole genome on this GPU alone must give the same tables
            full = eng.upload(ptr, cols)
            res = eng.detect(full, params)
            s1, p1 = res.segments(TOL), res.call_pmd_segments(TOL, MAX_DISTANCE, PMD_LAMBDA2)
            res.free()
            full.free()
            same = all(np.array_equal(seg[k], s1[k], equal_nan=True) for k in s1) and \
                all(np.array_equal(pmd[k], p1[k], equal_nan=True) for k in p1)
        U = int(sum(np.unique(t["pos"]).size for t in tables))
        line = {"metric": "CpGs/sec through IRLS+weighted 1D FLSA, ONE genome sharded over the ranks (aux)",
                "value": U / (float(ms.item()) * 1e-3), "unit": "CpG/s", "n_gpus": world, "steps": args.steps,
                "ms_per_step": float(ms.item()), "scaling": "strong", "host_wall_ms_each_step_rank0": wall,
                "jobs_per_rank": np.bincount(rank_of, minlength=world).tolist(),
                "rows_per_rank": np.bincount(rank_of, weights=sizes, minlength=world).astype(np.int64).tolist(),
                "segments": int(seg["start"].size), "pmd_segments": int(pmd["start"].size),
                "gathered_tables_equal_single_gpu": same,
                "rank0_ms_per_step_shard_work": t_parts[0] * 1e3 / args.steps,
                "rank0_ms_per_step_exchange_and_merge": t_parts[1] * 1e3 / args.steps,
                "timed": "detect + segments + call table + PMD fusion of the rank's chromosomes, fetch of its segment and "
                         "PMD tables, NCCL all-gather of both (one packed matrix per table: every rank holds all shards "
                         "in HBM afterwards), copy of the gathered matrices to rank 0's host; max over ranks"}
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(line) + "\n")
    batch.free()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strong", action="store_true",
                    help="auxiliary measurement: one genome sharded over the ranks (strong scaling) instead of the contract line")
    ap.add_argument("--dmr", action="store_true",
                    help="auxiliary measurement: configs[2] (3 vs 3 DMR calling) instead of the contract line")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpgs", type=int, default=synth.GENOME_CPGS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-groups", type=int, default=6,
                    help="engines (streams + host threads) the e2e step spreads the chromosomes over; 1 = single call")
    ap.add_argument("--resident-groups", type=int, default=0,
                    help="> 1: also time the resident step with the chromosomes spread over this many engines")
    ap.add_argument("--e2e-edge", type=float, default=0.0,
                    help="share of the rows given to the first and to the last group of the e2e pipeline (0 = balanced)")
    ap.add_argument("--e2e-no-mu", action="store_true",
                    help="auxiliary: the e2e step does not download mu (the reference returns it, but neither MethyLasso.R "
                         "nor R/call.R ever reads ret$mu); 39 %% fewer bytes come back")
    ap.add_argument("--out", default=None, help="also append the JSON line to this file")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed resident steps (ncu --profile-from-start off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.dmr:
        if rank == 0:
            run_dmr(args, local)
        return
    if args.strong:
        run_strong(args, rank, world, local)
        return

    import torch
    import torch.distributed as dist
    from methylasso_b200 import api
    from methylasso_b200._native import MlParams

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: methylasso_b200 has no CPU path")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_cpus(local) if world > 1 else None      # before any pinned buffer is allocated (first touch)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    eng = api.Engine(local)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    params = MlParams(LAMBDA2, LAMBDA2 + 1, 0.0, TOL, 50.0, 1.0, 1, 1, 1, 0, 0, 0)

    # started seconds before anything is timed (NVML initialisation, or nvidia-smi's start-up in the fall-back, can
    # stall CUDA calls for 50-350 ms)
    sampler = ClockSampler(local)
    sampler.start()
    tables, ptr, cols = make_workload(args.cpgs, seed_offset=rank)
    N = int(ptr[-1])
    # pinned host copies of the inputs (e2e) and of the outputs
    pin = {k: torch.from_numpy(v).pin_memory() for k, v in cols.items()}
    pcols = {k: v.numpy() for k, v in pin.items()}
    batch = eng.upload(ptr, pcols)

    def step_resident():
        res = eng.detect(batch, params)
        nseg = res.segments_count(TOL)
        res.call_table(TOL, MAX_DISTANCE, fetch=False)
        npmd = res.call_pmd_segments(TOL, MAX_DISTANCE, PMD_LAMBDA2, fetch=False)
        return res, nseg, npmd

    # sizes (one untimed call)
    res, nseg, npmd = step_resident()
    sizes = [res.sizes(j) for j in range(res.njobs)]
    assert all(s[4] == 0 for s in sizes), "a synthetic chromosome failed validation"
    U = int(sum(s[1] for s in sizes))
    EP = int(sum(s[1] * s[2] for s in sizes))
    D = 3
    words = int(sum((int(t["pos"].max()) - int(t["pos"].min())) // 32 + 1 for t in tables))
    res.free()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: resident inputs ------------------------------------------------------------------
    lw0 = eng.launch_count
    for _ in range(args.warmup):
        r, _, _ = step_resident()
        r.free()
    per_step = (eng.launch_count - lw0) // max(1, args.warmup)
    eng.set("pool_headroom_percent", 25)     # no cuMemMap inside a timed step
    gc.collect()
    gc.disable()
    barrier()
    # The GPU boxes are shared: now and then something outside this process (another tenant's driver queries take a
    # node-wide lock) stalls one step for 0.1-1.5 s (seen with no sampler of our own running: steps of 969 and 1535 ms
    # among 15 ms ones).  A K-step loop in which a step took more than DISTURBED x the loop's median step is recorded under
    # `disturbed_attempts` and the whole loop is timed again, at most twice; the reported numbers are those of ONE
    # contiguous K-step loop.
    disturbed = []
    for attempt in range(3):
        sampler.mark()
        sampler.sample()
        l0 = eng.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.profiler_range and attempt == 0:
            torch.cuda.cudart().cudaProfilerStart()
        ev0.record(stream)
        step_wall_res = []
        for _ in range(args.steps):
            tw0 = time.perf_counter()
            r, nseg, npmd = step_resident()
            r.free()
            step_wall_res.append(round((time.perf_counter() - tw0) * 1e3, 2))
        ev1.record(stream)
        sampler.sample()
        if args.profiler_range and attempt == 0:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        barrier()
        launches = eng.launch_count - l0
        ms = ev0.elapsed_time(ev1) / args.steps
        bad = torch.tensor([1.0 if max(step_wall_res) > DISTURBED * float(np.median(step_wall_res)) else 0.0], device="cuda")
        if world > 1:
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)       # every rank repeats or none does
        if bad.item() == 0.0 or attempt == 2 or args.profiler_range:
            break
        disturbed.append({"ms_per_step": ms, "host_wall_ms_each_step": step_wall_res})

    def one_more():
        r_, _, _ = step_resident()
        r_.free()
    sampler.sample_during(one_more)
    clocks = sampler.stop()
    # the same K steps once more with a pair of CUDA events around every kernel (the per-kernel durations of the
    # roofline and of `kernels`); kept out of `value`: ~800 event records cost a step about 0.4 ms
    eng.set("profile", 1)
    r, _, _ = step_resident()                # fills the engine's pool of timing events
    r.free()
    eng.set("profile_reserve_events", int(2.2 * per_step * (args.steps + 1)))   # no cudaEventCreate while timing
    eng.profile_reset()
    barrier()
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pv0.record(stream)
    for _ in range(args.steps):
        r, _, _ = step_resident()
        r.free()
    pv1.record(stream)
    barrier()
    gc.enable()
    ms_profiled = pv0.elapsed_time(pv1) / args.steps
    prof = eng.profile()
    prof_work = eng.profile_work()
    eng.set("profile", 0)
    t_ms = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    u_all = torch.tensor([U], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(u_all, op=dist.ReduceOp.SUM)
    value = float(u_all.item()) / (ms_max * 1e-3)

    # ---- resident inputs, chromosomes spread over several engines (streams + host threads) ---------
    # The latency-bound phases of one group (FLSA left-neighbour rounds, PMD chains, host round trips) overlap
    # with the bandwidth-bound kernels of the others.
    resident_pool = None
    if args.resident_groups > 1:
        Gr = args.resident_groups
        rpool = api.EnginePool(Gr, local)
        rgroups = []
        for k, (jobs, gptr, gcols) in enumerate(api.split_jobs(ptr, cols, Gr)):
            rgroups.append((rpool.engines[k % Gr].upload(gptr, gcols), jobs))

        def run_group_resident(eng_g, grp):
            bt, jobs = grp
            rr = eng_g.detect(bt, params)
            ns = rr.segments_count(TOL)
            rr.call_table(TOL, MAX_DISTANCE, fetch=False)
            npm = rr.call_pmd_segments(TOL, MAX_DISTANCE, PMD_LAMBDA2, fetch=False)
            rr.free()
            eng_g.synchronize()
            return ns, npm
        for _ in range(max(3, args.warmup)):
            outs_r = rpool.map(run_group_resident, rgroups)
        for e_g in rpool.engines:
            e_g.set("pool_headroom_percent", 10)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lr0 = sum(e_g.launch_count for e_g in rpool.engines)
        r0.record(stream)
        for _ in range(args.steps):
            outs_r = rpool.map(run_group_resident, rgroups)
        torch.cuda.synchronize()
        r1.record(stream)
        barrier()
        rms = r0.elapsed_time(r1) / args.steps
        t3 = torch.tensor([rms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        resident_pool = {"engines": Gr, "ms_per_step": float(t3.item()), "value": float(u_all.item()) / (float(t3.item()) * 1e-3),
                         "unit": "CpG/s", "gpu_launches": int(sum(e_g.launch_count for e_g in rpool.engines) - lr0),
                         "segments": [int(sum(o[0] for o in outs_r)), int(sum(o[1] for o in outs_r))]}
        for bt, _ in rgroups:
            bt.free()
        rpool.close()

    # ---- e2e: host buffers through the C ABI -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # The user-level call: chromosomes split into groups, one engine (stream) + host thread per group,
        # so H2D of one group, kernels of another and D2H of a third overlap.  Inputs and outputs are
        # pinned host buffers; sizes are known from the untimed call above.
        G = max(1, args.e2e_groups)
        pool = api.EnginePool(G, local)
        groups = []
        for jobs, gptr, gcols in api.split_jobs(ptr, cols, G, edge_share=(args.e2e_edge or None)):
            pin_in = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in gcols.items()}
            n_g = int(gptr[-1])
            ep_g = int(sum(sizes[j][1] * sizes[j][2] for j in jobs))
            outb = dict(mu=torch.empty(n_g, dtype=torch.float64).pin_memory().numpy(),
                        off=torch.empty(3 * len(jobs), dtype=torch.float64).pin_memory().numpy(),   # pageable would make the
                        # asynchronous copy call block until the detect kernels are done
                        est_id=torch.empty(ep_g, dtype=torch.int32).pin_memory().numpy(),
                        **{k: torch.empty(ep_g, dtype=torch.float64).pin_memory().numpy()
                           for k in ("start", "beta", "betahat", "pseudoweight")})
            groups.append((jobs, gptr, pin_in, outb))

        def run_group(eng_g, grp):
            jobs, gptr, pin_in, o = grp
            ts = [time.perf_counter()]
            with pool.h2d_turn:
                ts.append(time.perf_counter())
                bt = eng_g.upload(gptr, pin_in)
            ts.append(time.perf_counter())
            rr = eng_g.detect(bt, params)
            bt.free()
            ts.append(time.perf_counter())
            # D2H of mu / estimates is enqueued on the copy stream and overlaps the segmentation kernels
            rr.copy_all_async(None if args.e2e_no_mu else o["mu"], o["off"], o["est_id"], o["start"], o["beta"],
                              o["betahat"], o["pseudoweight"])
            ts.append(time.perf_counter())
            rr.segments_count(TOL)                      # resident; the tables are fetched after the big copies
            rr.call_table(TOL, MAX_DISTANCE, fetch=False)
            rr.call_pmd_segments(TOL, MAX_DISTANCE, PMD_LAMBDA2, fetch=False)
            ts.append(time.perf_counter())
            eng_g.synchronize()
            seg = rr.segments(TOL, reuse_pinned=True)               # the engine's own page-locked buffers
            call = rr.call_table(TOL, MAX_DISTANCE, reuse_pinned=True)
            pmd = rr.call_pmd_segments(TOL, MAX_DISTANCE, PMD_LAMBDA2)
            pmd["call_table"] = call
            rr.free()
            ts.append(time.perf_counter())
            timeline.append((len(jobs), [round((x - t_step0[0]) * 1e3, 1) for x in ts]))
            return seg, pmd

        timeline, t_step0 = [], [0.0]

        def step_e2e():
            timeline.clear()
            t_step0[0] = time.perf_counter()
            return pool.map(run_group, groups)
        for _ in range(max(3, args.warmup)):
            outs = step_e2e()
        for e_g in pool.engines:          # pre-grow each engine's pool a little: no cuMemMap inside a timed step
            e_g.set("pool_headroom_percent", 25)
        gc.collect()
        gc.disable()                      # no collector pause inside a 50 ms step
        barrier()
        e2e_sampler = sampler
        e2e_disturbed = []
        for attempt in range(3):             # same policy as for the resident loop
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2e_sampler.mark()
            e2e_sampler.sample()
            t0 = time.perf_counter()
            e0.record(stream)
            step_wall = []
            for _ in range(args.steps):
                ts0 = time.perf_counter()
                outs = step_e2e()
                step_wall.append(round((time.perf_counter() - ts0) * 1e3, 1))
            torch.cuda.synchronize()
            e1.record(stream)
            wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            barrier()
            bad = torch.tensor([1.0 if max(step_wall) > DISTURBED * float(np.median(step_wall)) else 0.0], device="cuda")
            if world > 1:
                dist.all_reduce(bad, op=dist.ReduceOp.MAX)
            if bad.item() == 0.0 or attempt == 2:
                break
            e2e_disturbed.append({"ms_per_step": e0.elapsed_time(e1) / args.steps, "host_wall_ms_each_step": step_wall})
        timeline_timed = sorted(timeline, key=lambda t: t[1][0])
        e2e_sampler.sample()
        e2e_sampler.sample_during(step_e2e)
        barrier()
        gc.enable()
        ems = e0.elapsed_time(e1) / args.steps
        t2 = torch.tensor([ems], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        seg_bytes = sum(v.nbytes for sg, pm in outs for v in list(sg.values()) + [x for x in pm.values() if not isinstance(x, dict)]
                        + list(pm["call_table"].values()))
        e2e = {"value": float(u_all.item()) / (float(t2.item()) * 1e-3), "unit": "CpG/s",
               "ms_per_step": float(t2.item()), "host_wall_ms_per_step": wall_ms, "h2d_bytes_per_step": 24 * N,
               "d2h_bytes_per_step": int((0 if args.e2e_no_mu else 8 * N) + 36 * EP + seg_bytes), "engines": G,
               "mu_downloaded": not args.e2e_no_mu,
               "host_wall_ms_each_step": step_wall, "disturbed_attempts": e2e_disturbed, "clocks": e2e_sampler.stop(),
               "timeline_ms_last_step": timeline_timed,  # per group: start, h2d start/end, detect end, d2h enqueued, segments+pmd end, all done
               "api": "EnginePool.map over chromosome groups: ml_batch_upload + ml_batch_detect + ml_result_copy_all + "
                      "ml_result_segments + ml_result_pmd_segments per group, transfers taking turns per direction"}
        pool.close()

    # ---- K16: all-gather of per-shard segment summaries over NCCL ---------------------------------
    gathered = None
    if world > 1:
        mine = torch.tensor([U, nseg, npmd], device="cuda", dtype=torch.int64)
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        gathered = [v.tolist() for v in allv]

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    pk, pk_kind = peaks()
    kb = kernel_bytes(N, EP, D, words)
    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full`
    # capture of this same command (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py)
    traffic, ncu_detail = {}, {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, ncu_detail = tj.get("dram_bytes_per_launch", {}), tj.get("detail", {})
    table = []
    for name, (tot_ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        per = tot_ms / max(1, n)
        # data-dependent kernels (FLSA chunk kernels) report the bytes of the DP steps they completed
        b = (prof_work[name] / max(1, n)) if name in prof_work else kb.get(name)
        gbs = (b / (per * 1e-3) / 1e9) if b and per > 0 else None
        table.append({"kernel": name, "ms_per_step": tot_ms / args.steps, "launches_per_step": n / args.steps,
                      "avg_launch_ms": per, "alg_bytes_per_launch": b, "GBps": gbs,
                      "frac_of_hbm": (gbs / pk["hbm_gbs"]) if gbs else None,
                      "ncu_dram_bytes_per_launch": traffic.get(name)})
    top = table[0] if table else None
    roofline = None
    if top:
        roofline = {"kernel": top["kernel"], "bound": "hbm", "achieved": top["GBps"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": top["frac_of_hbm"], "traffic": top["ncu_dram_bytes_per_launch"], "peak_source": pk_kind,
                    "alg_bytes_per_launch": top["alg_bytes_per_launch"], "avg_launch_ms": top["avg_launch_ms"],
                    # what bounds a latency-bound kernel: DP steps per second from this run, issue
                    # and occupancy figures from the committed ncu capture
                    "latency_bound": ({"dp_steps_per_s": (top["alg_bytes_per_launch"] / (16.0 if top["kernel"] == "flsa_warm_kernel" else 32.0))
                                                         / (top["avg_launch_ms"] * 1e-3),
                                       "ncu_issue_active_pct": ncu_detail.get(top["kernel"], {}).get("mean_issue_active_pct"),
                                       "ncu_warps_active_pct": ncu_detail.get(top["kernel"], {}).get("mean_warps_active_pct"),
                                       "ncu_fp64_pipe_pct": ncu_detail.get(top["kernel"], {}).get("mean_fp64_pipe_pct")}
                                      if top["kernel"].startswith("flsa_") and top["alg_bytes_per_launch"] else None),
                    "note": "FLSA chunk kernels run a dependent fp64 compare/add/divide chain per element (latency-bound, "
                            "not HBM-bound): their HBM fraction is reported as measured; the bandwidth-shaped kernels "
                            "of the step are listed in `kernels` with their own fractions"}

    # the largest kernel of the step that IS bandwidth-shaped (elementwise / gather / reduction sweeps of the IRLS loop)
    latency_bound = ("flsa_warm_kernel", "flsa_main_kernel", "flsa_chain_kernel", "flsa_pass_c_kernel")
    top_bw = next((t for t in table if t["kernel"] not in latency_bound and t["GBps"]), None)
    roofline_bw = None
    if top_bw:
        roofline_bw = {"kernel": top_bw["kernel"], "bound": "hbm", "achieved": top_bw["GBps"], "peak": pk["hbm_gbs"],
                       "unit": "GB/s", "frac": top_bw["frac_of_hbm"], "traffic": top_bw["ncu_dram_bytes_per_launch"],
                       "peak_source": pk_kind, "alg_bytes_per_launch": top_bw["alg_bytes_per_launch"],
                       "avg_launch_ms": top_bw["avg_launch_ms"]}

    # ---- auxiliary, untimed part of no metric: the rule engine of segment_methylation_singlechr (R/call.R:75-215) on the
    # resident fit of this workload - LMR / UMR annotation, valleys, PMD split / calls, PMD status.  Every repetition uses
    # a new key (flank_dist_max 5000 / 5000.5: distances are integers, same results) so that nothing is reused.
    rule_engine = None
    if rank == 0:
        try:
            res_r, _, _ = step_resident()
            eng.synchronize()
            walls, fin, pmd = [], None, None
            for rep in range(4):
                t0 = time.perf_counter()
                kw = dict(flank_dist_max=5000.0 + 0.5 * (rep & 1))
                pmd = res_r.call_pmds(TOL, MAX_DISTANCE, PMD_LAMBDA2, **kw)
                fin = res_r.call_pmd_status(TOL, MAX_DISTANCE, PMD_LAMBDA2, **kw)
                eng.synchronize()
                walls.append((time.perf_counter() - t0) * 1e3)
            res_r.free()
            rule_engine = {"host_wall_ms": float(np.median(walls[1:])), "host_wall_ms_each": [round(w, 2) for w in walls],
                           "rows": int(fin["category"].size), "lmr": int((fin["category"] == 1).sum()),
                           "umr": int((fin["category"] == 2).sum()), "valleys": int((fin["category"] == 3).sum()),
                           "pmds": int((pmd["category"] == 1).sum()),
                           "what": "R/call.R:75-215 on the resident call table, result tables downloaded; not part of the "
                                   "timed step (the reference arm does not run it)"}
        except Exception as exc:                                   # never let the auxiliary line break the bench
            rule_engine = {"error": repr(exc)}

    # ---- CPU baseline on rank 0 (bounded sample) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        dp = "reference tf_dp_weight (oracle/_ref)" if O.use_reference_dp(True) else "oracle DP restatement"
        t0 = time.perf_counter()
        Us = reference_step(tables, O, 1)        # the whole workload once, on one host thread (~12 s)
        dt = time.perf_counter() - t0
        cpu = {"value": Us / dt, "unit": "CpG/s", "cores": 1, "kind": "port",
               "sample": f"one pass over the whole workload ({Us} CpGs, {dt:.1f} s), single thread; oracle C restatement "
                         f"of methylation_detection, DP = {dp}"}

    if rank == 0:
        line = {
            "metric": "CpGs/sec through IRLS+weighted 1D FLSA", "value": value, "unit": "CpG/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, tables, ptr, U), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "roofline_largest_bandwidth_kernel": roofline_bw, "cpu_baseline": cpu,
            "kernels": table[:16], "resident_pipelined": resident_pool, "ms_per_step_with_kernel_events": ms_profiled,
            "kernel_ms_per_step_total": sum(t["ms_per_step"] for t in table),
            "host_wall_ms_each_step": step_wall_res, "disturbed_attempts": disturbed, "cpu_binding": numa,
            "segments": {"umr_lmr_segments": int(nseg), "pmd_segments": int(npmd)}, "gathered": gathered,
            "rule_engine": rule_engine,
        }
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(line) + "\n")
    batch.free()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
