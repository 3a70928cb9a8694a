"""CPU tests of the multi-GPU host logic (gloo, world size 2): LPT sharding, the all-gather of
per-shard segment tables, and that the sharded result equals the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest

from methylasso_b200 import shard, synth


def test_lpt_balance_and_determinism():
    sizes = [n * 3 for n in synth.chromosome_cpgs()]
    costs = [shard.job_cost(n) for n in sizes]
    for world in (1, 2, 4, 8):
        r = shard.lpt_assign(costs, world)
        assert np.array_equal(r, shard.lpt_assign(costs, world))
        loads = np.bincount(r, weights=costs, minlength=world)
        assert loads.max() / loads.mean() < 1.08            # 24 chromosomes over 8 ranks: ~97 % balance
        assert sorted(sum((shard.my_jobs(r, k) for k in range(world)), [])) == list(range(24))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tables = [synth.chromosome_table(n, 300 + i, (2,)) for i, n in enumerate((4000, 900, 2500, 300, 1700))]
    rank_of = shard.lpt_assign([shard.job_cost(t["pos"].size) for t in tables], world)
    segs = {k: [] for k in ("job", "estimate_id", "segment_id", "start", "end", "beta")}
    for j in shard.my_jobs(rank_of, rank):           # the compute stands in for the GPU engine here
        r = O.detect(tables[j], O.MODEL_FAST, lambda2=25.0)
        s = O.segment_methylation_helper(r["estimates"], 0.01)
        segs["job"].append(np.full(s["start"].size, j, np.int32))
        for k in ("estimate_id", "segment_id", "start", "end", "beta"):
            segs[k].append(s[k])
    local = {k: (np.concatenate(v) if v else np.empty(0, np.float64 if k == "beta" else np.int32))
             for k, v in segs.items()}
    gathered, lens = shard.all_gather_tables(local, dist)
    packed, lens2 = shard.all_gather_tables_packed(local, dist)      # one collective per table: same result
    assert lens2 == lens and all(np.array_equal(packed[k], gathered[k]) and packed[k].dtype == gathered[k].dtype
                                 for k in gathered)
    gathered = shard.reorder_by_job(gathered)
    if rank == 0:
        q.put((gathered, lens, rank_of.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_equals_single_process(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, lens, rank_of = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(lens) == gathered["job"].size and set(rank_of) == {0, 1}
    tables = [synth.chromosome_table(n, 300 + i, (2,)) for i, n in enumerate((4000, 900, 2500, 300, 1700))]
    at = 0
    for j, t in enumerate(tables):
        r = oracle.detect(t, 0, lambda2=25.0)
        s = oracle.segment_methylation_helper(r["estimates"], 0.01)
        n = s["start"].size
        assert (gathered["job"][at:at + n] == j).all()
        for k in ("estimate_id", "segment_id", "start", "end", "beta"):
            assert np.array_equal(gathered[k][at:at + n], s[k]), (j, k)
        at += n
    assert at == gathered["job"].size


# ---- the same exchange for the tables of the rule engine (R/call.R:31-221): lmr_umr_valley and pmd per shard ------------------
RULE_SIZES = (30_000, 6_000, 14_000)
RULE_KEYS = {"lmr_umr_valley": ("condition", "start", "end", "segment_id", "width", "beta", "coverage", "pseudoweight",
                                "num_cpgs", "std", "category", "pmd_status"),
             "pmd": ("condition", "category", "start", "end", "width", "num_cpgs", "fused_std", "beta", "std")}


def _rule_tables(O, t, j):
    from oracle import rule_engine as RE
    from test_calltable import pooled_from_table
    r = O.detect(t, O.MODEL_FAST, lambda2=25.0)
    out = RE.segment_methylation_singlechr(pooled_from_table(t), r["estimates"], pmd_lambda2=100.0, pmd_valley_min_width=1500.0,
                                           pmd_std_threshold=0.1)["all"]
    res = {}
    for name, keys in RULE_KEYS.items():
        tab = {k: (np.asarray(out[name][k], np.int32) if np.asarray(out[name][k]).dtype.kind in "iu" else np.asarray(out[name][k], np.float64))
               for k in keys}
        tab["job"] = np.full(tab["start"].size, j, np.int32)
        res[name] = tab
    return res


def _rule_worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import oracle as O
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tables = [synth.chromosome_table(n, 400 + i, (2,)) for i, n in enumerate(RULE_SIZES)]
    rank_of = shard.lpt_assign([shard.job_cost(t["pos"].size) for t in tables], world)
    mine = [_rule_tables(O, tables[j], j) for j in shard.my_jobs(rank_of, rank)]    # stands in for the GPU engine
    gathered = {}
    for name, keys in RULE_KEYS.items():
        cols = keys + ("job",)
        local = {k: (np.concatenate([m[name][k] for m in mine]) if mine else
                     np.empty(0, np.int32 if k in ("job", "condition", "category", "pmd_status", "segment_id", "num_cpgs") else np.float64))
                 for k in cols}
        tab, _ = shard.all_gather_tables_packed(local, dist)           # NaN (NA std / beta of a valley) travels as it is
        gathered[name] = shard.reorder_by_job(tab)
    if rank == 0:
        q.put((gathered, rank_of.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_rule_engine_tables(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rule_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, rank_of = q.get(timeout=90)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert set(rank_of) == {0, 1}
    tables = [synth.chromosome_table(n, 400 + i, (2,)) for i, n in enumerate(RULE_SIZES)]
    want = [_rule_tables(oracle, t, j) for j, t in enumerate(tables)]
    for name, keys in RULE_KEYS.items():
        for k in keys + ("job",):
            w = np.concatenate([x[name][k] for x in want])
            assert np.array_equal(gathered[name][k], w, equal_nan=True) and gathered[name][k].dtype == w.dtype, (name, k)
    assert gathered["pmd"]["start"].size > 10 and (gathered["lmr_umr_valley"]["category"] > 0).sum() > 30


def test_gather_layout_matches_the_python_unpack():
    """ml_gather_layout (the slot layout of the native K16 exchange, no GPU needed): 64-byte header, the 13 + 11 columns of
    the two region tables back to back, each 16-byte aligned and sized by its dtype; RegionGather.fetch slices the
    fetched bytes with exactly these offsets."""
    import ctypes as C
    from methylasso_b200 import _native
    from methylasso_b200.api import Result
    lib = _native.load()
    cols = list(Result.REGION_SEG_COLUMNS) + list(Result.REGION_PMD_COLUMNS)
    assert len(cols) == 24
    for ns, npm in ((0, 0), (1, 1), (31147, 1458), (5, 100003)):
        off = (C.c_int64 * 24)()
        tot = C.c_int64()
        assert lib.ml_gather_layout(ns, npm, off, C.byref(tot)) == 0
        at = 64
        for c, (name, dt) in enumerate(cols):
            assert off[c] == at and at % 16 == 0, (name, off[c], at)
            n = ns if c < 13 else npm
            at += (n * np.dtype(dt).itemsize + 15) // 16 * 16
        assert tot.value == at
    assert lib.ml_gather_layout(-1, 0, None, C.byref(tot)) != 0
