"""CPU tests of the multi-GPU host logic (gloo, world size 2): LPT sharding, the all-gather of
per-shard segment tables, and that the sharded result equals the single-process result."""
import os
import socket

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
