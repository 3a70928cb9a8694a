"""Multi-GPU host logic: jobs (chromosome x sample) are independent, so they are sharded across
ranks with no data-path collective; only per-shard summaries are exchanged at the end.

Mirrors what the reference does with `foreach(ch = chromosomes) %dopar%` + `combine_ret`
(/root/reference/R/genome_wide.R:6-15, 42-43): one worker per chromosome, results concatenated in
chromosome order.  Here a worker is a GPU rank and the concatenation is an all-gather.
"""
from __future__ import annotations

import numpy as np


def lpt_assign(costs, world_size):
    """Longest-processing-time-first: returns rank of each job.  Deterministic (ties by job index)."""
    costs = np.asarray(costs, dtype=np.float64)
    order = sorted(range(costs.size), key=lambda j: (-costs[j], j))
    load = [0.0] * world_size
    rank_of = np.empty(costs.size, dtype=np.int64)
    for j in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        rank_of[j] = r
        load[r] += costs[j]
    return rank_of


def job_cost(n_rows, n_est=1, irls_steps=1):
    """Relative cost of a job: rows swept per evaluation + parameters through the FLSA per step."""
    return float(n_rows) * (1.0 + irls_steps) + float(n_rows) * n_est * irls_steps


def my_jobs(rank_of, rank):
    return [int(j) for j in np.nonzero(np.asarray(rank_of) == rank)[0]]


def all_gather_tables(local, dist, device=None):
    """All-gather a dict of equally long 1-D arrays with a different length per rank.

    Two collectives: lengths, then one padded payload per column (float64 / int64 carriers).
    Returns the rank-ordered concatenation as numpy arrays.  `dist` is torch.distributed
    (backend nccl on GPUs, gloo in the CPU tests)."""
    import torch
    world = dist.get_world_size()
    keys = sorted(local)
    n = int(len(local[keys[0]])) if keys else 0
    dev = device or "cpu"
    lens = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(lens, torch.tensor([n], dtype=torch.int64, device=dev))
    lens = [int(x.item()) for x in lens]
    cap = max(lens + [1])
    out = {}
    for k in keys:
        a = np.asarray(local[k])
        is_int = np.issubdtype(a.dtype, np.integer)
        carrier = torch.int64 if is_int else torch.float64
        buf = torch.zeros(cap, dtype=carrier, device=dev)
        if n:
            buf[:n] = torch.from_numpy(a.astype(np.int64 if is_int else np.float64)).to(dev)
        parts = [torch.zeros(cap, dtype=carrier, device=dev) for _ in range(world)]
        dist.all_gather(parts, buf)
        out[k] = np.concatenate([p[:m].cpu().numpy() for p, m in zip(parts, lens)]).astype(a.dtype)
    return out, lens


def all_gather_packed(local, dist, device=None):
    """The exchange itself, for tables whose columns are float64 or 32-bit integers: ONE payload collective per table,
    the columns travelling as the rows of one float64 matrix (32-bit integers are exact in it).  Returns
    (parts, lens, keys, dtypes): parts[r] is rank r's matrix as a tensor on `device` (every rank ends up with all of
    them), lens[r] its number of rows.  unpack_gathered() turns that into the concatenated table on the host."""
    import torch
    world = dist.get_world_size()
    keys = sorted(local)
    n = int(len(local[keys[0]])) if keys else 0
    dev = device or "cpu"
    dtypes = [np.asarray(local[k]).dtype for k in keys]
    for k, dt in zip(keys, dtypes):
        if not (dt == np.float64 or (np.issubdtype(dt, np.integer) and dt.itemsize <= 4)):
            raise TypeError(f"column {k}: {dt} does not travel exactly in a float64 matrix")
    lens = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(lens, torch.tensor([n], dtype=torch.int64, device=dev))
    lens = [int(x.item()) for x in lens]
    cap = max(lens + [1])
    mat = np.zeros((len(keys), cap), np.float64)
    for i, k in enumerate(keys):
        mat[i, :n] = local[k]
    buf = torch.from_numpy(mat).to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return parts, lens, keys, dtypes


def unpack_gathered(parts, lens, keys, dtypes):
    host = [p.cpu().numpy() if hasattr(p, "cpu") else np.asarray(p) for p in parts]
    return {k: np.concatenate([h[i, :m] for h, m in zip(host, lens)]).astype(dt)
            for i, (k, dt) in enumerate(zip(keys, dtypes))}


def all_gather_tables_packed(local, dist, device=None):
    """all_gather_packed + unpack_gathered: same result as all_gather_tables with one collective per table."""
    parts, lens, keys, dtypes = all_gather_packed(local, dist, device)
    return unpack_gathered(parts, lens, keys, dtypes), lens


def reorder_by_job(table, job_key="job"):
    """Stable sort of a gathered table by global job index (chromosome order of the input)."""
    order = np.argsort(table[job_key], kind="stable")
    return {k: v[order] for k, v in table.items()}
