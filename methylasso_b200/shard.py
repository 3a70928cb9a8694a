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


# ---------------------------------------------------------------------------------------------
# K16, native: the region tables of every shard pushed into every rank's HBM over NVLink
# ---------------------------------------------------------------------------------------------
class _DeviceBytes:
    """A raw device range as an object torch.as_tensor understands (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class RegionGather:
    """All-gather of the two tables `segment_methylation` returns, straight from the resident tables of each rank's
    result (`ml_gather_*`, include/methylasso_b200.h): the reference's rbindlist over its per-chromosome workers
    (R/call.R:230-244, R/genome_wide.R:6-15).

    transport "ipc": every rank maps every other rank's window (CUDA IPC) and a push is one kernel storing into all of
    them; "nccl": the push fills the local slot only and `torch.distributed.all_gather_into_tensor` moves the slots
    (used when the windows cannot be mapped).  Either way the tables go HBM -> HBM; `dist` is only used for the
    64-byte handles at start-up and for the per-call counts, whose collective is also the barrier between the pushes
    and the reads.  world_size 1 needs no `dist`."""

    def __init__(self, eng, rank, world, slot_bytes, dist=None, device=None, transport="ipc"):
        import ctypes as C
        self.eng, self.rank, self.world, self.dist, self.device = eng, int(rank), int(world), dist, device
        self.h = C.c_void_p()
        eng._check(eng.lib.ml_gather_create(eng.h, self.rank, self.world, int(slot_bytes), C.byref(self.h)))
        eng._adopt(self)           # the window is released before the engine, whatever order they are let go of in
        sb = C.c_int64()
        p = C.c_void_p()
        eng._check(eng.lib.ml_gather_window(self.h, 0, C.byref(p), C.byref(sb)))
        self.slot_bytes = sb.value
        self.parity = 0
        self.transport = transport if world > 1 else "local"
        self._host = None
        if world > 1 and transport == "ipc":
            try:
                self._open_peers()
            except Exception as exc:           # windows cannot be mapped here: let NCCL move the slots
                self.transport = "nccl"
                self.ipc_error = repr(exc)
            # every rank must use the same transport
            import torch
            flag = torch.tensor([1 if self.transport == "ipc" else 0], dtype=torch.int64, device=device or "cpu")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.transport = "nccl"

    def _open_peers(self):
        import ctypes as C
        import torch
        mine = (C.c_ubyte * 64)()
        self.eng._check(self.eng.lib.ml_gather_ipc_handle(self.h, mine))
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device or "cpu")
        parts = [torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t)
        allh = np.concatenate([p.cpu().numpy() for p in parts]).astype(np.uint8)
        self.eng._check(self.eng.lib.ml_gather_open(self.h, allh.ctypes.data_as(C.c_void_p)))

    def close(self):
        if self.h:
            self.eng.lib.ml_gather_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def layout(self, n_seg, n_pmd):
        import ctypes as C
        off = (C.c_int64 * 24)()
        tot = C.c_int64()
        self.eng._check(self.eng.lib.ml_gather_layout(int(n_seg), int(n_pmd), off, C.byref(tot)))
        return list(off), tot.value

    def push(self, res, job_ids, tol_val=0.01, max_distance=500.0, drop=True):
        """Rule engine (if not done yet) + pack + store into this rank's slot everywhere.  Returns (n_seg, n_pmd)."""
        import ctypes as C
        ids = np.ascontiguousarray(job_ids, np.int32)
        ns, npm = C.c_int64(), C.c_int64()
        self.eng._check(self.eng.lib.ml_gather_push(self.eng.h, self.h, res.h, float(tol_val), float(max_distance), None, None,
                                                    None, 1 if drop else 0, ids.ctypes.data_as(C.c_void_p), self.parity,
                                                    1 if self.transport == "ipc" else 0, C.byref(ns), C.byref(npm)))
        return ns.value, npm.value

    def exchange_counts(self, n_seg, n_pmd):
        """All ranks' (n_seg, n_pmd); with the "nccl" transport the slots travel here too.  Returns an int64 [world, 2]."""
        if self.world == 1:
            return np.array([[n_seg, n_pmd]], np.int64)
        import ctypes as C
        import torch
        mine = torch.tensor([n_seg, n_pmd], dtype=torch.int64, device=self.device or "cpu")
        parts = [torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(parts, mine)
        if self.transport == "nccl":
            p = C.c_void_p()
            self.eng._check(self.eng.lib.ml_gather_window(self.h, self.parity, C.byref(p), None))
            win = torch.as_tensor(_DeviceBytes(p.value, self.world * self.slot_bytes), device=self.device)
            self.dist.all_gather_into_tensor(win, win[self.rank * self.slot_bytes:(self.rank + 1) * self.slot_bytes].clone())
            torch.cuda.current_stream().synchronize()
        return torch.stack(parts).cpu().numpy()

    def fetch(self, counts):
        """The gathered tables on the host, shards in rank order: (lmr_umr_valley, pmd) as dicts of columns."""
        import ctypes as C
        from .api import Result, pinned_empty
        lay = [self.layout(int(a), int(b)) for a, b in counts]
        used = np.array([t for _, t in lay], np.int64)
        if self._host is None:
            self._host = pinned_empty(self.world * self.slot_bytes, np.uint8, self.eng.lib)
        self.eng._check(self.eng.lib.ml_gather_fetch(self.eng.h, self.h, self.parity, used.ctypes.data_as(C.c_void_p),
                                                     self._host.ctypes.data_as(C.c_void_p)))
        cols = list(Result.REGION_SEG_COLUMNS) + list(Result.REGION_PMD_COLUMNS)
        seg = {k: [] for k, _ in Result.REGION_SEG_COLUMNS}
        pmd = {k: [] for k, _ in Result.REGION_PMD_COLUMNS}
        for w, ((ns, npm), (off, _)) in enumerate(zip(counts, lay)):
            base = w * self.slot_bytes
            hdr = self._host[base:base + 16].view(np.int64)
            if int(hdr[0]) != int(ns) or int(hdr[1]) != int(npm):
                raise RuntimeError(f"slot of rank {w}: header {hdr.tolist()} does not match the counts {(int(ns), int(npm))}")
            for c, (k, dt) in enumerate(cols):
                n = int(ns) if c < 13 else int(npm)
                a = self._host[base + off[c]:base + off[c] + n * np.dtype(dt).itemsize].view(dt)
                (seg if c < 13 else pmd)[k].append(a.copy())
        self.parity ^= 1
        return ({k: np.concatenate(v) for k, v in seg.items()}, {k: np.concatenate(v) for k, v in pmd.items()})

    def all_gather(self, res, job_ids, tol_val=0.01, max_distance=500.0, drop=True, fetch=True):
        ns, npm = self.push(res, job_ids, tol_val, max_distance, drop)
        counts = self.exchange_counts(ns, npm)
        if not fetch:
            self.parity ^= 1
            return counts
        return self.fetch(counts)
