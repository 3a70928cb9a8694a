"""Summarise an `ncu --set full --csv --page raw` export into profiles/ncu_traffic.json.

    python scripts/ncu_traffic.py gpurun_out/ncu_full_raw.csv profiles/ncu_traffic.json

Per kernel (function name): launches captured, mean DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum), mean duration, achieved occupancy and issue-slot utilisation.
"""
import csv
import json
import re
import sys
from collections import defaultdict


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


def main(src, dst):
    rows = list(csv.reader(open(src, newline="")))
    # ncu raw page: header row, units row, then one row per launch
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    unit = {n: units[i] for n, i in col.items()}

    def scaled(r, key):
        if key not in col:
            return None
        v = num(r[col[key]])
        if v is None:
            return None
        u = unit[key].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
                "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "second": 1e3}.get(u, 1)
        return v * mult

    acc = defaultdict(lambda: defaultdict(list))
    seen_main = defaultdict(int)
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        full = r[col["Kernel Name"]]
        k = re.sub(r"^void ", "", full.strip())
        k = re.sub(r"^ml::", "", k)
        k = re.sub(r"^(<?unnamed>|\(anonymous namespace\))::", "", k)
        k = re.sub(r"[<(].*", "", k)
        acc[k]["dram"].append((scaled(r, "dram__bytes_read.sum") or 0) + (scaled(r, "dram__bytes_write.sum") or 0))
        acc[k]["ms"].append(scaled(r, "gpu__time_duration.sum"))
        for key, short in (("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
                           ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
                           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                           ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
                           ("launch__registers_per_thread", "registers")):
            if key in col:
                acc[k][short].append(num(r[col[key]]))
    import datetime
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    # bench.py refuses the numbers below when the kernel sources have changed since this capture
    out = {"source": src, "kernel_source_hash": bench.kernel_source_hash(),
           "captured": datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%MZ"),
           "dram_bytes_per_launch": {}, "detail": {}}
    for k, d in acc.items():
        n = len(d["dram"])
        out["dram_bytes_per_launch"][k] = sum(d["dram"]) / n
        det = {"launches_captured": n}
        for key, v in d.items():
            v = [x for x in v if x is not None]
            if v:
                det["mean_" + key] = sum(v) / len(v)
        out["detail"][k] = det
    # bench.py names a launch by what it does, not by the specialisation that ran
    for fn, name in (("k_pseudodata_fast", "k_pseudodata"),):
        if fn in out["dram_bytes_per_launch"] and name not in out["dram_bytes_per_launch"]:
            out["dram_bytes_per_launch"][name] = out["dram_bytes_per_launch"][fn]
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(out["dram_bytes_per_launch"], indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
