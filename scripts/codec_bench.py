"""Throughput of the text codec on a genome-sized Bismark .cov file (one replicate, ~28 M lines): GPU parse through the
C ABI (upload of the text from pageable and from page-locked memory included, fetch of the columns separately) next to
pandas.read_csv's C parser on the same bytes (the closest thing to data.table::fread in this image)."""
import ctypes as C
import io
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from methylasso_b200 import api, codec, synth  # noqa: E402

n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 28_000_000
rng = np.random.default_rng(1)
pos = np.cumsum(rng.geometric(1 / 60.0, n_lines)).astype(np.int64) + 10_000      # stays below 2^31
cov = rng.poisson(20, n_lines) + 1
nm = rng.binomial(cov, 0.8)
t0 = time.perf_counter()
pct = np.char.mod("%.15g", 100.0 * nm / cov)
lines = np.char.add(np.char.add(np.char.add(np.char.add("chr1\t", pos.astype(str)), "\t"), np.char.add(pos.astype(str), "\t")),
                    np.char.add(np.char.add(pct, "\t"), np.char.add(np.char.add(nm.astype(str), "\t"), (cov - nm).astype(str))))
text = ("\n".join(lines.tolist()) + "\n").encode()
print(f"text: {len(text) / 1e6:.0f} MB, {n_lines} lines, built in {time.perf_counter() - t0:.0f} s", flush=True)

eng = api.Engine(0)
lib = eng.lib


def gpu_parse(buf, nbytes):
    h, n, runs = C.c_void_p(), C.c_int64(0), C.c_int64(0)
    eng._check(lib.ml_text_parse(eng.h, buf, nbytes, 0, 4, 5, b"\x00", C.byref(h), C.byref(n), C.byref(runs)))
    return h, n.value


for _ in range(2):
    h, n = gpu_parse(text, len(text))
    lib.ml_text_free(h)
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    h, n = gpu_parse(text, len(text))
    ts.append(time.perf_counter() - t0)
    lib.ml_text_free(h)
print(f"GPU parse, text in pageable memory : {min(ts) * 1e3:8.1f} ms  {len(text) / min(ts) / 1e9:6.2f} GB/s of text", flush=True)
pin = api.pinned_empty(len(text), np.uint8, lib)
pin[:] = np.frombuffer(text, np.uint8)
pbuf = pin.ctypes.data_as(C.c_char_p)
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    h, n = gpu_parse(pbuf, len(text))
    ts.append(time.perf_counter() - t0)
    if _ < 2:
        lib.ml_text_free(h)
print(f"GPU parse, text in page-locked memory: {min(ts) * 1e3:8.1f} ms  {len(text) / min(ts) / 1e9:6.2f} GB/s of text", flush=True)
prof_on = eng.set("profile", 1)
lib.ml_text_free(h)
eng.profile_reset()
h, n = gpu_parse(pbuf, len(text))
for k, (ms, cnt) in sorted(eng.profile().items(), key=lambda kv: -kv[1][0]):
    print(f"    {k:20s} {ms:8.3f} ms x{cnt}")
out = [np.empty(n, np.int32), np.empty(n), np.empty(n), np.empty(n, np.int8)]
t0 = time.perf_counter()
eng._check(lib.ml_text_fetch(eng.h, h, *[api._p(o) for o in out], None, None, None, None))
print(f"fetch of pos, a, b, status (pageable): {(time.perf_counter() - t0) * 1e3:8.1f} ms")
lib.ml_text_free(h)
assert np.array_equal(out[0], pos.astype(np.int32)) and np.array_equal(out[1], nm.astype(float)) and not out[3].any()

import pandas as pd  # noqa: E402
t0 = time.perf_counter()
df = pd.read_csv(io.BytesIO(text), sep="\t", header=None, usecols=[0, 1, 4, 5], engine="c")
dt = time.perf_counter() - t0
print(f"pandas.read_csv (C engine, 1 thread) : {dt * 1e3:8.1f} ms  {len(text) / dt / 1e9:6.2f} GB/s of text")
assert np.array_equal(df[1].to_numpy(), pos)
