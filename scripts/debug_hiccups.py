"""Where do the occasional +25 ms steps of the resident bench loop come from?  Step times with (a) no clock sampler,
(b) `nvidia-smi -lms 100` running, (c) an in-process NVML thread polling every 20 ms."""
import gc, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from methylasso_b200 import api
from methylasso_b200._native import MlParams

eng = api.Engine(0)
stream = torch.cuda.current_stream()
eng.set_stream(stream.cuda_stream)
tables, ptr, cols = bench.make_workload(28_000_000)
batch = eng.upload(ptr, cols)
params = MlParams(25.0, 26.0, 0.0, 0.01, 50.0, 1.0, 1, 1, 1, 0, 0, 0)

def step():
    res = eng.detect(batch, params)
    res.segments_count(0.01)
    res.call_table(0.01, 500.0, fetch=False)
    res.call_pmd_segments(0.01, 500.0, 1000.0, fetch=False)
    res.free()

for _ in range(4):
    step()
eng.set("pool_headroom_percent", 25)

def run(label, n=40):
    gc.collect(); gc.disable()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); step(); ts.append((time.perf_counter() - t0) * 1e3)
    gc.enable()
    ts = np.array(ts)
    print(f"{label:28s} median {np.median(ts):6.2f} ms  max {ts.max():6.2f}  steps > 1.3 x median: {(ts > 1.3 * np.median(ts)).sum()} of {n}", flush=True)

run("no sampler")
s = bench.ClockSampler(0); s.start(); time.sleep(1.0)
run("nvidia-smi -lms 100")
print(s.stop())
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = threading.Event(); samples = []
def poll():
    while not stop.is_set():
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.02)
th = threading.Thread(target=poll, daemon=True); th.start(); time.sleep(0.5)
run("NVML thread, 20 ms")
stop.set(); th.join()
print(len(samples), samples[-3:])
run("no sampler again")
eng.set("profile", 1)
for _ in range(3):
    step()
eng.set("profile_reserve_events", 2 * 400 * 45)
eng.profile_reset()
run("per-kernel events on")
eng.profile_reset()
run("per-kernel events on, again")
