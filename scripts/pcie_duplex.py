"""PCIe characterisation of the GPU box: H2D alone, D2H alone, both at once (pinned, two streams)."""
import torch, time
n = 1 << 28  # 1 GiB of float32
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device="cuda"); d_b = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
gb = n * 4 / 1e9
print(f"H2D {gb/t(h2d):.1f} GB/s  D2H {gb/t(d2h):.1f} GB/s  both: {2*gb/t(both):.1f} GB/s aggregate ({t(both)*1e3:.1f} ms for {gb:.2f}+{gb:.2f} GB)")
