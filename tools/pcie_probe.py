"""Raw pinned-memory PCIe rates on this box (the ceiling of the end-to-end numbers): python tools/pcie_probe.py"""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 4, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n // 4, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
print("H2D 1 GiB: %.1f GB/s" % (n / t(lambda: d.copy_(h, non_blocking=True)) / 1e9))
print("D2H 1 GiB: %.1f GB/s" % (n / t(lambda: h.copy_(d, non_blocking=True)) / 1e9))
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
dt = t(both)
print("D2H 1 GiB + H2D 256 MiB concurrently: %.2f ms -> %.1f GB/s of D2H payload" % (dt * 1e3, n / dt / 1e9))
def chunks():
    c = 64 << 20
    for o in range(0, n, c): h[o:o + c].copy_(d[o:o + c], non_blocking=True)
print("D2H 1 GiB in 64 MiB copies: %.1f GB/s" % (n / t(chunks) / 1e9))
