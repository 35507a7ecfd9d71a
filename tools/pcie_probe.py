"""Pinned-memory PCIe bandwidth of the box (floor for the host-pointer path)."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, it=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / it
a = t(lambda: d.copy_(h, non_blocking=True))
b = t(lambda: h.copy_(d, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print(f"H2D {n / a / 1e9:.1f} GB/s, D2H {n / b / 1e9:.1f} GB/s, both at once {n / c / 1e9:.1f} GB/s each direction")
