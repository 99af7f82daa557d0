"""PCIe probe for the end-to-end path: pinned H2D / D2H bandwidth alone and concurrently (full duplex), sizes of the 1M-sphere e2e step."""
import torch, time
a = torch.empty(72_000_000, dtype=torch.uint8).pin_memory(); b = torch.empty(48_000_000, dtype=torch.uint8).pin_memory()
da = torch.empty_like(a, device="cuda"); db = torch.empty_like(b, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): b.copy_(db, non_blocking=True)
def both(): h2d(); d2h()
th, td, tb = t(h2d), t(d2h), t(both)
print("H2D 72MB %.3f ms (%.1f GB/s)  D2H 48MB %.3f ms (%.1f GB/s)  both %.3f ms" % (th, 72 / th, td, 48 / td, tb))
