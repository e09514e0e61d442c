"""Raw pinned-memory copy rates of the box (context for bench.py's e2e leg, which is PCIe-bound)."""
import torch, time
dev = torch.device("cuda", 0)
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
t = run(lambda: d.copy_(h, non_blocking=True)); print("H2D alone  %.1f GB/s" % (n / t / 1e9))
t = run(lambda: h2.copy_(d2, non_blocking=True)); print("D2H alone  %.1f GB/s" % (n / t / 1e9))
def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
t = run(both); print("H2D + D2H concurrently: %.1f GB/s each way" % (n / t / 1e9))
m = 41 << 20
hs, ds = h[:m], d[:m]
t = run(lambda: ds.copy_(hs, non_blocking=True), reps=50); print("H2D 41 MB chunks %.1f GB/s" % (m / t / 1e9))
