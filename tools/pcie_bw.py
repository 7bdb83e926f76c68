import torch, time
n = 25_000_000  # 200 MB of float64
h = torch.empty(n, dtype=torch.float64).pin_memory(); h2 = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, k=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k
a = t(lambda: d.copy_(h, non_blocking=True)); print(f"H2D 200MB: {a*1e3:.2f} ms = {0.2/a:.1f} GB/s")
b = t(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H 200MB: {b*1e3:.2f} ms = {0.2/b:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print(f"H2D || D2H 200MB each: {c*1e3:.2f} ms = {0.4/c:.1f} GB/s aggregate")
