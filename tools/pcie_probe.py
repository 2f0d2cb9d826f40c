import torch, time
n = 52_900_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(13_728_000, dtype=torch.uint8).pin_memory()
d2 = torch.empty(13_728_000, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("H2D 52.9 MB: %.3f ms  %.1f GB/s" % (ms, n / ms / 1e6))
t0 = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / 20
print("H2D 52.9 MB + D2H 13.7 MB concurrent: %.3f ms per pair" % ms)
