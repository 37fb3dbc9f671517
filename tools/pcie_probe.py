"""Pinned host <-> device copy bandwidth of the box (one direction, both directions at once,
two streams in one direction).  usage: python tools/pcie_probe.py"""
import torch

n = 64 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
h3 = torch.empty(n, dtype=torch.uint8).pin_memory(); h4 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty_like(d1); d3 = torch.empty_like(d1); d4 = torch.empty_like(d1)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s1.synchronize(); s2.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def on(stream, f):
    with torch.cuda.stream(stream):
        f()


print("H2D one stream      %.1f GB/s" % (n / timed(lambda: on(s1, lambda: d1.copy_(h1, non_blocking=True))) / 1e9))
print("D2H one stream      %.1f GB/s" % (n / timed(lambda: on(s1, lambda: h2.copy_(d2, non_blocking=True))) / 1e9))
t = timed(lambda: (on(s1, lambda: d1.copy_(h1, non_blocking=True)), on(s2, lambda: h2.copy_(d2, non_blocking=True))))
print("H2D + D2H at once   %.1f GB/s each" % (n / t / 1e9))
t = timed(lambda: (on(s1, lambda: h2.copy_(d2, non_blocking=True)), on(s2, lambda: h4.copy_(d4, non_blocking=True))))
print("D2H on two streams  %.1f GB/s total" % (2 * n / t / 1e9))
t = timed(lambda: (on(s1, lambda: d1.copy_(h1, non_blocking=True)), on(s2, lambda: d3.copy_(h3, non_blocking=True))))
print("H2D on two streams  %.1f GB/s total" % (2 * n / t / 1e9))
for mb in (1, 4):
    m = mb << 20
    t = timed(lambda: on(s1, lambda: h2[:m].copy_(d2[:m], non_blocking=True)), reps=50)
    print("D2H %d MB pieces     %.1f GB/s" % (mb, m / t / 1e9))
