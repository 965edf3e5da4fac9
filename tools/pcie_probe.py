"""experiment: host<->device bandwidth of the box with pinned buffers (one direction, and both at once on two streams)."""
import time
import torch

n = 1 << 30  # 4 GiB of f32
hx = torch.empty(n, dtype=torch.float32, pin_memory=True)
hy = torch.empty(n, dtype=torch.float32, pin_memory=True)
dx = torch.empty(n, dtype=torch.float32, device="cuda")
dy = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 4 / 1e9


def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


def h2d():
    with torch.cuda.stream(s1):
        dx.copy_(hx, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hy.copy_(dy, non_blocking=True)


def both():
    h2d(); d2h()


def chunked(chunk_mb):
    c = chunk_mb * (1 << 20) // 4
    def f():
        for o in range(0, n, c):
            with torch.cuda.stream(s1):
                dx[o:o + c].copy_(hx[o:o + c], non_blocking=True)
            with torch.cuda.stream(s2):
                hy[o:o + c].copy_(dy[o:o + c], non_blocking=True)
    return f


a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D {gb / a:.1f} GB/s   D2H {gb / b:.1f} GB/s   both at once: {gb / c:.1f} GB/s each direction ({c * 1e3:.1f} ms for 4.29 GB each way)")
for mb in (16, 64, 128):
    d = t(chunked(mb))
    print(f"both at once in {mb} MB chunks: {gb / d:.1f} GB/s each direction ({d * 1e3:.1f} ms)")
