"""Developer probe: host<->device copy rates of 1 GiB pinned buffers -- each direction alone, both at once, and in
chunks -- to tell what bounds the end-to-end number of bench.py (which moves 1 GiB each way per batch).
usage: python tools/pcie_probe.py"""
import time
import torch

N = 1 << 30
dev = torch.device("cuda:0")
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
h_in.fill_(7)
d_a = torch.empty(N, dtype=torch.uint8, device=dev)
d_b = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


def both_chunked(chunks=8):
    c = N // chunks
    for i in range(chunks):
        with torch.cuda.stream(s1):
            d_a[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[i * c:(i + 1) * c].copy_(d_b[i * c:(i + 1) * c], non_blocking=True)


cases = [("H2D alone", h2d), ("D2H alone", d2h), ("H2D + D2H at once", both)]
for c in (2, 4, 8, 16, 32, 64, 128, 256):
    cases.append((f"H2D + D2H, {c} chunks each", lambda c=c: both_chunked(c)))
for name, fn in cases:
    t = timed(fn)
    print(f"{name:28s} {t * 1e3:8.2f} ms per GiB (each way)  {N / t / 1e9:6.1f} GB/s per direction", flush=True)
