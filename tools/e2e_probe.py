"""Developer probe: the three-lane end-to-end pipeline of bench.py (upload i+2 | sort i+1 | download i through the C ABI,
pinned host buffers) with the host<->device copies issued whole or in pieces (RSB200_COPY_CHUNK_MIB), on ONE box.
usage: python tools/e2e_probe.py [chunks=0,32,128,512] [reps=2]"""
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

rsb = importlib.import_module("radix-sort_b200")


def main():
    args = dict(a.split("=") for a in sys.argv[1:])
    chunks = [int(c) for c in args.get("chunks", "0,32,128,512").split(",")]
    reps = int(args.get("reps", 2))
    n = 1 << 28
    ctx = rsb.Context(0)
    h_in = rsb.HostBuffer(n, np.uint32)
    h_in.array[:] = np.random.default_rng(1).integers(0, 2**32, size=n, dtype=np.uint32)
    lanes = [(rsb.Sorter(ctx, "u32", n), rsb.Stream(ctx), rsb.HostBuffer(n, np.uint32)) for _ in range(3)]

    def run(steps):
        for i in range(steps):
            srt, stream, hout = lanes[i % 3]
            stream.sync()
            srt.upload(stream, h_in, None, n)
            srt.sort(stream, n)
            srt.download_async(stream, hout, None, n)
        for _, stream, _ in lanes:
            stream.sync()

    for rep in range(reps):
        for c in chunks:
            os.environ["RSB200_COPY_CHUNK_MIB"] = str(c)
            run(3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(9)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3 / 9
            ok = bool(np.all(lanes[0][2].array[:-1] <= lanes[0][2].array[1:]))
            print(f"rep {rep} pieces of {c:4d} MiB: {ms:7.2f} ms per batch  {n / ms / 1e6:6.2f} Gkeys/s  {n * 4 / ms / 1e6:5.1f} GB/s each way  sorted={ok}", flush=True)


main()
