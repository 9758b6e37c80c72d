"""Size sweep 2^0 .. 2^28 (BASELINE.json configs[2]): device-resident sort time per size, keys-only and key+index.
usage: python tools/sweep.py [type=u32] [max=28]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

rsb = importlib.import_module("radix-sort_b200")


def main():
    args = dict(a.split("=") for a in sys.argv[1:])
    t = args.get("type", "u32")
    mx = int(args.get("max", 28))
    kb = 4 if t in ("u32", "i32") else 8
    tdt = torch.int32 if kb == 4 else torch.int64
    ctx = rsb.Context(0)
    st = rsb.Stream(ctx, torch.cuda.current_stream().cuda_stream)
    nmax = 1 << mx
    d_in = torch.empty(nmax, dtype=tdt, device="cuda")
    bufs = [torch.empty(nmax, dtype=tdt, device="cuda") for _ in range(2)]
    vin = torch.arange(nmax, dtype=torch.int32, device="cuda")
    vb = [torch.empty(nmax, dtype=torch.int32, device="cuda") for _ in range(2)]
    ctx.fill_dataset(st, t, "uniform", d_in.data_ptr(), nmax)
    s = rsb.Sorter(ctx, t, nmax, rsb.RS_NO_KEY_BUFFERS | rsb.RS_NO_PROFILE)
    print(f"{'log2n':>5} {'keys us':>10} {'Gkeys/s':>9} {'key+idx us':>11} {'Gkeys/s':>9}")
    for k in range(0, mx + 1):
        n = 1 << k
        row = []
        for perm in (0, 1):
            def once():
                s.sort_device(st, d_in.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), n,
                              vin.data_ptr() if perm else None, vb[0].data_ptr() if perm else None,
                              vb[1].data_ptr() if perm else None)
            for _ in range(3):
                once()
            iters = 200 if k < 20 else (20 if k < 26 else 5)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                once()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            row += [us, n / us / 1e3]
        print(f"{k:5d} {row[0]:10.1f} {row[1]:9.3f} {row[2]:11.1f} {row[3]:9.3f}", flush=True)
    s.close()


main()
