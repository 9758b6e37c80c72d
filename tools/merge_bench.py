"""Developer probe: device time of one local merge level (two sorted runs of n/2 keys each -> n keys).
usage: python tools/merge_bench.py [log2n=29] [type=u32] [iters=5] [jobs=1]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

rsb = importlib.import_module("radix-sort_b200")


def main():
    args = dict(a.split("=") for a in sys.argv[1:])
    log2n, t, iters, jobs = int(args.get("log2n", 29)), args.get("type", "u32"), int(args.get("iters", 5)), int(args.get("jobs", 1))
    n = 1 << log2n
    half = n // 2
    per = half // jobs
    kb = 4 if t in ("u32", "i32") else 8
    tdt = torch.int32 if kb == 4 else torch.int64
    ctx = rsb.Context(0)
    st = rsb.Stream(ctx, torch.cuda.current_stream().cuda_stream)
    src = torch.empty(half, dtype=tdt, device="cuda")
    bufs = [torch.empty(half, dtype=tdt, device="cuda") for _ in range(2)]
    runs = []
    s = rsb.Sorter(ctx, t, half, rsb.RS_NO_KEY_BUFFERS)
    for i in range(2):
        ctx.fill_dataset(st, t, "uniform", src.data_ptr(), half, seed=1234 + i)
        keep = torch.empty(half, dtype=tdt, device="cuda")
        r = s.sort_device(st, src.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), half)
        torch.cuda.synchronize()
        keep.copy_(bufs[r])
        runs.append(keep)
    s.close()
    del bufs, src
    out = torch.empty(n, dtype=tdt, device="cuda")
    splits = torch.empty(n // 1024 + 64, dtype=torch.int64, device="cuda")   # enough for the tile-shape variants too
    # `jobs` independent merges of equal size (slices of the two sorted runs are sorted runs themselves)
    job_list = [(runs[0].data_ptr() + j * per * kb, per, runs[1].data_ptr() + j * per * kb, per, out.data_ptr() + 2 * j * per * kb)
                for j in range(jobs)]
    torch.cuda.synchronize()
    for _ in range(2):
        ctx.merge_runs(st, t, job_list, splits.data_ptr(), splits.numel())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ctx.merge_runs(st, t, job_list, splits.data_ptr(), splits.numel())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bad = sum(ctx.check(st, t, out.data_ptr() + 2 * j * per * kb, 2 * per)["unordered"] for j in range(jobs))
    print(f"merge {t} n=2^{log2n} jobs={jobs}: {ms:.3f} ms  {n / ms / 1e6:.1f} Gkeys/s  {2 * kb * n / ms / 1e6:.0f} GB/s  unordered={bad}", flush=True)


main()
