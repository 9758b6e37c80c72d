"""Developer probe (not the contract bench): per-phase device times of a device-resident sort.
usage: python tools/quick_bench.py [log2n=28] [types=u32,...] [kinds=uniform,...] [perm=0|1] [notma=0|1]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

rsb = importlib.import_module("radix-sort_b200")


def main():
    args = dict(a.split("=") for a in sys.argv[1:])
    log2n = int(args.get("log2n", 28))
    types = args.get("types", "u32").split(",")
    kinds = args.get("kinds", "uniform,random,range,inverted,zeros").split(",")
    perm = int(args.get("perm", 0))
    notma = int(args.get("notma", 0))
    noatomic = int(args.get("noatomic", 0))
    iters = int(args.get("iters", 5))
    n = 1 << log2n
    ctx = rsb.Context(0)
    st = rsb.Stream(ctx, torch.cuda.current_stream().cuda_stream)
    for t in types:
        kb = 4 if t in ("u32", "i32") else 8
        tdt = torch.int32 if kb == 4 else torch.int64
        d_in = torch.empty(n, dtype=tdt, device="cuda")
        bufs = [torch.empty(n, dtype=tdt, device="cuda") for _ in range(2)]
        vin = torch.arange(n, dtype=torch.int32, device="cuda") if perm else None
        vb = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)] if perm else None
        s = rsb.Sorter(ctx, t, n, rsb.RS_NO_KEY_BUFFERS | (rsb.RS_NO_TMA if notma else 0) | (rsb.RS_NO_ATOMIC_RANK if noatomic else 0))
        for kind in kinds:
            ctx.fill_dataset(st, t, kind, d_in.data_ptr(), n)
            for _ in range(2):
                r = s.sort_device(st, d_in.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), n,
                                  vin.data_ptr() if perm else None, vb[0].data_ptr() if perm else None,
                                  vb[1].data_ptr() if perm else None)
            torch.cuda.synchronize()
            s.reset_runtimes()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                s.sort_device(st, d_in.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), n,
                              vin.data_ptr() if perm else None, vb[0].data_ptr() if perm else None,
                              vb[1].data_ptr() if perm else None)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            rt = s.runtimes()
            chk = ctx.check(st, t, bufs[r].data_ptr(), n)
            passes = kb
            bpk = kb + passes * 2 * (kb + (4 if perm else 0))
            print(f"{t} {kind:9s} n=2^{log2n} perm={perm} notma={notma} atomic={int(ctx.atomic_rank and not noatomic)}: {ms:8.3f} ms  {n / ms / 1e6:7.2f} Gkeys/s  "
                  f"{bpk * n / ms / 1e6:7.1f} GB/s algorithmic | hist {rt['time_histo']['avg']:.3f} scan "
                  f"{rt['time_scan']['avg']:.3f} passes {rt['time_reorder']['avg']:.3f} "
                  f"({rt['time_reorder']['avg'] / passes:.3f}/pass) total {rt['time_total']['avg']:.3f} "
                  f"unordered={chk['unordered']}", flush=True)
        s.close()


if __name__ == "__main__":
    main()
