"""Small workload for compute-sanitizer (tools/sanitize.sh): every kernel of the library on inputs of a few tiles --
the onesweep pass in both ranking modes (and the pipelined variant with RSB200_SWEEP_V2=1), 4- and 8-byte keys with and
without the index payload, 2-byte and float keys, the one-CTA small sort, the merge tile kernels and the fused gather +
histogram kernel -- each checked against numpy so that a sanitizer run is also a parity run.
usage: python tools/sanitize_driver.py [log2n=17]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

rsb = importlib.import_module("radix-sort_b200")
dsort = importlib.import_module("radix-sort_b200.dist")


def main():
    args = dict(a.split("=") for a in sys.argv[1:])
    n = (1 << int(args.get("log2n", 17))) + 777
    rng = np.random.default_rng(9)
    ctx = rsb.Context(0)
    st = rsb.Stream(ctx, torch.cuda.current_stream().cuda_stream)
    bad = 0
    for t in ("u32", "i64", "u16", "f32"):
        dt = rsb.key_dtype(t)
        for flags in (0, rsb.RS_NO_ATOMIC_RANK):
            for with_perm in (False, True):
                for m in (n, 3000):   # multi-kernel path and the one-CTA path
                    if dt.kind == "f":
                        data = rng.standard_normal(m).astype(dt)
                    else:
                        info = np.iinfo(dt)
                        data = rng.integers(info.min, info.max, size=m, dtype=dt, endpoint=True)
                        data[: m // 3] = data[0]      # heavy ties
                    s = rsb.Sorter(ctx, t, m, flags | (rsb.RS_WITH_PERMUTATION if with_perm else 0))
                    s.upload(st, data, None, m)
                    s.sort(st, m)
                    out = np.empty_like(data)
                    perm = np.empty(m, np.uint32) if with_perm else None
                    s.download(st, out, perm, m)
                    s.close()
                    ok = np.array_equal(out, np.sort(data))
                    if with_perm:
                        ok = ok and np.array_equal(perm.astype(np.int64), np.argsort(data, kind="stable"))
                    bad += int(not ok)
                    print(f"sort {t} flags={flags} perm={with_perm} n={m}: {'OK' if ok else 'FAIL'}", flush=True)
    # merge kernels + merge tree (the multi-GPU final step on one GPU), fused gather + histogram
    import types
    for t in ("u32", "i64"):
        dt = rsb.key_dtype(t)
        info = np.iinfo(dt)
        sizes = [n // 2, 0, n // 3, 1, 4096]
        runs = [np.sort(rng.integers(info.min, info.max, size=k, dtype=dt, endpoint=True)) for k in sizes]
        tdt = torch.int32 if dt.itemsize == 4 else torch.int64
        pool = torch.from_numpy(np.concatenate(runs).view(np.int32 if dt.itemsize == 4 else np.int64).copy()).cuda()
        ptrs, off = [], 0
        for k in sizes:
            ptrs.append(pool.data_ptr() + off * dt.itemsize)
            off += k
        total = sum(sizes)
        bx, by = (torch.empty(total + 32 * len(sizes), dtype=tdt, device="cuda") for _ in range(2))
        torch.cuda.synchronize()
        out = dsort._merge_tree(types.SimpleNamespace(ctx=ctx, stream=st), t, ptrs, sizes, bx, by, total, dt.itemsize, {})
        st.sync()
        ok = np.array_equal(out.cpu().numpy().view(dt), np.sort(np.concatenate(runs)))
        bad += int(not ok)
        print(f"merge tree {t}: {'OK' if ok else 'FAIL'}", flush=True)
        srt = rsb.Sorter(ctx, t, total, rsb.RS_NO_KEY_BUFFERS | rsb.RS_NO_PROFILE)
        g, b0 = (torch.empty(total, dtype=tdt, device="cuda") for _ in range(2))
        r = srt.sort_segments_device(st, ptrs, sizes, 1, g.data_ptr(), b0.data_ptr(), g.data_ptr())
        st.sync()
        ok = np.array_equal((b0, g)[r].cpu().numpy().view(dt), np.sort(np.concatenate(runs)))
        srt.close()
        bad += int(not ok)
        print(f"gather + sort {t}: {'OK' if ok else 'FAIL'}", flush=True)
    print("sanitize_driver", "PASSED" if bad == 0 else f"FAILED ({bad})")
    sys.exit(1 if bad else 0)


main()
