import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
rsb = importlib.import_module("radix-sort_b200")
from oracle import Oracle
o = Oracle()
ctx = rsb.Context(0); st = rsb.Stream(ctx)
for t in ("u32","u64"):
  for n in (11264*3, 11264*7+5, 100000, 1<<20):
    for kind in ("uniform","zeros","range"):
        data = o.dataset(kind, t, n)
        s = rsb.Sorter(ctx, t, n, rsb.RS_WITH_PERMUTATION)
        s.upload(st, data, None, n); s.sort(st, n)
        out = np.empty_like(data); perm = np.empty(n, np.uint32)
        s.download(st, out, perm, n)
        h = s.download_histograms(st)
        ok = np.array_equal(out, np.sort(data)) and np.array_equal(perm, np.argsort(data, kind="stable").astype(np.uint32))
        print(t, n, kind, "OK" if ok else "FAIL", int(h[0].sum()), flush=True)
        if not ok:
            bad = np.nonzero(out != np.sort(data))[0]
            print("  first bad", bad[:5], "count", bad.size)
        s.close()
