"""Generate tests/golden/ from the REFERENCE's own code (oracle/_ref/libref_radix.so, i.e.
/root/reference/src/{CRadixSortCPU,Dataset,Parameters}.h compiled by oracle/Makefile).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Outputs (committed; they travel to the GPU box where /root/reference does not exist):
  golden_small.npz   datasets and RadixSortCPU / std::sort outputs at N = 1024 and a ragged tiny set
  golden_digests.json  sha256 of dataset bytes and RadixSortCPU outputs at N = 2^16 and 2^20
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import KEY_TYPES, Reference, build  # noqa: E402

UNIFORM_SEED = 0x5EED
DET_KINDS = ("zeros", "range", "inverted", "random")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    build()
    ref = Reference()
    small = {}
    digests = {"uniform_seed": UNIFORM_SEED, "datasets": {}, "radix_sort_cpu": {}, "notes": {}}
    for t in KEY_TYPES:
        # the deterministic generator head the survey quotes (src/Dataset.h:110-120)
        small[f"random_head_{t}"] = ref.dataset("random", t, 4096)
        small[f"uniform_head_{t}"] = ref.uniform_seeded(t, 1024, UNIFORM_SEED)
        for kind in DET_KINDS + ("uniform",):
            for n in (1024,):
                data = ref.uniform_seeded(t, n, UNIFORM_SEED) if kind == "uniform" else ref.dataset(kind, t, n)
                small[f"in_{kind}_{t}_{n}"] = data
                small[f"radixcpu_{kind}_{t}_{n}"] = ref.radix_sort_cpu(data, t)
                small[f"stdsort_{kind}_{t}_{n}"] = ref.std_sort(data, t)
            for n in (1 << 16, 1 << 20):
                data = ref.uniform_seeded(t, n, UNIFORM_SEED) if kind == "uniform" else ref.dataset(kind, t, n)
                digests["datasets"][f"{kind}_{t}_{n}"] = sha(data)
                digests["radix_sort_cpu"][f"{kind}_{t}_{n}"] = sha(ref.radix_sort_cpu(data, t))
        # tiny inputs where RadixSortCPU is known to differ from std::sort: keep what it returns
        for kind, n in (("inverted", 2), ("random", 3), ("random", 4), ("range", 7), ("random", 100)):
            if t[0] == "i" and kind in ("range", "inverted") and n == 1:
                continue
            data = ref.dataset(kind, t, n)
            small[f"in_{kind}_{t}_{n}"] = data
            small[f"radixcpu_{kind}_{t}_{n}"] = ref.radix_sort_cpu(data, t)
    digests["notes"]["num_bins"] = {"32": int(ref.lib.ref_num_bins_32()), "64": int(ref.lib.ref_num_bins_64())}
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **small)
    with open(os.path.join(HERE, "golden_digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)
    print("wrote", len(small), "arrays,", len(digests["radix_sort_cpu"]), "digests")


if __name__ == "__main__":
    main()
