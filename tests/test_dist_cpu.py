"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU host logic: exact splitters, tie splitting by
count, cut positions, the all-to-all-v plumbing.  The device operations are replaced by a numpy test
double here (the product backend is CudaOps over librsb200.so; the N>1 GPU path is exercised by
tools/dist_check.py under torchrun on the GPU box)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyOps:
    """Test double for radix-sort_b200.dist.CudaOps (CPU tensors, numpy sort)."""

    def __init__(self, key_type):
        self.key_type = key_type
        self.np_dt = {"u32": np.uint32, "i32": np.int32, "u64": np.uint64, "i64": np.int64}[key_type]
        self.tdt = torch.int32 if self.np_dt().itemsize == 4 else torch.int64

    def empty(self, n):
        return torch.empty(n, dtype=self.tdt)

    def _view(self, t):
        return t.numpy().view(self.np_dt)

    def local_sort(self, keys, buf0, buf1):
        n = keys.numel()
        out = buf0[:n]
        self._view(out)[:] = np.sort(self._view(keys))
        return out

    def lower_bound(self, sorted_keys, queries):
        q = queries.astype(np.uint64).astype(self.np_dt) if self.np_dt in (np.uint32, np.uint64) else \
            queries.astype(np.uint64).view(np.int64).astype(self.np_dt) if self.np_dt == np.int64 else \
            queries.astype(np.uint32).view(np.int32)
        return np.searchsorted(self._view(sorted_keys), q, side="left").astype(np.int64)


def _make_shard(case, key_type, rank, world, n):
    dt = {"u32": np.uint32, "i32": np.int32, "u64": np.uint64, "i64": np.int64}[key_type]
    rng = np.random.default_rng(1000 + rank)
    info = np.iinfo(dt)
    if case == "uniform":
        return rng.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)
    if case == "zeros":
        return np.zeros(n, dtype=dt)
    if case == "maxes":  # every key is the largest value: the splitter's successor is 2^bits itself
        return np.full(n, info.max, dtype=dt)
    if case == "ends":   # only the smallest and the largest value, and one rank holds nothing
        return np.empty(0, dtype=dt) if rank == 0 else np.where(rng.integers(0, 2, size=n) == 0, info.min, info.max).astype(dt)
    if case == "few":  # heavy duplicates: ties must be split by count
        return rng.integers(0, 3, size=n).astype(dt)
    if case == "range":  # global range, sharded contiguously
        return (np.arange(rank * n, (rank + 1) * n, dtype=np.int64) + int(info.min)).astype(dt)
    if case == "inverted":
        return (np.arange(world * n - 1 - rank * n, world * n - 1 - (rank + 1) * n, -1, dtype=np.int64) + int(info.min)).astype(dt)
    if case == "skewed":  # everything on one rank is small, the other large, plus ragged sizes
        lo = rng.integers(0, 1000, size=n + 17 * rank).astype(dt)
        return lo if rank % 2 == 0 else (lo + dt(100000)).astype(dt)
    raise ValueError(case)


def _worker(rank, world, port, key_type, cases, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dsort = importlib.import_module("radix-sort_b200.dist")
    ops = NumpyOps(key_type)
    try:
        for case in cases:
            shard = _make_shard(case, key_type, rank, world, n)
            t = torch.from_numpy(shard.view(np.int32 if shard.itemsize == 4 else np.int64).copy())
            out = dsort.distributed_sort(t, ops, dist, key_type)
            np.save(os.path.join(out_dir, f"in_{case}_{rank}.npy"), shard)
            np.save(os.path.join(out_dir, f"out_{case}_{rank}.npy"), out.numpy().view(shard.dtype).copy())
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,key_type", [(2, "u32"), (3, "i32"), (2, "i64"), (2, "u64")])
def test_distributed_sort_host_logic(tmp_path, world, key_type):
    cases = ["uniform", "zeros", "few", "range", "inverted", "skewed", "maxes", "ends"]
    n = 5000
    mp.spawn(_worker, args=(world, _free_port(), key_type, cases, n, str(tmp_path)), nprocs=world, join=True)
    for case in cases:
        ins = [np.load(tmp_path / f"in_{case}_{r}.npy") for r in range(world)]
        outs = [np.load(tmp_path / f"out_{case}_{r}.npy") for r in range(world)]
        want = np.sort(np.concatenate(ins))
        got = np.concatenate(outs)
        assert np.array_equal(got, want), (case, key_type)          # contiguous global order across ranks
        total = len(want)
        for r in range(world):                                      # exact balance, whatever the distribution
            assert len(outs[r]) == ((r + 1) * total) // world - (r * total) // world, (case, r)
