"""CPU test of the splitter planner behind rs_dist_sort: tests/cpp/libdistplan_check.so drives the SAME planning
functions the device kernels call (radix-sort_b200/csrc/rs_dist_plan.h) over emulated ranks.  The cuts must give every
rank exactly its share of the global order, for the distributions the gloo tests of the Python planner use, and they
must equal the Python planner's cuts (radix-sort_b200/dist.py:plan_exchange) computed through an in-process stand-in for
torch.distributed."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "cpp", "libdistplan_check.so")
_DT = {"u32": np.uint32, "i32": np.int32, "u64": np.uint64, "i64": np.int64}


@pytest.fixture(scope="module")
def planner():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "libdistplan_check.so"], check=True, capture_output=True)
    lib = C.CDLL(LIB)
    lib.distplan_emulate.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_void_p]
    lib.distplan_emulate.restype = C.c_int
    return lib


def _shards(case, key_type, world, n):
    dt = _DT[key_type]
    info = np.iinfo(dt)
    out = []
    for rank in range(world):
        rng = np.random.default_rng(1000 + rank)
        if case == "uniform":
            a = rng.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)
        elif case == "zeros":
            a = np.zeros(n, dtype=dt)
        elif case == "maxes":
            a = np.full(n, info.max, dtype=dt)
        elif case == "mins":
            a = np.full(n, info.min, dtype=dt)
        elif case == "ends":
            a = np.empty(0, dtype=dt) if rank == 0 else np.where(rng.integers(0, 2, size=n) == 0, info.min, info.max).astype(dt)
        elif case == "few":
            a = rng.integers(0, 3, size=n).astype(dt)
        elif case == "range":
            a = (np.arange(rank * n, (rank + 1) * n, dtype=np.int64) + int(info.min)).astype(dt)
        elif case == "inverted":
            a = (np.arange(world * n - 1 - rank * n, world * n - 1 - (rank + 1) * n, -1, dtype=np.int64) + int(info.min)).astype(dt)
        elif case == "skewed":
            lo = rng.integers(0, 1000, size=n + 17 * rank).astype(dt)
            a = lo if rank % 2 == 0 else (lo + dt(100000)).astype(dt)
        elif case == "empty":
            a = np.empty(0, dtype=dt)
        else:
            raise ValueError(case)
        out.append(np.sort(a))
    return out


def _emulate(planner, shards, key_type):
    world = len(shards)
    ptrs = (C.c_void_p * world)(*[s.ctypes.data if s.size else None for s in shards])
    ns = (C.c_uint64 * world)(*[s.size for s in shards])
    cuts = np.zeros((world, world + 1), dtype=np.uint64)
    rc = planner.distplan_emulate(shards[0].itemsize, int(key_type[0] == "i"), world, ptrs, ns, cuts.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return cuts.astype(np.int64)


CASES = ["uniform", "zeros", "few", "range", "inverted", "skewed", "maxes", "mins", "ends", "empty"]


@pytest.mark.parametrize("key_type", ["u32", "i32", "u64", "i64"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_cuts_give_every_rank_its_share_of_the_global_order(planner, key_type, world):
    for case in CASES:
        shards = _shards(case, key_type, world, 3000)
        cuts = _emulate(planner, shards, key_type)
        total = sum(s.size for s in shards)
        want = np.sort(np.concatenate(shards))
        got = []
        for r in range(world):
            assert cuts[r, 0] == 0 and cuts[r, -1] == shards[r].size and (np.diff(cuts[r]) >= 0).all(), (case, r)
        for dst in range(world):
            runs = [shards[src][cuts[src, dst]:cuts[src, dst + 1]] for src in range(world)]
            mine = np.sort(np.concatenate(runs))          # what the merge tree produces from the runs
            assert mine.size == ((dst + 1) * total) // world - (dst * total) // world, (case, key_type, dst)   # exact balance
            got.append(mine)
        assert np.array_equal(np.concatenate(got), want), (case, key_type)


class _InProcessDist:
    """torch.distributed stand-in that runs the P ranks of plan_exchange in lockstep threads."""

    def __init__(self, world):
        import threading
        self.world, self.barrier, self.lock = world, threading.Barrier(world), threading.Lock()
        self.slots, self.local = {}, threading.local()

    class ReduceOp:
        SUM = "sum"

    def get_world_size(self, group=None):
        return self.world

    def get_rank(self, group=None):
        return self.local.rank

    def _exchange(self, value):
        self.barrier.wait()
        self.slots[self.local.rank] = value.clone()
        self.barrier.wait()
        vals = [self.slots[r] for r in range(self.world)]
        self.barrier.wait()
        return vals

    def all_reduce(self, t, op=None, group=None):
        vals = self._exchange(t)
        t.copy_(sum(vals[1:], vals[0].clone()))

    def all_gather(self, outs, t, group=None):
        for o, v in zip(outs, self._exchange(t)):
            o.copy_(v)


@pytest.mark.parametrize("key_type,world", [("u32", 2), ("i64", 3), ("u64", 4), ("i32", 8)])
def test_cuts_equal_the_python_planner(planner, key_type, world):
    import threading

    import torch
    dsort = importlib.import_module("radix-sort_b200.dist")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    tdc = importlib.import_module("test_dist_cpu")
    for case in ("uniform", "few", "skewed", "maxes", "ends", "range"):
        shards = _shards(case, key_type, world, 2000)
        want = _emulate(planner, shards, key_type)
        fake = _InProcessDist(world)
        got = [None] * world
        errors = []

        def run(rank):
            try:
                fake.local.rank = rank
                ops = tdc.NumpyOps(key_type)
                t = torch.from_numpy(shards[rank].view(np.int32 if shards[rank].itemsize == 4 else np.int64).copy())
                got[rank], _ = dsort.plan_exchange(t, ops, fake, key_type, need_recv_counts=False)
            except Exception as e:  # noqa: BLE001
                errors.append(e)
                fake.barrier.abort()

        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        assert not errors, errors
        assert np.array_equal(np.stack(got), want), (case, key_type)
