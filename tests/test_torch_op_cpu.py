"""Host logic of the torch entry (radix-sort_b200/torch_op.py) with a numpy test double as the backend: argument
checks, sorter cache growth, which ping-pong buffer is returned.  The product backend (librsb200.so) is exercised by
tests/test_gpu_parity.py::test_torch_sort_operator on the GPU."""
import ctypes
import importlib

import numpy as np
import pytest
import torch


class _FakeStream:
    def sync(self):
        pass


class _FakeSorter:
    """Stands in for capi.Sorter: stable sort with numpy on the raw host pointers it is given; result in buffer 1."""
    created = []

    def __init__(self, key_type, capacity):
        self.np_dt = {"u32": np.uint32, "i32": np.int32, "u64": np.uint64, "i64": np.int64, "u16": np.uint16, "i16": np.int16,
                      "f32": np.float32, "f64": np.float64, "f16": np.float16, "bf16": np.uint16}[key_type]
        self.key_type = key_type
        self.capacity, self.closed = capacity, False
        _FakeSorter.created.append(self)

    def _view(self, ptr, n, dt):
        return np.ctypeslib.as_array((ctypes.c_uint8 * (n * np.dtype(dt).itemsize)).from_address(ptr)).view(dt)

    def sort_device(self, stream, d_in, d_buf0, d_buf1, n, d_vin=None, d_vbuf0=None, d_vbuf1=None, begin_bit=0, end_bit=None):
        assert not self.closed and n <= self.capacity and begin_bit == 0
        keys = self._view(d_in, n, self.np_dt)
        bits = self.np_dt().itemsize * 8 if end_bit is None else end_bit
        digits = keys if end_bit is None else (keys.astype(np.int64) & ((1 << bits) - 1))   # only bits [0, end_bit) order the keys
        if self.key_type[0] == "f" or self.key_type == "bf16":
            # the library's float order (include/rsb200.h): by value, -0.0 before +0.0, every NaN last
            u = keys.view({2: np.uint16, 4: np.uint32, 8: np.uint64}[keys.itemsize]).astype(np.uint64)
            nbits = keys.itemsize * 8
            sign, allb = np.uint64(1 << (nbits - 1)), np.uint64((1 << nbits) - 1)
            inf = np.uint64({"f16": 0x7c00, "bf16": 0x7f80, "f32": 0x7f800000, "f64": 0x7ff0000000000000}[self.key_type])
            digits = np.where((u & sign) != 0, u ^ allb, u | sign)
            digits = np.where((u & (allb ^ sign)) > inf, allb, digits)
        order = np.argsort(digits, kind="stable")
        self._view(d_buf1, n, self.np_dt)[:] = keys[order]
        self._view(d_buf0, n, self.np_dt)[:] = 0
        if d_vin is not None:
            self._view(d_vbuf1, n, np.int32)[:] = self._view(d_vin, n, np.int32)[order]
        return 1

    def close(self):
        self.closed = True


class _FakeBackend:
    def stream(self, cuda_stream):
        return _FakeStream()

    def make_sorter(self, key_type, capacity):
        return _FakeSorter(key_type, capacity)


@pytest.fixture()
def op():
    return importlib.import_module("radix-sort_b200.torch_op")


def test_sort_values_and_stable_indices(op):
    ts = op.TorchSorter(0, backend=_FakeBackend())
    rng = np.random.default_rng(0)
    for dt in (torch.int32, torch.int64):
        keys = torch.from_numpy(rng.integers(-5, 5, size=1000)).to(dt)
        before = keys.clone()
        values, indices = ts.sort(keys)
        ref_v, ref_i = torch.sort(keys, stable=True)
        assert torch.equal(values, ref_v) and torch.equal(indices.long(), ref_i) and indices.dtype == torch.int32
        assert torch.equal(keys, before)
        values2, none = ts.sort(keys, return_indices=False)
        assert none is None and torch.equal(values2, ref_v)
    ts.close()


def test_sorter_cache_grows_and_is_reused(op):
    _FakeSorter.created.clear()
    ts = op.TorchSorter(0, backend=_FakeBackend())
    ts.sort(torch.arange(10, dtype=torch.int32))
    ts.sort(torch.arange(100, dtype=torch.int32))
    assert len(_FakeSorter.created) == 1 and _FakeSorter.created[0].capacity >= 1 << 20
    big = (1 << 20) + 5
    ts.sort(torch.zeros(big, dtype=torch.int32), return_indices=False)
    assert len(_FakeSorter.created) == 2 and _FakeSorter.created[0].closed and _FakeSorter.created[1].capacity >= big
    ts.sort(torch.arange(7, dtype=torch.int64))
    assert len(_FakeSorter.created) == 3   # another key type, another sorter
    ts.close()
    assert all(s.closed for s in _FakeSorter.created)


def test_float_keys_total_order(op):
    ts = op.TorchSorter(0, backend=_FakeBackend())
    rng = np.random.default_rng(3)
    for dt, np_dt in ((torch.float32, np.float32), (torch.float64, np.float64)):
        host = rng.standard_normal(5000).astype(np_dt) * np_dt(1e3)
        host[:8] = [0.0, -0.0, np.inf, -np.inf, 1.0, -1.0, np.finfo(np_dt).tiny, -np.finfo(np_dt).max]
        keys = torch.from_numpy(host.copy())
        values, indices = ts.sort(keys)
        assert values.dtype == dt
        got = values.numpy()
        assert np.array_equal(np.sort(host), got)                      # no NaNs here: numeric order == total order
        assert np.array_equal(host[indices.numpy()], got)
        z = np.flatnonzero(got == 0.0)
        assert np.signbit(got[z[0]]) and not np.signbit(got[z[-1]])    # -0.0 before +0.0
        assert torch.equal(keys, torch.from_numpy(host))               # input untouched
        with_nan = torch.tensor([1.0, float("nan"), -2.0, float("inf"), -float("nan")], dtype=dt)
        v, idx = ts.sort(with_nan)
        assert torch.isnan(v[-1]) and torch.isnan(v[-2]) and v[0].item() == -2.0 and torch.isinf(v[-3])
        assert idx.tolist() == [2, 0, 3, 1, 4]                         # NaNs of either sign last, in input order
    ts.close()


def test_half_precision_keys(op):
    ts = op.TorchSorter(0, backend=_FakeBackend())
    for dt in (torch.float16, torch.bfloat16):
        keys = torch.tensor([3.5, -1.0, float("inf"), 0.0, -0.0, float("nan"), -float("inf"), 2.0], dtype=dt)
        values, indices = ts.sort(keys)
        assert values.dtype == dt and indices.tolist() == [6, 1, 4, 3, 7, 0, 2, 5]
    ts.close()


def test_16_bit_keys_two_passes(op):
    ts = op.TorchSorter(0, backend=_FakeBackend())
    rng = np.random.default_rng(4)
    for dt, np_dt in ((torch.int16, np.int16), (torch.uint16, np.uint16)):
        info = np.iinfo(np_dt)
        host = rng.integers(info.min, info.max, size=20_000, dtype=np_dt, endpoint=True)
        host[:2] = [info.min, info.max]
        keys = torch.from_numpy(host.view(np.int16).copy()).view(dt)
        values, indices = ts.sort(keys)
        assert values.dtype == dt
        order = np.argsort(host, kind="stable")
        assert np.array_equal(values.view(torch.int16).numpy().view(np_dt), host[order])
        assert np.array_equal(indices.numpy().astype(np.int64), order)
    ts.close()


def test_argument_checks(op):
    ts = op.TorchSorter(0, backend=_FakeBackend())
    with pytest.raises(TypeError):
        ts.sort(torch.zeros(4, dtype=torch.complex64))
    with pytest.raises(ValueError):
        ts.sort(torch.zeros((2, 2), dtype=torch.int32))
    with pytest.raises(ValueError):
        ts.sort(torch.arange(10, dtype=torch.int32)[::2])
    v, i = ts.sort(torch.empty(0, dtype=torch.int64))
    assert v.numel() == 0 and i.numel() == 0
    with pytest.raises(ValueError):       # the module-level entry is the product path: CUDA tensors only, no CPU fallback
        op.sort(torch.arange(4, dtype=torch.int32))
