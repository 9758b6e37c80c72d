"""torch entry of the sort: ``sort(tensor) -> (values, indices)`` on device tensors, zero copy.

SURVEY.md 8(f).3.  What a PyTorch user of the reference's path would call: the key tensor goes through
``rs_sorter_sort_device`` (include/rsb200.h) on torch's current stream; torch only owns the memory.  The result
tensors are the sorter's ping-pong buffers themselves (no copy); ``indices`` is the stable sorting permutation
(``torch.sort(stable=True)`` semantics), stored as int32 bit patterns of the library's uint32 payload, so inputs
are limited to 2^31 - 1 keys when indices are requested.

Every dtype is a key type of the C ABI (include/rsb200.h rs_key_type): (u)int16/32/64 and -- additive, the
reference excludes floats (doc/doc.tex:64) -- float16 / bfloat16 / float32 / float64.  Floating-point keys are
ordered by value inside the kernels' digit extraction (no conversion passes around the sort):
-inf < ... < -0.0 < +0.0 < ... < +inf < NaN, every NaN last and in input order, as torch.sort places them.
16-bit keys take two onesweep passes, 32-bit keys four, 64-bit keys eight.

There is no CPU fallback: a tensor that is not on a CUDA device, or a missing ``librsb200.so``, raises.
The reference has no counterpart (its user entry is examples/basic_sort/basic_sort.cpp:24-139, host buffers only).
"""
from __future__ import annotations

import threading

from . import capi

_KEY_OF_DTYPE = {"torch.int32": "i32", "torch.int64": "i64", "torch.uint32": "u32", "torch.uint64": "u64",
                 "torch.int16": "i16", "torch.uint16": "u16", "torch.float32": "f32", "torch.float64": "f64",
                 "torch.float16": "f16", "torch.bfloat16": "bf16"}


class _NativeBackend:
    """The product backend: librsb200.so through the ctypes binding."""

    def __init__(self, device_index: int):
        self.ctx = capi.Context(device_index)
        self._streams = {}

    def stream(self, cuda_stream: int):
        st = self._streams.get(cuda_stream)
        if st is None:
            st = self._streams[cuda_stream] = capi.Stream(self.ctx, cuda_stream)
        return st

    def make_sorter(self, key_type: str, capacity: int):
        return capi.Sorter(self.ctx, key_type, capacity, capi.RS_NO_KEY_BUFFERS | capi.RS_NO_PROFILE)


class TorchSorter:
    """Per-device sorter cache + scratch reuse for ``sort``.  ``backend`` is injectable for the host-logic tests."""

    def __init__(self, device_index: int = 0, backend=None):
        self.device_index = device_index
        self.backend = backend if backend is not None else _NativeBackend(device_index)
        self._sorters = {}      # (key type, cuda stream) -> (sorter, capacity): scratch is never shared across streams
        self._lock = threading.Lock()

    def _sorter(self, key_type: str, cuda_stream: int, n: int):
        entry = self._sorters.get((key_type, cuda_stream))
        if entry is None or entry[1] < n:
            capacity = max(n, 2 * entry[1] if entry is not None else 1 << 20)
            if entry is not None:
                self.backend.stream(cuda_stream).sync()   # the old scratch may still be in use by an earlier sort
                entry[0].close()
            entry = self._sorters[(key_type, cuda_stream)] = (self.backend.make_sorter(key_type, capacity), capacity)
        return entry[0]

    def sort(self, keys, return_indices: bool = True, end_bit: int | None = None):
        import torch
        key_type = _KEY_OF_DTYPE.get(str(keys.dtype))
        if key_type is None:
            raise TypeError(f"unsupported key dtype {keys.dtype}: (u)int16/32/64, float16, bfloat16, float32, float64")
        if keys.dim() != 1 or not keys.is_contiguous():
            raise ValueError("keys must be a contiguous 1-D tensor")
        if self.backend.__class__ is _NativeBackend and not keys.is_cuda:
            raise ValueError("keys must live on a CUDA device (there is no CPU path)")
        n = keys.numel()
        if return_indices and n > 0x7FFFFFFF:
            raise ValueError("indices are a 32-bit payload: at most 2^31 - 1 keys")
        if n == 0:
            return keys.clone(), (torch.empty(0, dtype=torch.int32, device=keys.device) if return_indices else None)
        # torch's CURRENT stream: the temporaries below go back to torch's caching allocator when this function returns,
        # which is only safe because allocator reuse is ordered on the stream the sort was enqueued on
        cuda_stream = torch.cuda.current_stream(keys.device).cuda_stream if keys.is_cuda else 0
        with self._lock:
            sorter = self._sorter(key_type, cuda_stream, n)
            stream = self.backend.stream(cuda_stream)
            kbuf = [torch.empty_like(keys), torch.empty_like(keys)]
            if return_indices:
                vin = torch.arange(n, dtype=torch.int32, device=keys.device)
                vbuf = [torch.empty_like(vin), torch.empty_like(vin)]
                r = sorter.sort_device(stream, keys.data_ptr(), kbuf[0].data_ptr(), kbuf[1].data_ptr(), n,
                                       vin.data_ptr(), vbuf[0].data_ptr(), vbuf[1].data_ptr(), 0, end_bit)
                return kbuf[r], vbuf[r]
            r = sorter.sort_device(stream, keys.data_ptr(), kbuf[0].data_ptr(), kbuf[1].data_ptr(), n, None, None, None, 0, end_bit)
            return kbuf[r], None

    def close(self):
        for sorter, _ in self._sorters.values():
            sorter.close()
        self._sorters = {}


_per_device: dict[int, TorchSorter] = {}


def sort(keys, return_indices: bool = True):
    """Ascending sort of a 1-D CUDA tensor of (u)int16/32/64 or float16/bfloat16/float32/float64 keys.

    Returns ``(values, indices)``; ``indices`` (int32) is the stable permutation with ``keys[indices] == values``,
    or ``None`` when ``return_indices`` is False (keys-only sorts run ~1.7x faster).  The work is enqueued on
    torch's current stream of ``keys.device`` and is asynchronous like any torch op; the input is not modified.
    """
    if not keys.is_cuda:
        raise ValueError("keys must live on a CUDA device (there is no CPU path)")
    index = keys.device.index if keys.device.index is not None else 0
    ts = _per_device.get(index)
    if ts is None:
        ts = _per_device[index] = TorchSorter(index)
    return ts.sort(keys, return_indices)
