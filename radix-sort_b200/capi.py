"""ctypes binding of include/rsb200.h (librsb200.so).  No torch types cross this boundary;
torch tensors are passed as raw device pointers (``tensor.data_ptr()``) by the callers."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "rsb200.h")
# RSB200_LIB: developer override to A/B a differently tuned build of the same library (tools/build_variants.sh)
_LIB_PATH = os.environ.get("RSB200_LIB") or os.path.join(HERE, "librsb200.so")

KEY_TYPES = {"u32": 0, "i32": 1, "u64": 2, "i64": 3, "u16": 4, "i16": 5, "f32": 6, "f64": 7, "f16": 8, "bf16": 9}
# host-side storage dtype of a key type (bf16 has no numpy dtype: its bit patterns travel as uint16)
_NP = {"u32": np.uint32, "i32": np.int32, "u64": np.uint64, "i64": np.int64, "u16": np.uint16, "i16": np.int16,
       "f32": np.float32, "f64": np.float64, "f16": np.float16, "bf16": np.uint16}
DATASET_KINDS = {"zeros": 0, "range": 1, "inverted": 2, "uniform": 3, "random": 4}
RS_WITH_PERMUTATION, RS_NO_TMA, RS_NO_PROFILE, RS_NO_KEY_BUFFERS, RS_NO_PASS_SKIP, RS_NO_ATOMIC_RANK, RS_NO_SMALL_PATH = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40

_lib = None


class RsError(RuntimeError):
    def __init__(self, status: int, detail: str):
        self.status = status
        super().__init__(f"rsb200: {status_string(status)} ({status}): {detail}")


def key_dtype(key_type: str):
    return np.dtype(_NP[key_type])


def library_path() -> str:
    return _LIB_PATH


def build_library(verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> radix-sort_b200/librsb200.so (in-tree)."""
    out = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("building librsb200.so failed")
    return _LIB_PATH


def declared_symbols() -> list[str]:
    """Every function include/rsb200.h declares (used by the CPU test that the library exports them)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", text)))


class _Stat(C.Structure):
    _fields_ = [("min", C.c_double), ("max", C.c_double), ("avg", C.c_double), ("sum", C.c_double), ("n", C.c_uint64)]


class _Runtimes(C.Structure):
    _fields_ = [("time_histo", _Stat), ("time_scan", _Stat), ("time_reorder", _Stat), ("time_paste", _Stat),
                ("time_total", _Stat)]


def load_library():
    """Load librsb200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise FileNotFoundError(f"{_LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
    lib = C.CDLL(_LIB_PATH)
    vp, u64, u32, i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    sig = {
        "rs_version": ([], i),
        "rs_status_string": ([i], C.c_char_p),
        "rs_last_error_string": ([], C.c_char_p),
        "rs_device_count": ([C.POINTER(i)], i),
        "rs_context_create": ([i, C.POINTER(vp)], i),
        "rs_context_destroy": ([vp], i),
        "rs_context_sm_count": ([vp, C.POINTER(i)], i),
        "rs_context_atomic_rank": ([vp, C.POINTER(i)], i),
        "rs_stream_create": ([vp, C.POINTER(vp)], i),
        "rs_stream_wrap": ([vp, vp, C.POINTER(vp)], i),
        "rs_stream_sync": ([vp], i),
        "rs_stream_destroy": ([vp], i),
        "rs_host_alloc": ([C.c_size_t, C.POINTER(vp)], i),
        "rs_host_free": ([vp], i),
        "rs_round_up_keys": ([u64], u64),
        "rs_sorter_create": ([vp, i, u64, u32, C.POINTER(vp)], i),
        "rs_sorter_destroy": ([vp], i),
        "rs_sorter_pad": ([vp, vp, u64, u64], i),
        "rs_sorter_upload": ([vp, vp, vp, vp, u64], i),
        "rs_sorter_sort": ([vp, vp, u64], i),
        "rs_sorter_download": ([vp, vp, vp, vp, u64], i),
        "rs_sorter_download_async": ([vp, vp, vp, vp, u64], i),
        "rs_sorter_download_histograms": ([vp, vp, vp, C.c_size_t, C.POINTER(i)], i),
        "rs_sorter_sort_device": ([vp, vp, vp, vp, vp, vp, vp, vp, u64, i, i, C.POINTER(i)], i),
        "rs_sorter_runtimes": ([vp, C.POINTER(_Runtimes)], i),
        "rs_sorter_last_skipped_passes": ([vp, vp, C.POINTER(i)], i),
        "rs_sorter_reset_runtimes": ([vp], i),
        "rs_sorter_launch_count": ([vp, C.POINTER(u64)], i),
        "rs_device_alloc": ([vp, C.c_size_t, C.POINTER(vp)], i),
        "rs_device_free": ([vp, vp], i),
        "rs_ipc_export": ([vp, vp, C.c_char_p], i),
        "rs_ipc_open": ([vp, C.c_char_p, C.POINTER(vp)], i),
        "rs_ipc_close": ([vp, vp], i),
        "rs_sorter_sort_segments_device": ([vp, vp, i, C.POINTER(vp), C.POINTER(u64), i, vp, vp, vp, i, i, C.POINTER(i)], i),
        "rs_merge_runs_device": ([vp, vp, i, i, C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), vp, u64], i),
        "rs_dataset_fill_device": ([vp, vp, i, i, u64, u64, u64, vp, u64], i),
        "rs_check_device": ([vp, vp, i, vp, u64, C.POINTER(u64 * 4)], i),
        "rs_check_permutation_device": ([vp, vp, i, vp, vp, vp, u64, C.POINTER(u64 * 2)], i),
        "rs_lower_bound_device": ([vp, vp, i, vp, u64, vp, u64, vp], i),
        "rs_dist_unique_id": ([C.c_char_p], i),
        "rs_dist_create": ([i, C.POINTER(i), i, u64, C.POINTER(vp)], i),
        "rs_dist_create_rank": ([vp, vp, i, u64, i, i, C.c_char_p, C.POINTER(vp)], i),
        "rs_dist_destroy": ([vp], i),
        "rs_dist_num_local": ([vp, C.POINTER(i)], i),
        "rs_dist_sort": ([vp, C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), C.POINTER(u64)], i),
        "rs_dist_sync": ([vp], i),
        "rs_dist_stage_ms": ([vp, i, C.POINTER(C.c_double * 6)], i),
    }
    for name, (args, res) in sig.items():
        f = getattr(lib, name)
        f.argtypes, f.restype = args, res
    _lib = lib
    return lib


def status_string(status: int) -> str:
    return load_library().rs_status_string(status).decode()


def _check(status: int):
    if status != 0:
        raise RsError(status, load_library().rs_last_error_string().decode())


def round_up_keys(n: int) -> int:
    return int(load_library().rs_round_up_keys(n))


def _addr(x):
    """numpy array / HostBuffer / int device pointer / None -> void*"""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(C.c_void_p)
    if isinstance(x, HostBuffer):
        return C.c_void_p(x.ptr)
    return C.c_void_p(int(x))


class Context:
    """CUDA device state; replaces ComputeState (reference Common/ComputeState.cpp:14-104)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.handle = C.c_void_p()
        _check(self.lib.rs_context_create(device, C.byref(self.handle)))
        self.device = device

    @staticmethod
    def device_count() -> int:
        n = C.c_int(0)
        load_library().rs_device_count(C.byref(n))
        return n.value

    @property
    def sm_count(self) -> int:
        n = C.c_int(0)
        _check(self.lib.rs_context_sm_count(self.handle, C.byref(n)))
        return n.value

    @property
    def atomic_rank(self) -> bool:
        n = C.c_int(0)
        _check(self.lib.rs_context_atomic_rank(self.handle, C.byref(n)))
        return bool(n.value)

    def close(self):
        if self.handle:
            self.lib.rs_context_destroy(self.handle)
            self.handle = C.c_void_p()

    # -- multi-GPU building blocks ---------------------------------------------------------------
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        _check(self.lib.rs_device_alloc(self.handle, nbytes, C.byref(p)))
        return p.value

    def device_free(self, ptr: int):
        _check(self.lib.rs_device_free(self.handle, C.c_void_p(ptr)))

    def ipc_export(self, ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        _check(self.lib.rs_ipc_export(self.handle, C.c_void_p(ptr), buf))
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        _check(self.lib.rs_ipc_open(self.handle, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def ipc_close(self, ptr: int):
        _check(self.lib.rs_ipc_close(self.handle, C.c_void_p(ptr)))

    def merge_runs(self, stream, key_type, jobs, d_splits, splits_capacity):
        """One level of pairwise merges: jobs = [(a_ptr, na, b_ptr, nb, out_ptr), ...] (<= 8), inputs may be peer memory."""
        k = len(jobs)
        ap = (C.c_void_p * k)(*[C.c_void_p(int(j[0])) for j in jobs])
        an = (C.c_uint64 * k)(*[int(j[1]) for j in jobs])
        bp = (C.c_void_p * k)(*[C.c_void_p(int(j[2])) for j in jobs])
        bn = (C.c_uint64 * k)(*[int(j[3]) for j in jobs])
        op = (C.c_void_p * k)(*[C.c_void_p(int(j[4])) for j in jobs])
        _check(self.lib.rs_merge_runs_device(self.handle, stream.handle, KEY_TYPES[key_type], k, ap, an, bp, bn, op,
                                             _addr(d_splits), splits_capacity))

    # -- device helpers -------------------------------------------------------------------------
    def fill_dataset(self, stream, key_type, kind, d_out, n, seed=0x5EED, global_offset=0, total_n=None):
        total_n = n if total_n is None else total_n
        _check(self.lib.rs_dataset_fill_device(self.handle, stream.handle, KEY_TYPES[key_type], DATASET_KINDS[kind],
                                               seed, global_offset, total_n, _addr(d_out), n))

    def check(self, stream, key_type, d_keys, n):
        """-> dict(unordered, sum, hash_sum, hash_xor) computed on the device (blocks)."""
        res = (C.c_uint64 * 4)()
        _check(self.lib.rs_check_device(self.handle, stream.handle, KEY_TYPES[key_type], _addr(d_keys), n, C.byref(res)))
        return {"unordered": res[0], "sum": res[1], "hash_sum": res[2], "hash_xor": res[3]}

    def check_permutation(self, stream, key_type, d_keys_in, d_keys_out, d_perm, n):
        res = (C.c_uint64 * 2)()
        _check(self.lib.rs_check_permutation_device(self.handle, stream.handle, KEY_TYPES[key_type], _addr(d_keys_in),
                                                    _addr(d_keys_out), _addr(d_perm), n, C.byref(res)))
        return {"mismatches": res[0], "unstable": res[1]}

    def lower_bound(self, stream, key_type, d_keys, n, d_queries, num_queries, d_out):
        _check(self.lib.rs_lower_bound_device(self.handle, stream.handle, KEY_TYPES[key_type], _addr(d_keys), n,
                                              _addr(d_queries), num_queries, _addr(d_out)))


class Stream:
    """A CUDA stream: replaces the in-order cl::CommandQueue the reference passes around."""

    def __init__(self, ctx: Context, cuda_stream: int | None = None):
        self.lib, self.ctx = ctx.lib, ctx
        self.handle = C.c_void_p()
        self.wrapped = cuda_stream          # the foreign cudaStream_t this object wraps, None for an own stream
        if cuda_stream is None:
            _check(self.lib.rs_stream_create(ctx.handle, C.byref(self.handle)))
        else:
            _check(self.lib.rs_stream_wrap(ctx.handle, C.c_void_p(cuda_stream), C.byref(self.handle)))

    def sync(self):
        _check(self.lib.rs_stream_sync(self.handle))

    def close(self):
        if self.handle:
            self.lib.rs_stream_destroy(self.handle)
            self.handle = C.c_void_p()


class HostBuffer:
    """Pinned host memory viewed as a numpy array (upload / download stage)."""

    def __init__(self, n: int, dtype):
        self.lib = load_library()
        self.dtype = np.dtype(dtype)
        p = C.c_void_p()
        _check(self.lib.rs_host_alloc(max(1, n) * self.dtype.itemsize, C.byref(p)))
        self.ptr = p.value
        self.n = n
        buf = (C.c_char * (max(1, n) * self.dtype.itemsize)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=n)

    def close(self):
        if self.ptr:
            self.array = None
            self.lib.rs_host_free(C.c_void_p(self.ptr))
            self.ptr = 0


class Sorter:
    """Binding of rs_sorter: what RadixSortGPU<T> + ComputeDeviceData<T> are in the reference
    (src/RadixSortGPU.h:38-124, src/ComputeDeviceData.h:12-28)."""

    def __init__(self, ctx: Context, key_type: str, max_keys: int, flags: int = 0):
        self.lib, self.ctx, self.key_type = ctx.lib, ctx, key_type
        self.dtype = key_dtype(key_type)
        self.max_keys, self.flags = max_keys, flags
        self.handle = C.c_void_p()
        _check(self.lib.rs_sorter_create(ctx.handle, KEY_TYPES[key_type], max_keys, flags, C.byref(self.handle)))

    def pad(self, stream, first_key, rounded):
        _check(self.lib.rs_sorter_pad(self.handle, stream.handle, first_key, rounded))

    def upload(self, stream, h_keys, h_perm=None, n=None):
        n = len(h_keys) if n is None else n
        _check(self.lib.rs_sorter_upload(self.handle, stream.handle, _addr(h_keys), _addr(h_perm), n))

    def sort(self, stream, n):
        _check(self.lib.rs_sorter_sort(self.handle, stream.handle, n))

    def download(self, stream, h_keys_out, h_perm_out=None, n=None):
        n = len(h_keys_out) if n is None else n
        _check(self.lib.rs_sorter_download(self.handle, stream.handle, _addr(h_keys_out), _addr(h_perm_out), n))

    def download_async(self, stream, h_keys_out, h_perm_out=None, n=None):
        """Enqueue the copies only; the (pinned) host buffers are valid after stream.sync()."""
        n = len(h_keys_out) if n is None else n
        _check(self.lib.rs_sorter_download_async(self.handle, stream.handle, _addr(h_keys_out), _addr(h_perm_out), n))

    def download_histograms(self, stream):
        buf = np.zeros(8 * 256, dtype=np.uint64)
        passes = C.c_int(0)
        _check(self.lib.rs_sorter_download_histograms(self.handle, stream.handle, _addr(buf), buf.size, C.byref(passes)))
        return buf[: passes.value * 256].reshape(passes.value, 256)

    def sort_device(self, stream, d_in, d_buf0, d_buf1, n, d_vin=None, d_vbuf0=None, d_vbuf1=None, begin_bit=0,
                    end_bit=None) -> int:
        """Device-pointer entry; returns the index (0/1) of the buffer that holds the result."""
        end_bit = self.dtype.itemsize * 8 if end_bit is None else end_bit
        idx = C.c_int(0)
        _check(self.lib.rs_sorter_sort_device(self.handle, stream.handle, _addr(d_in), _addr(d_buf0), _addr(d_buf1),
                                              _addr(d_vin), _addr(d_vbuf0), _addr(d_vbuf1), n, begin_bit, end_bit,
                                              C.byref(idx)))
        return idx.value

    def last_skipped_passes(self, stream) -> int:
        v = C.c_int(0)
        _check(self.lib.rs_sorter_last_skipped_passes(self.handle, stream.handle, C.byref(v)))
        return v.value

    def sort_segments_device(self, stream, seg_ptrs, seg_counts, first_segment, d_gathered, d_buf0, d_buf1,
                             begin_bit=0, end_bit=None) -> int:
        """Fused exchange + sort: pull the (possibly peer-memory) segments into d_gathered while histogramming, then
        sort; returns the index (0/1) of the buffer that holds the result (buf1 may alias d_gathered)."""
        end_bit = self.dtype.itemsize * 8 if end_bit is None else end_bit
        k = len(seg_ptrs)
        ptrs = (C.c_void_p * k)(*[C.c_void_p(int(p)) for p in seg_ptrs])
        cnts = (C.c_uint64 * k)(*[int(c) for c in seg_counts])
        idx = C.c_int(0)
        _check(self.lib.rs_sorter_sort_segments_device(self.handle, stream.handle, k, ptrs, cnts, first_segment,
                                                       _addr(d_gathered), _addr(d_buf0), _addr(d_buf1), begin_bit, end_bit,
                                                       C.byref(idx)))
        return idx.value

    def runtimes(self) -> dict:
        rt = _Runtimes()
        _check(self.lib.rs_sorter_runtimes(self.handle, C.byref(rt)))
        out = {}
        for name, _ in _Runtimes._fields_:
            st = getattr(rt, name)
            out[name] = {"min": st.min, "max": st.max, "avg": st.avg, "sum": st.sum, "n": st.n}
        return out

    def reset_runtimes(self):
        _check(self.lib.rs_sorter_reset_runtimes(self.handle))

    @property
    def launch_count(self) -> int:
        v = C.c_uint64(0)
        _check(self.lib.rs_sorter_launch_count(self.handle, C.byref(v)))
        return v.value

    def close(self):
        if self.handle:
            self.lib.rs_sorter_destroy(self.handle)
            self.handle = C.c_void_p()


DIST_STAGES = ("local_sort", "splitters", "merge_level_1", "merge_levels_local", "closing", "total")


def dist_unique_id() -> bytes:
    """ncclUniqueId for rs_dist_create_rank: create on one rank, hand the 128 bytes to the others."""
    buf = C.create_string_buffer(128)
    _check(load_library().rs_dist_unique_id(buf))
    return buf.raw


class DistSorter:
    """Binding of rs_dist: the multi-GPU sort of one NVSwitch box behind the C ABI (NCCL + peer memory inside).
    `DistSorter.rank(...)`: this process drives one GPU of `world`; `DistSorter.all_gpus(...)`: one process, all GPUs."""

    def __init__(self, handle, key_type: str, n_local: int):
        self.lib, self.handle, self.key_type, self.n_local_gpus = load_library(), handle, key_type, n_local

    @classmethod
    def rank(cls, ctx: Context, stream: Stream, key_type: str, max_keys_per_gpu: int, rank: int, world: int, unique_id: bytes):
        h = C.c_void_p()
        _check(ctx.lib.rs_dist_create_rank(ctx.handle, stream.handle, KEY_TYPES[key_type], max_keys_per_gpu, rank, world,
                                           C.create_string_buffer(unique_id, 128), C.byref(h)))
        return cls(h, key_type, 1)

    @classmethod
    def all_gpus(cls, key_type: str, max_keys_per_gpu: int, devices=None, ngpus=None):
        lib = load_library()
        n = len(devices) if devices is not None else ngpus
        devs = (C.c_int * n)(*devices) if devices is not None else None
        h = C.c_void_p()
        _check(lib.rs_dist_create(n, devs, KEY_TYPES[key_type], max_keys_per_gpu, C.byref(h)))
        return cls(h, key_type, n)

    def sort(self, shard_ptrs, counts):
        """shard_ptrs / counts: one entry per GPU this handle drives.  Returns [(device pointer, keys)] per GPU: this GPU's
        slice of the global order, in buffers owned by the handle (valid until the next sort); the tail of the work is
        stream-ordered (see rs_dist_sort)."""
        k = self.n_local_gpus
        ins = (C.c_void_p * k)(*[C.c_void_p(int(p)) if p else None for p in shard_ptrs])
        ns = (C.c_uint64 * k)(*[int(c) for c in counts])
        outs, n_out = (C.c_void_p * k)(), (C.c_uint64 * k)()
        _check(self.lib.rs_dist_sort(self.handle, ins, ns, outs, n_out))
        return [(outs[i] or 0, int(n_out[i])) for i in range(k)]

    def sync(self):
        _check(self.lib.rs_dist_sync(self.handle))

    def stage_ms(self, local_index: int = 0) -> dict:
        out = (C.c_double * 6)()
        _check(self.lib.rs_dist_stage_ms(self.handle, local_index, C.byref(out)))
        return dict(zip(DIST_STAGES, [float(x) for x in out]))

    def close(self):
        if self.handle:
            self.lib.rs_dist_destroy(self.handle)
            self.handle = C.c_void_p()
