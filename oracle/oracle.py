"""ctypes doors onto oracle/liboracle.so (the C restatement, radix_oracle.c) and, where it was
built, oracle/_ref/libref_radix.so (the reference's own CPU code, ref_shim.cpp).

TEST INFRASTRUCTURE ONLY: the checker / the timed CPU baseline, never the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libref_radix.so")

KEY_TYPES = ("u32", "i32", "u64", "i64")
# kind ordinals shared by radix_oracle.c, ref_shim.cpp and include/rsb200.h (rs_dataset_kind)
DATASETS = {"zeros": 0, "range": 1, "inverted": 2, "uniform": 3, "random": 4}
# Dataset<T>::name() strings of the reference (src/Dataset.h:40,67,76,49,58)
DATASET_NAMES = {"zeros": "Zeros", "range": "Range", "inverted": "Inverted Range",
                 "uniform": "Random Uniform", "random": "Random Random"}
_DT = {"u32": np.uint32, "i32": np.int32, "u64": np.uint64, "i64": np.int64}


def numpy_dtype(key_type: str):
    return np.dtype(_DT[key_type])


def build(verbose: bool = False) -> None:
    """Compile the restatement (always) and the reference shim (when /root/reference exists)."""
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_reference() -> bool:
    return os.path.exists(LIB_REF)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """The C restatement of the reference (radix_oracle.c)."""

    def __init__(self):
        if not os.path.exists(LIB_ORACLE):
            build()
        self.lib = C.CDLL(LIB_ORACLE)
        for t in KEY_TYPES:
            f = getattr(self.lib, f"oracle_radix_sort_cpu_{t}")
            f.argtypes, f.restype = [C.c_void_p, C.c_size_t], C.c_int
            f = getattr(self.lib, f"oracle_opencl_path_{t}")
            f.argtypes, f.restype = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p], C.c_int
            f = getattr(self.lib, f"oracle_dataset_{t}")
            f.argtypes, f.restype = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint64], C.c_int

    def radix_sort_cpu(self, keys: np.ndarray, key_type: str) -> np.ndarray:
        """RadixSortCPU<T>::sort (src/CRadixSortCPU.h:57) on a copy."""
        a = np.ascontiguousarray(keys, dtype=_DT[key_type]).copy()
        rc = getattr(self.lib, f"oracle_radix_sort_cpu_{key_type}")(_ptr(a), a.size)
        assert rc == 0, rc
        return a

    def opencl_path(self, keys: np.ndarray, key_type: str, with_perm: bool = False):
        """The reference's OpenCL GPU path emulated on the host; len(keys) % 1024 == 0.
        Returns (sorted_keys, perm_or_None, histograms_last_pass, globsum_last_pass)."""
        a = np.ascontiguousarray(keys, dtype=_DT[key_type]).copy()
        perm = np.arange(a.size, dtype=np.uint32) if with_perm else None
        histo = np.zeros(16 * 1024, dtype=np.uint32)
        globsum = np.zeros(512, dtype=np.uint32)
        rc = getattr(self.lib, f"oracle_opencl_path_{key_type}")(
            _ptr(a), _ptr(perm) if with_perm else None, a.size, _ptr(histo), _ptr(globsum))
        assert rc == 0, rc
        return a, perm, histo, globsum

    def dataset(self, kind: str, key_type: str, n: int, seed: int = 0x5EED) -> np.ndarray:
        """Dataset generators of src/Dataset.h:84-137 (`seed` replaces the clock of 'uniform')."""
        a = np.empty(n, dtype=_DT[key_type])
        rc = getattr(self.lib, f"oracle_dataset_{key_type}")(DATASETS[kind], _ptr(a), n, seed)
        assert rc == 0, rc
        return a


class Reference:
    """The reference's own CPU code (oracle/_ref/libref_radix.so). Absent => constructor raises."""

    def __init__(self):
        if not os.path.exists(LIB_REF):
            raise FileNotFoundError(LIB_REF)
        self.lib = C.CDLL(LIB_REF)
        for t in KEY_TYPES:
            for name in ("ref_radix_sort_cpu", "ref_std_sort"):
                f = getattr(self.lib, f"{name}_{t}")
                f.argtypes, f.restype = [C.c_void_p, C.c_size_t], C.c_int
            f = getattr(self.lib, f"ref_dataset_{t}")
            f.argtypes, f.restype = [C.c_int, C.c_void_p, C.c_size_t], C.c_int
            f = getattr(self.lib, f"ref_uniform_seeded_{t}")
            f.argtypes, f.restype = [C.c_void_p, C.c_size_t, C.c_uint64], C.c_int
            for name in ("ref_time_radix_sort_cpu_ms", "ref_time_std_sort_ms"):
                f = getattr(self.lib, f"{name}_{t}")
                f.argtypes, f.restype = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p], C.c_double
        self.lib.ref_num_bins_32.restype = C.c_uint
        self.lib.ref_num_bins_64.restype = C.c_uint
        self.lib.ref_num_performance_iterations.restype = C.c_uint

    def radix_sort_cpu(self, keys, key_type):
        a = np.ascontiguousarray(keys, dtype=_DT[key_type]).copy()
        assert getattr(self.lib, f"ref_radix_sort_cpu_{key_type}")(_ptr(a), a.size) == 0
        return a

    def std_sort(self, keys, key_type):
        a = np.ascontiguousarray(keys, dtype=_DT[key_type]).copy()
        assert getattr(self.lib, f"ref_std_sort_{key_type}")(_ptr(a), a.size) == 0
        return a

    def dataset(self, kind, key_type, n):
        a = np.empty(n, dtype=_DT[key_type])
        assert getattr(self.lib, f"ref_dataset_{key_type}")(DATASETS[kind], _ptr(a), n) == 0
        return a

    def uniform_seeded(self, key_type, n, seed):
        a = np.empty(n, dtype=_DT[key_type])
        assert getattr(self.lib, f"ref_uniform_seeded_{key_type}")(_ptr(a), n, seed) == 0
        return a

    def time_radix_sort_cpu_ms(self, keys, key_type, iterations):
        a = np.ascontiguousarray(keys, dtype=_DT[key_type])
        scratch = np.empty_like(a)
        return getattr(self.lib, f"ref_time_radix_sort_cpu_ms_{key_type}")(_ptr(a), a.size, iterations, _ptr(scratch))

    def time_std_sort_ms(self, keys, key_type, iterations):
        a = np.ascontiguousarray(keys, dtype=_DT[key_type])
        scratch = np.empty_like(a)
        return getattr(self.lib, f"ref_time_std_sort_ms_{key_type}")(_ptr(a), a.size, iterations, _ptr(scratch))
