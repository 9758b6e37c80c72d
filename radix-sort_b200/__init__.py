"""radix-sort_b200 -- B200-native drop-in for the GPU sort path of gyatskov/radix-sort.

The product is the C-ABI shared library ``librsb200.so`` (``include/rsb200.h``, built from
``csrc/``) and the C++ facade in ``cpp/`` that keeps the reference's class names
(``RadixSortGPU<T>``, ``ComputeDeviceData<T>``, ``CRadixSortTask<T>``, ``IComputeTask`` ...).
This Python package is the thin ctypes binding the tests and ``bench.py`` drive the library
through, plus the multi-GPU host logic (``dist``: one process per GPU over ``torch.distributed``) and the torch
entry ``torch_op.sort(tensor) -> (values, indices)``.

The directory name contains a hyphen, so import it with
``importlib.import_module("radix-sort_b200")`` (see ``tests/conftest.py``).
There is no CPU fallback: loading fails loudly when ``librsb200.so`` is missing.
"""
from .capi import (  # noqa: F401
    KEY_TYPES, DATASET_KINDS, RS_WITH_PERMUTATION, RS_NO_TMA, RS_NO_PROFILE, RS_NO_KEY_BUFFERS, RS_NO_PASS_SKIP, RS_NO_ATOMIC_RANK, RS_NO_SMALL_PATH,
    RsError, Context, Stream, Sorter, HostBuffer, build_library, library_path, load_library,
    declared_symbols, round_up_keys, status_string, key_dtype, DistSorter, dist_unique_id, DIST_STAGES,
)
