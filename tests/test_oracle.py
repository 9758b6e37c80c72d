"""CPU tests: the oracle restatement (oracle/radix_oracle.c) against the reference's golden vectors,
against the reference's own code where it is built (oracle/_ref), and against numpy as the
independent ground truth (std::sort / std::stable_sort argsort)."""
import hashlib

import numpy as np
import pytest

from oracle import KEY_TYPES, Reference, have_reference, numpy_dtype

KINDS = ("zeros", "range", "inverted", "uniform", "random")
SEED = 0x5EED


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_random_dataset_golden_head(oracle):
    # first values of Random<uint32_t> computed from the reference header
    head = oracle.dataset("random", "u32", 8)
    assert head.tolist() == [2421477274, 811668573, 145020712, 106868501, 1537199182, 2398500640, 300868536, 942266821]
    assert int(oracle.dataset("random", "i32", 1)[0]) == -1873490022
    # 64-bit types hold the same 32-bit values zero-extended
    assert oracle.dataset("random", "u64", 8).tolist() == head.tolist()
    assert oracle.dataset("random", "i64", 8).tolist() == head.tolist()


@pytest.mark.parametrize("t", KEY_TYPES)
def test_datasets_match_golden(oracle, golden_small, golden_digests, t):
    assert np.array_equal(oracle.dataset("random", t, 4096), golden_small[f"random_head_{t}"])
    assert np.array_equal(oracle.dataset("uniform", t, 1024, SEED), golden_small[f"uniform_head_{t}"])
    for kind in KINDS:
        assert np.array_equal(oracle.dataset(kind, t, 1024, SEED), golden_small[f"in_{kind}_{t}_1024"]), kind
        for n in (1 << 16, 1 << 20):
            assert sha(oracle.dataset(kind, t, n, SEED)) == golden_digests["datasets"][f"{kind}_{t}_{n}"], (kind, n)


@pytest.mark.parametrize("t", KEY_TYPES)
def test_radix_sort_cpu_matches_golden(oracle, golden_small, golden_digests, t):
    for kind in KINDS:
        got = oracle.radix_sort_cpu(golden_small[f"in_{kind}_{t}_1024"], t)
        assert np.array_equal(got, golden_small[f"radixcpu_{kind}_{t}_1024"]), kind
        # N >= 2^10 on the five datasets: RadixSortCPU == std::sort
        assert np.array_equal(got, golden_small[f"stdsort_{kind}_{t}_1024"]), kind
        assert np.array_equal(got, np.sort(golden_small[f"in_{kind}_{t}_1024"])), kind
    for kind in KINDS:
        n = 1 << 16
        data = oracle.dataset(kind, t, n, SEED)
        got = oracle.radix_sort_cpu(data, t)
        assert sha(got) == golden_digests["radix_sort_cpu"][f"{kind}_{t}_{n}"], kind
        assert np.array_equal(got, np.sort(data)), kind


def test_radix_sort_cpu_2_20_config1(oracle, golden_digests):
    # BASELINE.json configs[0]: CRadixSortCPU on 2^20 uint32 keys, random distribution
    n = 1 << 20
    data = oracle.dataset("random", "u32", n)
    got = oracle.radix_sort_cpu(data, "u32")
    assert sha(got) == golden_digests["radix_sort_cpu"][f"random_u32_{n}"]
    assert np.array_equal(got, np.sort(data))


@pytest.mark.parametrize("t", KEY_TYPES)
def test_radix_sort_cpu_reproduces_reference_bugs(oracle, golden_small, t):
    # tiny inputs: the restatement returns what the reference returns, right or wrong
    wrong = 0
    for kind, n in (("inverted", 2), ("random", 3), ("random", 4), ("range", 7), ("random", 100)):
        data = golden_small[f"in_{kind}_{t}_{n}"]
        got = oracle.radix_sort_cpu(data, t)
        assert np.array_equal(got, golden_small[f"radixcpu_{kind}_{t}_{n}"]), (kind, n)
        wrong += int(not np.array_equal(got, np.sort(data)))
    if t in ("u32", "u64"):
        assert wrong >= 1  # InvertedRange at N = 2: max = 1 => zero digits sorted


@pytest.mark.parametrize("t", KEY_TYPES)
@pytest.mark.parametrize("kind", KINDS)
def test_opencl_path_is_stable_lsd_sort(oracle, t, kind):
    # the restated OpenCL path == std::sort for keys and == stable argsort once the payload moves
    for n in (1024, 5 * 1024):
        data = oracle.dataset(kind, t, n, SEED)
        keys, perm, histo, globsum = oracle.opencl_path(data, t, with_perm=True)
        assert np.array_equal(keys, np.sort(data))
        assert np.array_equal(perm, np.argsort(data, kind="stable").astype(np.uint32))
        assert int(histo.max()) <= n and histo.size == 16384 and globsum.size == 512
        # last pass's scanned table: last group/item/digit offset + its count == n
        keys2, none, _, _ = oracle.opencl_path(data, t, with_perm=False)
        assert none is None and np.array_equal(keys2, keys)


def test_opencl_path_rejects_unpadded(oracle):
    import ctypes as C
    a = np.zeros(1000, dtype=np.uint32)
    rc = oracle.lib.oracle_opencl_path_u32(a.ctypes.data_as(C.c_void_p), None, a.size, None, None)
    assert rc == -2  # the reference requires n % 1024 == 0 (RadixSortGPU::Resize)


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("t", KEY_TYPES)
def test_restatement_equals_reference_code(oracle, t):
    ref = Reference()
    assert ref.lib.ref_num_bins_32() == 8 and ref.lib.ref_num_bins_64() == 16
    rng = np.random.default_rng(1234)
    for kind in ("zeros", "range", "inverted", "random"):
        for n in (2, 3, 5, 64, 1000, 1024, 4096, 100003):
            a, b = oracle.dataset(kind, t, n), ref.dataset(kind, t, n)
            assert np.array_equal(a, b), (kind, n)
            assert np.array_equal(oracle.radix_sort_cpu(a, t), ref.radix_sort_cpu(b, t)), (kind, n)
    for n, seed in ((2, 1), (1000, SEED), (65536, 2**40 + 17)):
        a, b = oracle.dataset("uniform", t, n, seed), ref.uniform_seeded(t, n, seed)
        assert np.array_equal(a, b), (n, seed)
        assert np.array_equal(oracle.radix_sort_cpu(a, t), ref.radix_sort_cpu(b, t))
    # arbitrary inputs incl. small magnitudes (exercise the digit-count rule)
    info = np.iinfo(numpy_dtype(t))
    for hi in (2, 9, 65, 4097, 2**31 - 1):
        a = rng.integers(0, hi, size=3000, dtype=np.int64).astype(numpy_dtype(t))
        assert np.array_equal(oracle.radix_sort_cpu(a, t), ref.radix_sort_cpu(a, t)), hi
    if info.min < 0:
        a = rng.integers(-1000, 1000, size=3000, dtype=np.int64).astype(numpy_dtype(t))
        assert np.array_equal(oracle.radix_sort_cpu(a, t), ref.radix_sort_cpu(a, t))
