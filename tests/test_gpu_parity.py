"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle, the
committed golden vectors and numpy (= std::sort / std::stable_sort argsort), bit-exact."""
import os
import numpy as np
import pytest

from oracle import KEY_TYPES, numpy_dtype

pytestmark = pytest.mark.gpu

KINDS = ("zeros", "range", "inverted", "uniform", "random")
SEED = 0x5EED
SIZES = (1, 2, 3, 31, 33, 100, 1023, 1024, 1025, 4097, 6144, 8192, 8193, 12289, 65536, (1 << 20) + 12345)


def host_sort(rsb, gpu, keys, t, with_perm=False, flags=0, perm_in=None):
    """upload -> sort -> download through rs_sorter_* (the RadixSortGPU call sequence)."""
    ctx, st = gpu
    n = len(keys)
    s = rsb.Sorter(ctx, t, max(n, 1), flags | (rsb.RS_WITH_PERMUTATION if with_perm else 0))
    try:
        s.upload(st, keys, perm_in, n)
        s.sort(st, n)
        out = np.empty(n, dtype=numpy_dtype(t))
        perm = np.empty(n, dtype=np.uint32) if with_perm else None
        s.download(st, out, perm, n)
        return out, perm, s.download_histograms(st) if n else None
    finally:
        s.close()


def test_library_is_the_cuda_one(rsb, gpu):
    ctx, _ = gpu
    assert ctx.sm_count == 148
    assert rsb.Context.device_count() >= 1


@pytest.mark.parametrize("t", KEY_TYPES)
@pytest.mark.parametrize("kind", KINDS)
def test_keys_only_bit_exact(rsb, gpu, oracle, t, kind):
    for n in SIZES:
        data = oracle.dataset(kind, t, n, SEED)
        got, _, hist = host_sort(rsb, gpu, data, t)
        assert np.array_equal(got, np.sort(data)), (t, kind, n)
        if n % 1024 == 0:
            # the reference's OpenCL path restated on the host (needs the padded length)
            assert np.array_equal(got, oracle.opencl_path(data, t)[0]), (t, kind, n)
        if n >= 1024 and (n & (n - 1)) == 0:
            # validity envelope of RadixSortCPU: equal to std::sort for N = 2^k >= 2^10
            assert np.array_equal(got, oracle.radix_sort_cpu(data, t)), (t, kind, n)
        # digit histogram read-back: pass p counts byte p of the (sign-flipped) key
        u = data.view(np.uint32 if data.itemsize == 4 else np.uint64).copy()
        if t[0] == "i":
            u ^= u.dtype.type(1) << u.dtype.type(u.itemsize * 8 - 1)
        assert hist.shape == (data.itemsize, 256)
        for p in range(data.itemsize):
            want = np.bincount(((u >> u.dtype.type(8 * p)) & u.dtype.type(255)).astype(np.int64), minlength=256)
            assert np.array_equal(hist[p].astype(np.int64), want), (t, kind, n, p)


@pytest.mark.parametrize("t", KEY_TYPES)
@pytest.mark.parametrize("kind", KINDS)
def test_stable_permutation_bit_exact(rsb, gpu, oracle, t, kind):
    for n in SIZES:
        data = oracle.dataset(kind, t, n, SEED)
        got, perm, _ = host_sort(rsb, gpu, data, t, with_perm=True)
        assert np.array_equal(got, np.sort(data)), (t, kind, n)
        assert np.array_equal(perm, np.argsort(data, kind="stable").astype(np.uint32)), (t, kind, n)
        if n % 1024 == 0:
            k2, p2, _, _ = oracle.opencl_path(data, t, with_perm=True)
            assert np.array_equal(got, k2) and np.array_equal(perm, p2), (t, kind, n)


def test_uploaded_permutation_payload_is_carried(rsb, gpu, oracle):
    # the reference uploads whatever h_Permut holds (src/RadixSortGPU.cpp:378-386); carry it stably
    n = 50000
    data = oracle.dataset("random", "u32", n) & np.uint32(0x3FF)  # many duplicates
    payload = (np.arange(n, dtype=np.uint32) * np.uint32(2654435761)).astype(np.uint32)
    got, perm, _ = host_sort(rsb, gpu, data, "u32", with_perm=True, perm_in=payload)
    order = np.argsort(data, kind="stable")
    assert np.array_equal(got, data[order]) and np.array_equal(perm, payload[order])


@pytest.mark.parametrize("t", KEY_TYPES)
def test_golden_vectors(rsb, gpu, golden_small, t):
    for kind in KINDS:
        data = golden_small[f"in_{kind}_{t}_1024"]
        got, perm, _ = host_sort(rsb, gpu, data, t, with_perm=True)
        assert np.array_equal(got, golden_small[f"radixcpu_{kind}_{t}_1024"]), kind
        assert np.array_equal(got, golden_small[f"stdsort_{kind}_{t}_1024"]), kind
        assert np.array_equal(data[perm], got)


def test_size_sweep_u32_key_value(rsb, gpu, oracle):
    # BASELINE.json configs[2]: 2^0 .. 2^k sweep of key + stable index (full 2^28 is a property test below)
    full = oracle.dataset("random", "u32", 1 << 22)
    for k in range(0, 23):
        n = 1 << k
        data = full[:n]
        got, perm, _ = host_sort(rsb, gpu, data, "u32", with_perm=True)
        assert np.array_equal(got, np.sort(data)), k
        assert np.array_equal(perm, np.argsort(data, kind="stable").astype(np.uint32)), k


@pytest.mark.parametrize("t", KEY_TYPES)
def test_ranking_modes_agree(rsb, gpu, oracle, t):
    # default = one shared atomic per key (enabled by the context's lane-order self-test), fallback = lane bitmap
    ctx, _ = gpu
    assert ctx.atomic_rank, "the B200 passes the same-word atomic ordering self-test"
    rng = np.random.default_rng(11)
    n = 200003
    few = rng.integers(0, 3, size=n).astype(numpy_dtype(t))          # long same-digit groups in every row
    two = np.where(np.arange(n) % 2 == 0, 5, 70000).astype(numpy_dtype(t))
    for data in (oracle.dataset("uniform", t, n, SEED), oracle.dataset("range", t, n), few, two):
        a, pa, _ = host_sort(rsb, gpu, data, t, with_perm=True)
        b, pb, _ = host_sort(rsb, gpu, data, t, with_perm=True, flags=rsb.RS_NO_ATOMIC_RANK)
        want = np.argsort(data, kind="stable").astype(np.uint32)
        assert np.array_equal(pa, want) and np.array_equal(pb, want)
        assert np.array_equal(a, data[want]) and np.array_equal(b, a)


@pytest.mark.parametrize("t", KEY_TYPES)
def test_small_input_single_cta_path(rsb, gpu, oracle, t):
    # n <= 4096 is sorted by one CTA in one launch; it must agree with the multi-kernel path bit for bit
    ctx, st = gpu
    rng = np.random.default_rng(5)
    for n in (1, 2, 31, 32, 33, 255, 256, 257, 1000, 4095, 4096):
        for data in (oracle.dataset("uniform", t, n, SEED), oracle.dataset("inverted", t, n),
                     rng.integers(0, 4, size=n).astype(numpy_dtype(t)), np.zeros(n, dtype=numpy_dtype(t))):
            a, pa, ha = host_sort(rsb, gpu, data, t, with_perm=True)
            b, pb, hb = host_sort(rsb, gpu, data, t, with_perm=True, flags=rsb.RS_NO_SMALL_PATH)
            want = np.argsort(data, kind="stable").astype(np.uint32)
            assert np.array_equal(pa, want) and np.array_equal(a, data[want]), (t, n)
            assert np.array_equal(pb, want) and np.array_equal(b, a), (t, n)
            assert np.array_equal(ha, hb), (t, n)
    s = rsb.Sorter(ctx, t, 4096)
    data = oracle.dataset("random", t, 3000)
    l0 = s.launch_count
    s.upload(st, data, None, 3000)
    s.sort(st, 3000)
    assert s.launch_count - l0 == 1
    out = np.empty_like(data)
    s.download(st, out, None, 3000)
    assert np.array_equal(out, np.sort(data))
    s.close()


def test_empty_input(rsb, gpu):
    got, perm, _ = host_sort(rsb, gpu, np.empty(0, dtype=np.uint32), "u32", with_perm=True)
    assert got.size == 0 and perm.size == 0


@pytest.mark.parametrize("t", ("u32", "i64"))
def test_plain_load_path_matches_tma_path(rsb, gpu, oracle, t):
    for kind in ("uniform", "range"):
        data = oracle.dataset(kind, t, 300001, SEED)
        a, pa, _ = host_sort(rsb, gpu, data, t, with_perm=True)
        b, pb, _ = host_sort(rsb, gpu, data, t, with_perm=True, flags=rsb.RS_NO_TMA)
        assert np.array_equal(a, b) and np.array_equal(pa, pb)
        assert np.array_equal(a, np.sort(data))


def test_reference_call_sequence_with_padding(rsb, gpu, oracle):
    # CRadixSortTask::ComputeGPU (src/CRadixSortTask.cpp:120-130): pad, then upload `rounded` keys whose
    # tail past N is zero in the host vector; sort `rounded`; compare the first N with the CPU sort of
    # the same rounded buffer (src/CRadixSortTask.cpp:176-221, 225-244).
    ctx, st = gpu
    n = 5000
    rounded = rsb.round_up_keys(n)
    host = np.zeros(rounded, dtype=np.int32)
    host[:n] = oracle.dataset("uniform", "i32", n, SEED)
    s = rsb.Sorter(ctx, "i32", rounded)
    s.pad(st, n, rounded)
    s.upload(st, host, None, rounded)
    s.sort(st, rounded)
    out = np.empty(rounded, dtype=np.int32)
    s.download(st, out, None, rounded)
    rt = s.runtimes()
    s.close()
    assert np.array_equal(out, np.sort(host))
    assert np.array_equal(out, oracle.opencl_path(host, "i32")[0])
    assert rt["time_total"]["n"] == 1 and rt["time_paste"]["avg"] == 0.0
    assert rt["time_total"]["avg"] >= rt["time_reorder"]["avg"] > 0.0


def test_pad_value_is_max_minus_one(rsb, gpu):
    ctx, st = gpu
    s = rsb.Sorter(ctx, "i64", 2048)
    host = np.zeros(1000, dtype=np.int64)
    s.upload(st, host, None, 1000)
    s.pad(st, 1000, 2048)
    s.sort(st, 2048)
    out = np.empty(2048, dtype=np.int64)
    s.download(st, out, None, 2048)
    s.close()
    assert (out[:1000] == 0).all() and (out[1000:] == np.iinfo(np.int64).max - 1).all()


def test_constant_digit_passes_are_skipped_on_device(rsb, gpu, oracle):
    ctx, st = gpu
    rng = np.random.default_rng(7)
    n = 300001
    cases = {
        # name: (key type, data, passes expected to be skipped)
        "zeros": ("u32", np.zeros(n, dtype=np.uint32), 4),
        "u64 holding 32-bit values (Random<uint64_t>)": ("u64", oracle.dataset("random", "u64", n), 4),
        "constant byte 1 (odd number of passes runs -> fix-up copy)":
            ("u32", (rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32) & np.uint32(0xFFFF00FF)) | np.uint32(0x5500), 1),
        "only byte 2 varies (one pass runs)": ("i32", (rng.integers(0, 256, size=n).astype(np.int32) << 16) | np.int32(0x01000203), 3),
        "negative constant high bytes": ("i64", -(rng.integers(1, 2**20, size=n).astype(np.int64)), 5),
    }
    for name, (t, data, want_skipped) in cases.items():
        for flags in (0, rsb.RS_NO_PASS_SKIP):
            s = rsb.Sorter(ctx, t, n, flags | rsb.RS_WITH_PERMUTATION)
            s.upload(st, data, None, n)
            s.sort(st, n)
            out = np.empty(n, dtype=data.dtype)
            perm = np.empty(n, dtype=np.uint32)
            s.download(st, out, perm, n)
            skipped = s.last_skipped_passes(st)
            s.close()
            assert np.array_equal(out, np.sort(data)), name
            assert np.array_equal(perm, np.argsort(data, kind="stable").astype(np.uint32)), name
            assert skipped == (0 if flags else want_skipped), (name, skipped)
    # device-pointer entry with a separate input buffer: the input is never written, the result is where promised
    torch = _torch()
    data = cases["constant byte 1 (odd number of passes runs -> fix-up copy)"][1]
    d_in = _dev(data)
    bufs = [torch.empty_like(d_in), torch.empty_like(d_in)]
    s = rsb.Sorter(ctx, "u32", n, rsb.RS_NO_KEY_BUFFERS)
    torch.cuda.synchronize()
    r = s.sort_device(st, d_in.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), n)
    st.sync()
    assert np.array_equal(_host(bufs[r], np.uint32), np.sort(data)) and np.array_equal(_host(d_in, np.uint32), data)
    s.close()


# ---- device-pointer entry (torch tensors are only the memory plumbing) -----------------------------

def _torch():
    import torch
    return torch


def _dev(a):
    torch = _torch()
    view = a.view(np.int32 if a.itemsize == 4 else np.int64)
    return torch.from_numpy(view.copy()).cuda()


def _host(tensor, dtype):
    return tensor.cpu().numpy().view(dtype)


@pytest.mark.parametrize("t", KEY_TYPES)
def test_sort_device_bit_ranges(rsb, gpu, oracle, t):
    torch = _torch()
    ctx, st = gpu
    n = 200003
    data = oracle.dataset("uniform", t, n, SEED)
    bits = data.itemsize * 8
    s = rsb.Sorter(ctx, t, n, rsb.RS_NO_KEY_BUFFERS)
    d_in = _dev(data)
    bufs = [torch.empty_like(d_in), torch.empty_like(d_in)]
    vin = torch.arange(n, dtype=torch.int32, device="cuda")
    vbufs = [torch.empty_like(vin), torch.empty_like(vin)]
    torch.cuda.synchronize()
    u = data.view(np.uint32 if data.itemsize == 4 else np.uint64)
    ranges = ((0, bits), (0, 8), (4, 13), (bits - 11, bits), (3, bits - 5), (8, 24))
    if bits == 64:   # five and seven whole-byte digits from bit 0: the split-half histogram with a partly idle upper half
        ranges += ((0, 40), (0, 56))
    for b, e in ranges:
        r = s.sort_device(st, d_in.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), n, vin.data_ptr(),
                          vbufs[0].data_ptr(), vbufs[1].data_ptr(), begin_bit=b, end_bit=e)
        st.sync()
        got, perm = _host(bufs[r], data.dtype), _host(vbufs[r], np.uint32)
        sub = (u >> u.dtype.type(b)) & u.dtype.type((1 << (e - b)) - 1)
        if t[0] == "i" and e == bits:
            sub = sub ^ (u.dtype.type(1) << u.dtype.type(e - b - 1))
        order = np.argsort(sub, kind="stable")
        assert np.array_equal(perm, order.astype(np.uint32)), (t, b, e)
        assert np.array_equal(got, data[order]), (t, b, e)
    assert np.array_equal(_host(d_in, data.dtype), data)  # the input buffer is never written
    s.close()


def test_sort_device_unaligned_pointers(rsb, gpu, oracle):
    torch = _torch()
    ctx, st = gpu
    n = 100000
    data = oracle.dataset("random", "u32", n + 3)
    s = rsb.Sorter(ctx, "u32", n, rsb.RS_NO_KEY_BUFFERS)
    d = _dev(data)
    bufs = [torch.empty(n + 3, dtype=torch.int32, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    for off in (1, 2, 3):
        r = s.sort_device(st, d.data_ptr() + 4 * off, bufs[0].data_ptr() + 4 * off, bufs[1].data_ptr() + 4, n)
        st.sync()
        got = _host(bufs[r], np.uint32)[(off if r == 0 else 1):][:n]
        assert np.array_equal(got, np.sort(data[off:off + n])), off
    s.close()


@pytest.mark.parametrize("t", KEY_TYPES)
def test_device_datasets_closed_forms(rsb, gpu, oracle, t):
    torch = _torch()
    ctx, st = gpu
    n = 70001
    d = torch.empty(n, dtype=torch.int32 if numpy_dtype(t).itemsize == 4 else torch.int64, device="cuda")
    torch.cuda.synchronize()
    for kind in ("zeros", "range", "inverted"):
        ctx.fill_dataset(st, t, kind, d.data_ptr(), n)
        st.sync()
        assert np.array_equal(_host(d, numpy_dtype(t)), oracle.dataset(kind, t, n)), kind
    # shard semantics: elements [off, off+m) of a larger dataset
    ctx.fill_dataset(st, t, "inverted", d.data_ptr(), 1000, global_offset=500, total_n=n)
    st.sync()
    assert np.array_equal(_host(d, numpy_dtype(t))[:1000], oracle.dataset("inverted", t, n)[500:1500])
    ctx.fill_dataset(st, t, "uniform", d.data_ptr(), n, seed=7)
    st.sync()
    u = _host(d, numpy_dtype(t))
    info = np.iinfo(numpy_dtype(t))
    assert u[0] == info.max and u[-1] == info.min and len(np.unique(u)) > n * 0.99
    ctx.fill_dataset(st, t, "random", d.data_ptr(), n, seed=7)
    st.sync()
    r = _host(d, numpy_dtype(t))
    if numpy_dtype(t).itemsize == 8:
        assert int(r.max()) < 2**32 and int(r.min()) >= 0  # 32-bit values zero-extended, like Random<T>


def test_check_kernels_detect_errors(rsb, gpu, oracle):
    torch = _torch()
    ctx, st = gpu
    n = 100000
    data = np.sort(oracle.dataset("uniform", "i32", n, SEED))
    d = _dev(data)
    torch.cuda.synchronize()
    ok = ctx.check(st, "i32", d.data_ptr(), n)
    assert ok["unordered"] == 0 and ok["sum"] == int(data.view(np.uint32).astype(np.uint64).sum())
    bad = data.copy()
    bad[[10, 5000]] = bad[[5000, 10]]
    d2 = _dev(bad)
    torch.cuda.synchronize()
    res = ctx.check(st, "i32", d2.data_ptr(), n)
    assert res["unordered"] >= 1 and res["hash_sum"] == ok["hash_sum"] and res["hash_xor"] == ok["hash_xor"]
    bad[7] ^= 1
    d3 = _dev(bad)
    torch.cuda.synchronize()
    res = ctx.check(st, "i32", d3.data_ptr(), n)
    assert res["hash_sum"] != ok["hash_sum"]
    # lower bound
    q = np.array([np.iinfo(np.int32).min, -5, 0, 77, np.iinfo(np.int32).max], dtype=np.int32)
    dq = _dev(q)
    out = torch.empty(len(q), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.lower_bound(st, "i32", d.data_ptr(), n, dq.data_ptr(), len(q), out.data_ptr())
    st.sync()
    assert np.array_equal(out.cpu().numpy(), np.searchsorted(data, q, side="left"))


@pytest.mark.parametrize("t", ("u32", "i64"))
def test_fused_gather_sort_of_segments(rsb, gpu, oracle, t):
    # the multi-GPU exchange kernel with all segments local: ragged, unaligned, empty segments; every rotation
    torch = _torch()
    ctx, st = gpu
    data = oracle.dataset("uniform", t, 400000, SEED)
    d = _dev(data)
    kb = data.itemsize
    bounds = [0, 1, 1, 70001, 70004, 250000, 399999, 400000]  # 7 segments, one empty, odd offsets
    ptrs = [d.data_ptr() + kb * bounds[i] for i in range(len(bounds) - 1)]
    cnts = [bounds[i + 1] - bounds[i] for i in range(len(bounds) - 1)]
    n = sum(cnts)
    g, b0 = torch.empty_like(d), torch.empty_like(d)
    s = rsb.Sorter(ctx, t, n, rsb.RS_NO_KEY_BUFFERS)
    torch.cuda.synchronize()
    for first in range(len(ptrs)):
        r = s.sort_segments_device(st, ptrs, cnts, first, g.data_ptr(), b0.data_ptr(), g.data_ptr())
        st.sync()
        out = _host((b0, g)[r], data.dtype)[:n]
        assert np.array_equal(out, np.sort(data)), (t, first)
    # a permuted segment order sorts to the same result and the source is never written
    order = [3, 0, 6, 5, 1, 4, 2]
    r = s.sort_segments_device(st, [ptrs[i] for i in order], [cnts[i] for i in order], 2, g.data_ptr(), b0.data_ptr(), g.data_ptr())
    st.sync()
    assert np.array_equal(_host((b0, g)[r], data.dtype)[:n], np.sort(data))
    assert np.array_equal(_host(d, data.dtype), data)
    s.close()


@pytest.mark.parametrize("t", KEY_TYPES)
def test_merge_runs(rsb, gpu, oracle, t):
    # pairwise merge of sorted runs (multi-GPU final step), all runs local here: ragged, empty, several jobs per call
    torch = _torch()
    ctx, st = gpu
    dt = numpy_dtype(t)
    kb = dt.itemsize
    rng = np.random.default_rng(3)
    sizes = [(0, 0), (1, 0), (0, 5), (1, 1), (4096, 4096), (4095, 4097), (100003, 7), (250000, 333333), (5, 60001)]
    runs = []
    for na, nb in sizes:
        a = np.sort(oracle.dataset("uniform", t, na + 1, SEED + na)[:na]) if na else np.empty(0, dt)
        b = np.sort(rng.integers(-50, 50, size=nb).astype(dt)) if nb else np.empty(0, dt)  # duplicates, negatives for signed
        runs.append((a, b))
    splits = torch.empty(1 << 16, dtype=torch.int64, device="cuda")
    for lo in range(0, len(runs), 4):  # up to 8 jobs per call; use 4 here
        batch = runs[lo:lo + 4]
        d_a = [_dev(a) if len(a) else torch.empty(1, dtype=torch.int32 if kb == 4 else torch.int64, device="cuda") for a, _ in batch]
        d_b = [_dev(b) if len(b) else torch.empty(1, dtype=torch.int32 if kb == 4 else torch.int64, device="cuda") for _, b in batch]
        d_o = [torch.empty(max(1, len(a) + len(b)), dtype=d_a[0].dtype, device="cuda") for a, b in batch]
        torch.cuda.synchronize()
        ctx.merge_runs(st, t, [(d_a[i].data_ptr(), len(batch[i][0]), d_b[i].data_ptr(), len(batch[i][1]), d_o[i].data_ptr())
                               for i in range(len(batch))], splits.data_ptr(), splits.numel())
        st.sync()
        for i, (a, b) in enumerate(batch):
            got = _host(d_o[i], dt)[:len(a) + len(b)]
            assert np.array_equal(got, np.sort(np.concatenate((a, b)))), (t, len(a), len(b))


@pytest.mark.parametrize("t", ["u32", "i64"])
def test_merge_runs_many_tiles_per_cta(rsb, gpu, t):
    # 8 jobs of very different sizes, ~3000 tiles: every persistent CTA walks several slots of the interleaved sequence
    torch = _torch()
    ctx, st = gpu
    dt = numpy_dtype(t)
    tdt = torch.int32 if dt.itemsize == 4 else torch.int64
    rng = np.random.default_rng(11)
    sizes = [(3_000_001, 2_500_000), (0, 900_000), (17, 4096 * 300), (1_234_567, 1), (4096 * 64, 4096 * 64), (5, 5), (700_000, 700_001), (0, 0)]
    info = np.iinfo(dt)
    runs = [(np.sort(rng.integers(info.min, info.max, size=na, dtype=dt, endpoint=True)),
             np.sort(rng.integers(info.min // 2, info.max // 2, size=nb, dtype=dt))) for na, nb in sizes]
    splits = torch.empty(1 << 14, dtype=torch.int64, device="cuda")
    d_a = [_dev(a) if len(a) else torch.empty(1, dtype=tdt, device="cuda") for a, _ in runs]
    d_b = [_dev(b) if len(b) else torch.empty(1, dtype=tdt, device="cuda") for _, b in runs]
    d_o = [torch.empty(max(1, len(a) + len(b)), dtype=tdt, device="cuda") for a, b in runs]
    torch.cuda.synchronize()
    ctx.merge_runs(st, t, [(d_a[i].data_ptr(), len(runs[i][0]), d_b[i].data_ptr(), len(runs[i][1]), d_o[i].data_ptr())
                           for i in range(len(runs))], splits.data_ptr(), splits.numel())
    st.sync()
    for i, (a, b) in enumerate(runs):
        got = _host(d_o[i], dt)[:len(a) + len(b)]
        assert np.array_equal(got, np.sort(np.concatenate((a, b)))), (t, i)


@pytest.mark.parametrize("t,n_runs", [("u32", 8), ("i32", 4), ("u64", 8), ("u32", 2), ("i64", 1)])
def test_merge_tree_levels(rsb, gpu, t, n_runs):
    # the multi-GPU final step on ONE GPU: P sorted runs of uneven size (one empty) through dist._merge_tree
    import importlib
    import types
    dsort = importlib.import_module("radix-sort_b200.dist")
    torch = _torch()
    ctx, st = gpu
    dt = numpy_dtype(t)
    kb = dt.itemsize
    tdt = torch.int32 if kb == 4 else torch.int64
    rng = np.random.default_rng(5)
    info = np.iinfo(dt)
    sizes = [300_001, 0, 123_457, 1, 70_000, 555_555, 4096, 250_003][:n_runs]
    runs = [np.sort(rng.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)) for n in sizes]
    pool = _dev(np.concatenate(runs)) if sum(sizes) else torch.empty(1, dtype=tdt, device="cuda")
    ptrs, off = [], 0
    for n in sizes:   # runs sit back to back: arbitrary alignment, like the cut positions of a peer shard
        ptrs.append(pool.data_ptr() + off * kb)
        off += n
    total = sum(sizes)
    slack = 32 * n_runs
    bx, by = (torch.empty(total + slack, dtype=tdt, device="cuda") for _ in range(2))
    ops = types.SimpleNamespace(ctx=ctx, stream=st)
    torch.cuda.synchronize()
    out = dsort._merge_tree(ops, t, ptrs, sizes, bx, by, total, kb, {})
    st.sync()
    assert np.array_equal(_host(out, dt), np.sort(np.concatenate(runs)))


# ---- BASELINE.json full sizes: size-independent properties, everything resident on the device -------

def _full_size_case(rsb, gpu, t, kind, n, with_perm, flags=0):
    torch = _torch()
    ctx, st = gpu
    tdt = torch.int32 if numpy_dtype(t).itemsize == 4 else torch.int64
    d_in = torch.empty(n, dtype=tdt, device="cuda")
    bufs = [torch.empty(n, dtype=tdt, device="cuda") for _ in range(2)]
    vin = vb = None
    if with_perm:
        vin = torch.arange(n, dtype=torch.int32, device="cuda")  # wraps past 2^31 harmlessly: n <= 2^28 here
        vb = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    ctx.fill_dataset(st, t, kind, d_in.data_ptr(), n, seed=SEED)
    s = rsb.Sorter(ctx, t, n, rsb.RS_NO_KEY_BUFFERS | flags)
    r = s.sort_device(st, d_in.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), n,
                      vin.data_ptr() if with_perm else None, vb[0].data_ptr() if with_perm else None,
                      vb[1].data_ptr() if with_perm else None)
    st.sync()
    before = ctx.check(st, t, d_in.data_ptr(), n)
    after = ctx.check(st, t, bufs[r].data_ptr(), n)
    # sorted + same multiset  =>  the unique sorted arrangement, i.e. bit-exact with std::sort
    assert after["unordered"] == 0, (t, kind)
    assert {k: after[k] for k in ("sum", "hash_sum", "hash_xor")} == {k: before[k] for k in ("sum", "hash_sum", "hash_xor")}
    if kind in ("range", "inverted"):
        # closed form: both sort to Range; compare fingerprints with a generated Range
        ctx.fill_dataset(st, t, "range", bufs[1 - r].data_ptr(), n)
        want = ctx.check(st, t, bufs[1 - r].data_ptr(), n)
        assert want == after
    if with_perm:
        res = ctx.check_permutation(st, t, d_in.data_ptr(), bufs[r].data_ptr(), vb[r].data_ptr(), n)
        assert res == {"mismatches": 0, "unstable": 0}, (t, kind)
    s.close()


@pytest.mark.parametrize("kind", KINDS)
def test_full_size_2_28_u32_keys(rsb, gpu, kind):
    _full_size_case(rsb, gpu, "u32", kind, 1 << 28, with_perm=False)      # configs[1]


@pytest.mark.parametrize("kind", KINDS)
def test_full_size_2_28_u32_key_value(rsb, gpu, kind):
    _full_size_case(rsb, gpu, "u32", kind, 1 << 28, with_perm=True)       # configs[2]


@pytest.mark.parametrize("t", ("i32", "u64", "i64"))
@pytest.mark.parametrize("kind", ("uniform", "inverted", "random"))
def test_full_size_2_28_other_types(rsb, gpu, t, kind):
    _full_size_case(rsb, gpu, t, kind, 1 << 28, with_perm=False)          # configs[3]


# ---- the other rows of the reference's type table (int16_t / uint16_t, Common/CLTypeInformation.h:12-22)
#      and floating-point keys as NATIVE key types of the C ABI (no conversion passes) ------------------------------------

def _sort_any(rsb, gpu, keys, t, with_perm, flags=0):
    ctx, st = gpu
    n = len(keys)
    s = rsb.Sorter(ctx, t, max(n, 1), flags | (rsb.RS_WITH_PERMUTATION if with_perm else 0))
    try:
        s.upload(st, keys, None, n)
        s.sort(st, n)
        out = np.empty(n, dtype=keys.dtype)
        perm = np.empty(n, dtype=np.uint32) if with_perm else None
        s.download(st, out, perm, n)
        return out, perm, s.download_histograms(st)
    finally:
        s.close()


@pytest.mark.parametrize("t", ("u16", "i16"))
@pytest.mark.parametrize("mode", ("atomic", "bitmap"))
def test_16_bit_keys_native(rsb, gpu, t, mode):
    dt = rsb.key_dtype(t)
    info = np.iinfo(dt)
    rng = np.random.default_rng(41)
    flags = rsb.RS_NO_ATOMIC_RANK if mode == "bitmap" else 0
    for n in (1, 7, 4096, 4097, 100_003, (1 << 21) + 5):
        for lo, hi in ((info.min, info.max), (0, 3)):
            data = rng.integers(lo, hi, size=n, dtype=dt, endpoint=True)
            got, perm, hist = _sort_any(rsb, gpu, data, t, True, flags)
            order = np.argsort(data, kind="stable")
            assert np.array_equal(got, data[order]), (t, n)
            assert np.array_equal(perm.astype(np.int64), order), (t, n)
            assert hist.shape == (2, 256)       # two 8-bit digits: two passes
    ctx, st = gpu
    s = rsb.Sorter(ctx, t, 2048)
    s.pad(st, 1000, 2048)                       # max - 1, src/RadixSortGPU.cpp:274-276
    s.sort(st, 2048)
    out = np.empty(2048, dtype=dt)
    s.download(st, out, None, 2048)
    assert (out[-1048:] == info.max - 1).all()
    s.close()


def _float_order_key(bits, inf_bits):
    """The order rsb200.h documents for float keys, on unsigned bit patterns: by value, -0.0 < +0.0, every NaN last."""
    nbits = bits.itemsize * 8
    u = bits.astype(np.uint64)
    sign, allb = np.uint64(1 << (nbits - 1)), np.uint64((1 << nbits) - 1)
    key = np.where((u & sign) != 0, u ^ allb, u | sign)
    return np.where((u & (allb ^ sign)) > np.uint64(inf_bits), allb, key)


@pytest.mark.parametrize("t,udt,inf_bits", [("f32", np.uint32, 0x7f800000), ("f64", np.uint64, 0x7ff0000000000000),
                                            ("f16", np.uint16, 0x7c00), ("bf16", np.uint16, 0x7f80)])
def test_float_keys_native(rsb, gpu, t, udt, inf_bits):
    dt = rsb.key_dtype(t)
    rng = np.random.default_rng(42)
    nbits = np.dtype(udt).itemsize * 8
    sign = 1 << (nbits - 1)
    specials = np.array([0, sign, inf_bits, inf_bits | sign, inf_bits + 1, (inf_bits + 1) | sign, (1 << nbits) - 1, 1, sign | 1,
                         inf_bits - 1, (inf_bits - 1) | sign], dtype=udt)   # +-0, +-inf, NaNs of both signs, denormals, +-max
    for n in (11, 4000, 250_007, (1 << 20) + 3):
        bits = rng.integers(0, (1 << nbits) - 1, size=n, dtype=udt, endpoint=True)   # every class of value, NaNs included
        bits[: len(specials)] = specials
        if n > 100_000:
            bits[n // 2:] = rng.choice(specials, size=n - n // 2)                    # heavy ties: stability
        data = bits.view(dt)
        got, perm, hist = _sort_any(rsb, gpu, data, t, True)
        order = np.argsort(_float_order_key(bits, inf_bits), kind="stable")
        assert np.array_equal(got.view(udt), bits[order]), (t, n)                    # bit-exact, NaN payloads included
        assert np.array_equal(perm.astype(np.int64), order), (t, n)
        assert hist.shape == (nbits // 8, 256)
        if dt.kind == "f":   # the same VALUES as numpy's sort (NaNs last; numpy does not tell -0.0 from +0.0)
            assert np.array_equal(got, np.sort(data), equal_nan=True), (t, n)
    ctx, st = gpu
    d = rng.integers(0, (1 << nbits) - 1, size=1 << 16, dtype=udt, endpoint=True)
    import torch
    dev = torch.from_numpy(d.view(np.int16 if nbits == 16 else np.int32 if nbits == 32 else np.int64).copy()).cuda()
    torch.cuda.synchronize()
    assert ctx.check(st, t, dev.data_ptr(), d.size)["unordered"] > 0                 # the device check knows the float order
    got, _, _ = _sort_any(rsb, gpu, d.view(dt), t, False)
    dev.copy_(torch.from_numpy(got.view(np.int16 if nbits == 16 else np.int32 if nbits == 32 else np.int64).copy()))
    torch.cuda.synchronize()
    assert ctx.check(st, t, dev.data_ptr(), d.size)["unordered"] == 0


def test_integer_only_entries_reject_float_keys(rsb, gpu):
    ctx, st = gpu
    with pytest.raises(rsb.RsError):
        ctx.lower_bound(st, "f32", 0, 0, 1, 1, 1)
    with pytest.raises(rsb.RsError):
        ctx.merge_runs(st, "f32", [(1, 1, 1, 1, 1)], 1, 16)
    with pytest.raises(rsb.RsError):
        ctx.merge_runs(st, "u16", [(1, 1, 1, 1, 1)], 1, 16)


def test_sort_device_rejects_aliased_buffers(rsb, gpu):
    torch = _torch()
    ctx, st = gpu
    a, b = (torch.zeros(1024, dtype=torch.int32, device="cuda") for _ in range(2))
    s = rsb.Sorter(ctx, "u32", 1024, rsb.RS_NO_KEY_BUFFERS)
    with pytest.raises(rsb.RsError):
        s.sort_device(st, a.data_ptr(), a.data_ptr(), b.data_ptr(), 1024)     # in == buf0
    with pytest.raises(rsb.RsError):
        s.sort_device(st, a.data_ptr(), b.data_ptr(), b.data_ptr(), 1024)     # buf0 == buf1
    s.sort_device(st, a.data_ptr(), b.data_ptr(), a.data_ptr(), 1024)         # in == buf1: the documented double buffer
    st.sync()
    s.close()


def test_dist_abi_single_gpu(rsb, gpu, oracle):
    # rs_dist_* on a one-GPU "box": NCCL is loaded and initialised, the sort degenerates to the local onesweep.
    # (2 .. 8 GPUs: tests/cpp/dist_runner and bench.py --gpus N run the exact compares on the multi-GPU boxes.)
    torch = _torch()
    for t in ("u32", "i64"):
        data = oracle.dataset("uniform", t, 300_007, SEED)
        d = rsb.DistSorter.all_gpus(t, len(data) + 1024, ngpus=1)
        dev = _dev(data)
        torch.cuda.synchronize()
        (ptr, n), = d.sort([dev.data_ptr()], [len(data)])
        d.sync()
        assert n == len(data)
        out = torch.as_tensor(importlib_dist()._DevicePtr(ptr, n, data.itemsize), device="cuda")
        assert np.array_equal(_host(out, data.dtype), np.sort(data))
        assert d.stage_ms(0)["local_sort"] > 0
        d.close()
    with pytest.raises(rsb.RsError):
        rsb.DistSorter.all_gpus("f32", 1024, ngpus=1)     # integer keys of 4 or 8 bytes only
    with pytest.raises(rsb.RsError):
        rsb.DistSorter.all_gpus("u32", 1024, ngpus=64)


def importlib_dist():
    import importlib
    return importlib.import_module("radix-sort_b200.dist")


# ---- BASELINE.json configs[1..3] at FULL size, EXACT: inputs from the reference's own generators (mt19937 /
#      uniform_int_distribution streams, oracle.dataset == src/Dataset.h:84-137), sorted through the reference-facing
#      upload -> sort -> download calls and compared byte for byte with np.sort -- what CRadixSortTask::ValidateResults
#      asserts with memcmp against std::sort (src/CRadixSortTask.cpp:225-252, tests/tests.cpp:18-28) ------------------

def _assert_stable_permutation(data, got, perm):
    """perm is THE stable sorting permutation of data (given got == sort(data)): gathers to got, and is increasing
    inside every run of equal keys -- which also makes it a bijection."""
    assert perm.dtype == np.uint32 and int(perm.max()) < len(data)
    assert np.array_equal(data[perm], got)
    eq = got[1:] == got[:-1]
    assert bool((perm[1:][eq] > perm[:-1][eq]).all())


@pytest.mark.parametrize("kind", KINDS)
def test_exact_2_28_u32_reference_datasets(rsb, gpu, oracle, kind):
    data = oracle.dataset(kind, "u32", 1 << 28, SEED)
    got, _, _ = host_sort(rsb, gpu, data, "u32")
    assert got.tobytes() == np.sort(data).tobytes(), kind


@pytest.mark.parametrize("kind,flags", [("random", 0), ("uniform", "bitmap")])
def test_exact_2_28_u32_key_index(rsb, gpu, oracle, kind, flags):
    # configs[2]: keys + stable permutation at full size, default ranking mode and the lane-bitmap fallback;
    # 2^28 draws of a 32-bit generator hold ~2^23 duplicate pairs, so stability is exercised
    data = oracle.dataset(kind, "u32", 1 << 28, SEED)
    got, perm, _ = host_sort(rsb, gpu, data, "u32", with_perm=True, flags=rsb.RS_NO_ATOMIC_RANK if flags else 0)
    assert got.tobytes() == np.sort(data).tobytes()
    _assert_stable_permutation(data, got, perm)


@pytest.mark.parametrize("t,kind", [("i64", "uniform"), ("i32", "random")])
def test_exact_2_28_signed_and_wide_keys(rsb, gpu, oracle, t, kind):
    # configs[3]: sign-bit flip and 8-byte keys at full size, exact
    data = oracle.dataset(kind, t, 1 << 28, SEED)
    got, _, _ = host_sort(rsb, gpu, data, t)
    assert got.tobytes() == np.sort(data).tobytes()


@pytest.mark.parametrize("flags", [0, "bitmap"])
def test_full_size_2_28_u64_key_value_both_modes(rsb, gpu, flags):
    # 8-byte keys + index at full size in both ranking modes (in-situ evidence for the lane-order property the default
    # mode rests on: any out-of-order service of same-word atomics shows up as a stability violation here)
    _full_size_case(rsb, gpu, "u64", "uniform", 1 << 28, with_perm=True, flags=rsb.RS_NO_ATOMIC_RANK if flags else 0)
    _full_size_case(rsb, gpu, "u64", "random", 1 << 28, with_perm=True, flags=rsb.RS_NO_ATOMIC_RANK if flags else 0)


# ---- transfer stage (src/RadixSortGPU.cpp:365-429): pinned host buffers, asynchronous download,
#      several sorters on several streams -- the call pattern bench.py's `e2e` runs on -----------------------------------

@pytest.mark.parametrize("t,with_perm", [("u32", False), ("i64", True)])
def test_pinned_async_lanes_bit_exact(rsb, gpu, oracle, t, with_perm):
    ctx, _ = gpu
    n, lanes_n, batches = (1 << 22) + 4321, 3, 7
    dt = numpy_dtype(t)
    flags = rsb.RS_WITH_PERMUTATION if with_perm else 0
    lanes = [dict(s=rsb.Sorter(ctx, t, n, flags), st=rsb.Stream(ctx), hin=rsb.HostBuffer(n, dt), hout=rsb.HostBuffer(n, dt),
                  hperm=rsb.HostBuffer(n, np.uint32) if with_perm else None, batch=None) for _ in range(lanes_n)]
    inputs = [oracle.dataset(("uniform", "random", "inverted")[b % 3], t, n, SEED + b) for b in range(batches)]

    def verify(lane):
        b = lane["batch"]
        if b is None:
            return
        lane["st"].sync()                      # the lane's previous batch (its pinned buffers) is complete
        got = lane["hout"].array
        assert np.array_equal(got, np.sort(inputs[b])), (t, b)
        if with_perm:
            _assert_stable_permutation(inputs[b], got, lane["hperm"].array)
        got[:] = 0                             # a stale buffer must not pass for the next batch

    old_piece = os.environ.get("RSB200_COPY_CHUNK_MIB")
    try:
        for b in range(batches):
            # host copies in 1 MiB pieces, whole, and in the default 128 MiB pieces (read by the library per call)
            os.environ["RSB200_COPY_CHUNK_MIB"] = ("1", "0", "128")[b % 3]
            lane = lanes[b % lanes_n]
            verify(lane)
            lane["hin"].array[:] = inputs[b]
            lane["s"].upload(lane["st"], lane["hin"], None, n)
            lane["s"].sort(lane["st"], n)
            lane["s"].download_async(lane["st"], lane["hout"], lane["hperm"], n)   # returns at once: nothing is synchronised here
            lane["batch"] = b
        for lane in lanes:
            verify(lane)
    finally:
        if old_piece is None:
            os.environ.pop("RSB200_COPY_CHUNK_MIB", None)
        else:
            os.environ["RSB200_COPY_CHUNK_MIB"] = old_piece
        for lane in lanes:
            lane["s"].close(); lane["st"].close(); lane["hin"].close(); lane["hout"].close()
            if lane["hperm"] is not None:
                lane["hperm"].close()


def test_lookback_chunk_handover_exact(tmp_path):
    # the >2^30-key path (a chunk's last tile hands the next chunk its base) checked EXACTLY at small sizes by shrinking
    # the chunk to 3 tiles through the library's test hook; separate process because the hook is read once
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RSB200_DEBUG_TILES_PER_CHUNK="3")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "debug_chunks.py")], capture_output=True, text=True,
                         env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count(" OK ") == 24 and "FAIL" not in out.stdout, out.stdout[-2000:]


def test_pipelined_pass_variant_exact():
    # RSB200_SWEEP_V2=1 selects the software-pipelined variant of the pass kernel (csrc/rs_sweep.cuh, an opt-in kept for
    # the record of the round-2 experiments): every key type / payload / size class of tools/sanitize_driver.py, exact
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RSB200_SWEEP_V2="1")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize_driver.py"), "log2n=20"], capture_output=True,
                         text=True, env=env, timeout=900)
    assert out.returncode == 0 and "sanitize_driver PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
    env["RSB200_DEBUG_TILES_PER_CHUNK"] = "3"     # and its chunk hand-over
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "debug_chunks.py")], capture_output=True, text=True,
                         env=env, timeout=600)
    assert out.returncode == 0 and out.stdout.count(" OK ") == 24 and "FAIL" not in out.stdout, out.stdout[-2000:]


def test_merge_walk_order_regions_exact():
    # RSB200_MERGE_ORDER_REGIONS=1: the region-strided walk order of the merge CTAs (csrc/rs_merge_order.h), the second
    # instantiation of the merge tile kernel -- merge tree and merge levels of tools/sanitize_driver.py, exact
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RSB200_MERGE_ORDER_REGIONS="1")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize_driver.py"), "log2n=21"], capture_output=True,
                         text=True, env=env, timeout=900)
    assert out.returncode == 0 and "sanitize_driver PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]


def test_more_than_one_lookback_chunk(rsb, gpu):
    # > 2^30 keys: the look-back chain restarts at a chunk boundary with offsets from the chunk histograms
    _full_size_case(rsb, gpu, "u32", "uniform", (1 << 30) + 77777, with_perm=False)
    _full_size_case(rsb, gpu, "u32", "zeros", (1 << 30) + 8192 * 3 + 1, with_perm=False)


@pytest.mark.parametrize("dt_name", ["int32", "int64", "uint32", "uint64"])
def test_torch_sort_operator(rsb, dt_name):
    # sort(tensor) -> (values, indices), zero copy on torch's current stream (radix-sort_b200/torch_op.py)
    import importlib
    op = importlib.import_module("radix-sort_b200.torch_op")
    torch = _torch()
    dt = getattr(torch, dt_name)
    np_dt = np.dtype(dt_name)
    rng = np.random.default_rng(21)
    info = np.iinfo(np_dt)
    for n, lo, hi in ((1, info.min, info.max), (5000, info.min, info.max), (300_000, 0, 50), (1 << 21, info.min, info.max)):
        host = rng.integers(lo, hi, size=n, dtype=np_dt, endpoint=True)
        keys = torch.from_numpy(host.view(np.int32 if np_dt.itemsize == 4 else np.int64).copy()).cuda().view(dt)
        values, indices = op.sort(keys)
        torch.cuda.synchronize()
        order = np.argsort(host, kind="stable")
        assert values.dtype == dt and indices.dtype == torch.int32
        got_v = values.view(torch.int32 if np_dt.itemsize == 4 else torch.int64).cpu().numpy().view(np_dt)
        assert np.array_equal(got_v, host[order]), (dt_name, n)
        assert np.array_equal(indices.cpu().numpy().astype(np.int64), order), (dt_name, n)
        keys_only, none = op.sort(keys, return_indices=False)
        torch.cuda.synchronize()
        assert none is None
        assert np.array_equal(keys_only.view(torch.int32 if np_dt.itemsize == 4 else torch.int64).cpu().numpy().view(np_dt), host[order])
        assert np.array_equal(keys.view(torch.int32 if np_dt.itemsize == 4 else torch.int64).cpu().numpy().view(np_dt), host)  # input untouched
    with pytest.raises(ValueError):
        op.sort(torch.arange(4, dtype=torch.int32))   # CPU tensor: no fallback


def test_torch_sort_operator_float_keys(rsb):
    # float keys ride on the integer sort through the order-preserving bit transform (torch_op.py)
    import importlib
    op = importlib.import_module("radix-sort_b200.torch_op")
    torch = _torch()
    rng = np.random.default_rng(22)
    for np_dt in (np.float32, np.float64):
        host = (rng.standard_normal(200_000) * 1e4).astype(np_dt)
        host[:4] = [0.0, -0.0, np.inf, -np.inf]
        values, indices = op.sort(torch.from_numpy(host.copy()).cuda())
        torch.cuda.synchronize()
        got = values.cpu().numpy()
        assert np.array_equal(got, np.sort(host)) and np.array_equal(host[indices.cpu().numpy()], got)


def test_torch_sort_operator_half_and_nan(rsb):
    import importlib
    op = importlib.import_module("radix-sort_b200.torch_op")
    torch = _torch()
    g = torch.Generator(device="cuda").manual_seed(5)
    for dt in (torch.float16, torch.bfloat16, torch.float32):
        keys = (torch.randn(300_000, device="cuda", generator=g) * 100).to(dt)
        keys[::1000] = float("nan")
        keys[1::1000] = float("-inf")
        values, indices = op.sort(keys)
        ref_v, ref_i = torch.sort(keys, stable=True)       # torch: NaNs last, stable -- the same placement
        torch.cuda.synchronize()
        assert torch.equal(values.isnan(), ref_v.isnan())
        assert torch.equal(values[~values.isnan()], ref_v[~ref_v.isnan()])
        assert torch.equal(keys[indices.long()].isnan(), values.isnan())
        assert torch.equal(keys[indices.long()][~values.isnan()], values[~values.isnan()])


@pytest.mark.parametrize("dt_name", ["int16", "uint16"])
def test_torch_sort_operator_16_bit_keys(rsb, dt_name):
    import importlib
    op = importlib.import_module("radix-sort_b200.torch_op")
    torch = _torch()
    np_dt = np.dtype(dt_name)
    info = np.iinfo(np_dt)
    host = np.random.default_rng(24).integers(info.min, info.max, size=700_001, dtype=np_dt, endpoint=True)
    keys = torch.from_numpy(host.view(np.int16).copy()).cuda().view(getattr(torch, dt_name))
    values, indices = op.sort(keys)
    torch.cuda.synchronize()
    order = np.argsort(host, kind="stable")
    assert values.dtype == getattr(torch, dt_name)
    assert np.array_equal(values.view(torch.int16).cpu().numpy().view(np_dt), host[order])
    assert np.array_equal(indices.cpu().numpy().astype(np.int64), order)


def test_torch_sort_operator_int16_keys(rsb):
    # 16-bit keys are a native key type: two onesweep passes over 2-byte keys
    import importlib
    op = importlib.import_module("radix-sort_b200.torch_op")
    torch = _torch()
    host = np.random.default_rng(23).integers(-32768, 32767, size=500_000, dtype=np.int16, endpoint=True)
    values, indices = op.sort(torch.from_numpy(host.copy()).cuda())
    torch.cuda.synchronize()
    order = np.argsort(host, kind="stable")
    assert values.dtype == torch.int16
    assert np.array_equal(values.cpu().numpy(), host[order]) and np.array_equal(indices.cpu().numpy().astype(np.int64), order)
