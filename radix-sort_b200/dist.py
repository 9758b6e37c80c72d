"""Multi-GPU sort on one NVSwitch box: one process per GPU (torch.distributed), keys-only.

The product path is `NativeDistSort` at the end of this file: a thin binding of rs_dist_* (include/rsb200.h) -- local
sorts, splitter refinement on the devices, NCCL collectives, peer mapping and the merge tree all live behind the C ABI
(csrc/rs_dist_api.inl); torch.distributed only hands the ncclUniqueId to the other ranks.  The rest of this file is the
first, Python-orchestrated version of the same path.  It is kept because its host logic (`plan_exchange`) is what the
gloo tests exercise on CPU and what the C++ planner is checked against (tests/test_dist_plan_cpu.py), and for A/B runs
(`bench.py --exchange p2p|nccl`).

Path (BASELINE.json north_star): every GPU sorts its shard with the local onesweep,
MSD splitter histograms (11 bits per refinement round) are all-reduced, every rank computes the same exact
splitters (ties on a splitter value are split by count -- equal keys are indistinguishable), key
ranges are exchanged with ONE all-to-all-v over NVLink (NCCL), and every GPU sorts what it received.
Afterwards rank r holds exactly its share of the global order: keys(r) <= keys(r+1), and the
concatenation over ranks is the sorted input.

The reference has no counterpart (one OpenCL device, Common/ComputeState.cpp:10-12); this is the new
L2' layer of SURVEY.md section 1.  torch is only plumbing here (device memory, NCCL collectives);
all sorting and searching on the GPU goes through librsb200.so (`CudaOps`).  The host logic is
written against the small `ops` interface so it can be exercised on CPU with gloo by the tests.
"""
from __future__ import annotations

import numpy as np

LEVEL_BITS = 11   # splitter refinement fan-out per round: 2048 sub-ranges (32-bit keys: 11 + 11 + 10 bits)


def _key_bits(key_type: str) -> int:
    return 32 if key_type in ("u32", "i32") else 64


def _sign(key_type: str) -> int:
    return (1 << (_key_bits(key_type) - 1)) if key_type[0] == "i" else 0


class CudaOps:
    """Device operations of the distributed sort, all through the C ABI (no CPU fallback)."""

    def __init__(self, rsb, ctx, stream, key_type: str, max_keys: int):
        import torch
        # the collectives that order this rank's kernels against its peers run on torch's current stream: the kernels
        # must run there too, or a peer could overwrite a shard that is still being read
        if getattr(stream, "wrapped", None) != torch.cuda.current_stream().cuda_stream:
            raise ValueError("CudaOps: `stream` must wrap torch's current CUDA stream")
        self.torch, self.rsb, self.ctx, self.stream, self.key_type = torch, rsb, ctx, stream, key_type
        self.tdt = torch.int32 if _key_bits(key_type) == 32 else torch.int64
        self.sorter = rsb.Sorter(ctx, key_type, max_keys, rsb.RS_NO_KEY_BUFFERS | rsb.RS_NO_PROFILE)
        self.max_keys = max_keys
        self._q = None
        self._pos = None

    def empty(self, n: int):
        return self.torch.empty(n, dtype=self.tdt, device="cuda")

    def local_sort(self, keys, buf0, buf1):
        """Sort `keys` (device tensor) using two scratch tensors; returns the tensor holding the result."""
        n = keys.numel()
        if n == 0:
            return buf0[:0]
        r = self.sorter.sort_device(self.stream, keys.data_ptr(), buf0.data_ptr(), buf1.data_ptr(), n)
        return (buf0, buf1)[r][:n]

    def lower_bound(self, sorted_keys, queries: np.ndarray) -> np.ndarray:
        """positions[i] = first index with key >= queries[i] (type-aware); queries are raw key bit patterns."""
        torch = self.torch
        nq = len(queries)
        signed = queries.astype(np.uint64).view(np.int64) if self.tdt == torch.int64 else \
            queries.astype(np.uint32).view(np.int32)
        q = torch.from_numpy(np.ascontiguousarray(signed)).to("cuda", non_blocking=False)
        pos = torch.empty(nq, dtype=torch.int64, device="cuda")
        torch.cuda.current_stream().synchronize()
        self.ctx.lower_bound(self.stream, self.key_type, sorted_keys.data_ptr(), sorted_keys.numel(), q.data_ptr(), nq,
                             pos.data_ptr())
        self.stream.sync()
        return pos.cpu().numpy()

    # ---- peer-memory exchange (NVLink, CUDA IPC) ---------------------------------------------------------
    def enable_p2p(self, dist, capacity_keys: int, group=None):
        """Allocate the two local-sort buffers with plain cudaMalloc, export them through CUDA IPC and map every
        peer's pair.  After this, distributed_sort replaces the NCCL all-to-all + histogram pass of the final
        sort by ONE kernel that pulls the peers' segments over NVLink while histogramming them."""
        torch = self.torch
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        kb = 4 if self.tdt == torch.int32 else 8
        self.p2p_capacity = capacity_keys
        self.p2p_local = [self.ctx.device_alloc(capacity_keys * kb) for _ in range(2)]
        mine = torch.tensor(list(b"".join(self.ctx.ipc_export(p) for p in self.p2p_local)), dtype=torch.uint8, device="cuda")
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        self.p2p_peer = []
        for r in range(world):
            if r == rank:
                self.p2p_peer.append(list(self.p2p_local))
            else:
                raw = bytes(parts[r].cpu().tolist())
                self.p2p_peer.append([self.ctx.ipc_open(raw[0:64]), self.ctx.ipc_open(raw[64:128])])
        self.p2p_rank, self.p2p_world = rank, world
        return self

    def p2p_enabled(self) -> bool:
        return getattr(self, "p2p_peer", None) is not None

    def close(self):
        if self.p2p_enabled():
            for r, pair in enumerate(self.p2p_peer):
                if r != self.p2p_rank:
                    for p in pair:
                        self.ctx.ipc_close(p)
            for p in self.p2p_local:
                self.ctx.device_free(p)
            self.p2p_peer = None
        self.sorter.close()


class RawBuf:
    """A device range known by address (cudaMalloc'd exchange buffers are not torch tensors)."""

    def __init__(self, ptr: int, n: int, device=None):
        self._ptr, self._n, self.device = ptr, n, device

    def data_ptr(self) -> int:
        return self._ptr

    def numel(self) -> int:
        return self._n


def plan_exchange(sorted_keys, ops, dist, key_type: str, group=None, device=None, need_recv_counts: bool = True):
    """Exact splitters + per-destination cut positions in this rank's sorted shard.

    Returns (cuts, recv_counts): cuts has world+1 entries (cuts[0]=0, cuts[-1]=n_local); rank d receives
    sorted_keys[cuts[d]:cuts[d+1]] from this rank.  Rank j ends with the keys whose global (stable) rank
    lies in [j*N/P, (j+1)*N/P), so shard sizes differ by at most one key whatever the distribution.
    recv_counts is None when need_recv_counts is False (the peer-memory path learns them from the cut table it
    gathers anyway, which saves one collective and one host sync).
    """
    import torch
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bits, sign = _key_bits(key_type), _sign(key_type)
    top = 1 << bits                      # one past the largest ordered-unsigned key
    n_local = int(sorted_keys.numel())
    dev = device if device is not None else sorted_keys.device

    def allreduce_np(a: np.ndarray) -> np.ndarray:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()

    def positions(ordered: np.ndarray, is_top: np.ndarray | None = None) -> np.ndarray:
        """lower bound of ordered-unsigned values (uint64; entries flagged is_top mean `top`) in the local shard"""
        vals = (ordered & np.uint64(top - 1)) ^ np.uint64(sign)
        pos = ops.lower_bound(sorted_keys, vals).astype(np.int64)
        if is_top is not None:
            pos[is_top] = n_local
        return pos

    if world == 1:
        return np.array([0, n_local], dtype=np.int64), np.array([n_local], dtype=np.int64)
    targets = None               # global rank of the first key of rank j; known after the first all-reduce
    prefix = [0] * (world - 1)   # per splitter: ordered-unsigned key bits resolved so far
    below = [0] * (world - 1)    # per splitter: number of keys globally below `prefix`
    resolved = 0                 # key bits (from the top) already fixed in every splitter prefix
    level = 0
    while resolved < bits:
        width = min(LEVEL_BITS, bits - resolved)
        fan = 1 << width
        shift = bits - resolved - width
        bases = [0] if level == 0 else prefix      # level 0: one histogram serves every splitter
        steps = np.arange(fan + 1, dtype=np.uint64) << np.uint64(shift)   # uint64 wrap-around == mod 2^64
        q = np.stack([np.uint64(b) + steps for b in bases])
        is_top = np.zeros(q.shape, dtype=bool)
        is_top[:, -1] = [b + (fan << shift) >= top for b in bases]         # the last boundary may be `top` itself
        pos = positions(q.reshape(-1), is_top.reshape(-1)).reshape(q.shape)
        counts = allreduce_np(np.diff(pos, axis=1))   # [len(bases)][fan] global sub-range counts
        if level == 0:   # the level-0 histogram covers the whole key space, so it also tells N (no extra all-reduce)
            n_total = int(counts[0].sum())
            targets = [(j * n_total) // world for j in range(1, world)]
        for j in range(world - 1):
            c = counts[0 if level == 0 else j]
            cum = below[j] + np.concatenate(([0], np.cumsum(c)))
            b = int(np.searchsorted(cum, targets[j], side="right") - 1)  # cum[b] <= target < cum[b+1]
            b = min(max(b, 0), fan - 1)
            prefix[j] += b << shift
            below[j] = int(cum[b])
        resolved += width
        level += 1
    # prefix[j] is now an exact key value v_j; the copies of v_j are split by count, in rank order
    pairs = [x for j in range(world - 1) for x in (prefix[j], prefix[j] + 1)]
    pos = positions(np.array([x & 0xFFFFFFFFFFFFFFFF for x in pairs], dtype=np.uint64),
                    np.array([x >= top for x in pairs], dtype=bool))
    lb, ub = pos[0::2], pos[1::2]
    eq_local = ub - lb
    mine = torch.from_numpy(np.ascontiguousarray(eq_local, dtype=np.int64)).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    eq_all = torch.stack(parts).cpu().numpy()  # [rank][splitter], one device->host copy
    cuts = np.zeros(world + 1, dtype=np.int64)
    cuts[-1] = n_local
    for j in range(world - 1):
        need_left = targets[j] - below[j]              # copies of v_j that belong to ranks <= j
        before_me = int(eq_all[:rank, j].sum())
        share = min(max(need_left - before_me, 0), int(eq_local[j]))
        cuts[j + 1] = lb[j] + share
    cuts = np.maximum.accumulate(cuts)  # equal splitters (heavy duplicates) keep the cuts monotone
    if not need_recv_counts:
        return cuts, None
    t_send = torch.from_numpy(np.diff(cuts).copy()).to(dev)
    t_recv = torch.empty_like(t_send)
    dist.all_to_all_single(t_recv, t_send, group=group)
    return cuts, t_recv.cpu().numpy()


def distributed_sort(keys, ops, dist, key_type: str, scratch=None, group=None):
    """Sort the union of all ranks' `keys`; returns this rank's slice of the global order (device tensor).

    keys     : this rank's shard (device tensor, int32/int64 storage of the key bit patterns)
    ops      : CudaOps (product) -- or a test double with local_sort / lower_bound / empty
    scratch  : optional dict of preallocated tensors {"a","b","recv","c"} (each >= the largest shard) so the
               timed region does not allocate
    """
    n_local = int(keys.numel())
    s = scratch or {}
    if getattr(ops, "p2p_enabled", None) and ops.p2p_enabled():
        return _distributed_sort_p2p(keys, ops, dist, key_type, s, group)

    def buf(name, n):
        t = s.get(name)
        return t if (t is not None and t.numel() >= n) else ops.empty(n)

    a, b = buf("a", n_local), buf("b", n_local)
    sorted_local = ops.local_sort(keys, a[:n_local], b[:n_local])
    cuts, recv_counts = plan_exchange(sorted_local, ops, dist, key_type, group=group)
    n_recv = int(recv_counts.sum())
    recv = buf("recv", n_recv)[:n_recv]
    dist.all_to_all_single(recv, sorted_local, output_split_sizes=[int(c) for c in recv_counts],
                           input_split_sizes=[int(c) for c in np.diff(cuts)], group=group)
    # final local sort of the received runs; the buffer the local result does not live in is free again
    free = a if sorted_local.data_ptr() != a.data_ptr() else b
    if free.numel() < n_recv:
        free = ops.empty(n_recv)
    c = buf("c", n_recv)
    return ops.local_sort(recv, free[:n_recv], c[:n_recv])


def _distributed_sort_p2p(keys, ops, dist, key_type: str, s, group):
    """Same result as the NCCL path; the exchange is the fused gather + histogram kernel over peer memory."""
    import os
    import time
    import torch
    timing = os.environ.get("RSB200_DIST_TIMING") == "1"   # developer switch: per-stage wall times (adds syncs)
    marks = []

    def mark(name):
        if timing:
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    mark("start")
    world, rank = ops.p2p_world, ops.p2p_rank
    n_local = int(keys.numel())
    if n_local > ops.p2p_capacity:
        raise ValueError("shard larger than the peer-exchange buffers")
    kb = 4 if _key_bits(key_type) == 32 else 8
    # 1. local sort into the IPC-exported pair; `sel` says which of the two holds the sorted shard
    sel = ops.sorter.sort_device(ops.stream, keys.data_ptr(), ops.p2p_local[0], ops.p2p_local[1], n_local) if n_local else 0
    sorted_local = RawBuf(ops.p2p_local[sel], n_local, device=keys.device)
    mark("local sort")
    # 2. exact splitters (all-reduces inside: every peer's local sort is complete once they return)
    cuts, _ = plan_exchange(sorted_local, ops, dist, key_type, group=group, device=keys.device, need_recv_counts=False)
    mine = torch.from_numpy(np.concatenate((cuts, [sel])).astype(np.int64)).to(keys.device)
    mark("splitters")
    # 3. everybody learns everybody's cut positions and buffer selector
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    table = torch.stack(parts).cpu().numpy()  # [source][world+1 cuts, sel], one device->host copy
    ptrs, counts = [], []
    for src in range(world):
        lo, hi = int(table[src, rank]), int(table[src, rank + 1])
        ptrs.append(ops.p2p_peer[src][int(table[src, world + 1])] + lo * kb)
        counts.append(hi - lo)
    n_recv = int(sum(counts))

    def buf(name, n):
        t = s.get(name)
        return t if (t is not None and t.numel() >= n) else ops.empty(max(n, 1))

    slack = 32 * world   # the merge tree starts every intermediate run on a 128-byte line
    gathered, other = buf("recv", n_recv + slack), buf("c", n_recv + slack)
    if n_recv == 0:
        out = gathered[:0]
    elif getattr(ops, "final_step", "merge") == "merge" and (world & (world - 1)) == 0:
        # 4a. final step as a merge tree: my P segments are sorted runs; level 1 merges pairs of them
        #     straight out of the peers' memory (exchange + merge in one kernel), later levels are local
        out = _merge_tree(ops, key_type, ptrs, counts, gathered, other, n_recv, kb, s)
    else:
        # 4b. fused exchange + histogram, then the onesweep passes of a final local sort
        r = ops.sorter.sort_segments_device(ops.stream, ptrs, counts, rank, gathered.data_ptr(), other.data_ptr(),
                                            gathered.data_ptr())
        out = (other, gathered)[r][:n_recv]
    mark("exchange + final sort")
    # 5. nobody may overwrite its sorted shard (next local sort) before every rank has pulled from it
    done = torch.zeros(1, dtype=torch.int32, device=keys.device)
    dist.all_reduce(done, group=group)
    mark("barrier")
    if timing:
        print(f"[rank {rank}] dist timing (ms): " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f}" for a, b in zip(marks, marks[1:])), flush=True)
    return out


def _merge_tree(ops, key_type, ptrs, counts, buf_x, buf_y, n_total, kb, scratch):
    """Merge the sorted runs (ptr, count) pairwise, level by level, ping-ponging between buf_x and buf_y."""
    import torch
    tiles_max = (n_total + 1023) // 1024 + 2 * len(ptrs) + 16   # enough for any tile shape of the A/B builds (>= 1024 outputs)
    splits = scratch.get("splits")
    if splits is None or splits.numel() < tiles_max:
        splits = torch.empty(tiles_max, dtype=torch.int64, device=buf_x.device)
        scratch["splits"] = splits
    runs = [(int(p), int(c)) for p, c in zip(ptrs, counts)]
    bufs = [buf_x, buf_y]
    level = 0
    if len(runs) == 1:  # single rank: the run is already the result; bring it into local memory through a 1-job merge
        runs.append((runs[0][0], 0))
    while len(runs) > 1:
        dst = bufs[level & 1]
        jobs, nxt, off = [], [], 0
        for i in range(0, len(runs) - 1, 2):
            (pa, na), (pb, nb) = runs[i], runs[i + 1]
            off = (off + 31) & ~31   # intermediate runs start on a 128-byte line: 16-byte stores now, aligned loads next level
            out_ptr = dst.data_ptr() + off * kb
            jobs.append((pa, na, pb, nb, out_ptr))
            nxt.append((out_ptr, na + nb))
            off += na + nb
        if len(runs) & 1:
            nxt.append(runs[-1])  # odd run out: joins the next level (it may still live in peer memory)
        for lo in range(0, len(jobs), 8):
            ops.ctx.merge_runs(ops.stream, key_type, jobs[lo:lo + 8], splits.data_ptr(), splits.numel())
        runs = nxt
        level += 1
    final_ptr, final_n = runs[0]
    dst = bufs[(level - 1) & 1]
    assert final_ptr == dst.data_ptr() and final_n == n_total
    return dst[:n_total]


class _DevicePtr:
    """A raw device range as a CUDA-array-interface object (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr: int, n: int, itemsize: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4" if itemsize == 4 else "<i8", "data": (ptr, False),
                                         "version": 2}


class NativeDistSort:
    """One rank of the multi-GPU sort behind the C ABI (rs_dist_create_rank / rs_dist_sort).  `stream` must wrap the
    CUDA stream the caller orders its own work on (torch's current stream in bench.py): the sort is enqueued there."""

    def __init__(self, rsb, ctx, stream, key_type: str, max_keys_per_gpu: int, dist, group=None):
        import torch
        cur = torch.cuda.current_stream().cuda_stream
        if getattr(stream, "wrapped", None) != cur:
            raise ValueError("NativeDistSort: `stream` must wrap torch's current CUDA stream (rsb.Stream(ctx, "
                             "torch.cuda.current_stream().cuda_stream)): results are consumed by torch on that stream")
        self.torch, self.rsb, self.key_type = torch, rsb, key_type
        self.kb = 4 if _key_bits(key_type) == 32 else 8
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            ident.copy_(torch.frombuffer(bytearray(rsb.dist_unique_id()), dtype=torch.uint8))
        dist.broadcast(ident, src=0, group=group)
        self.sorter = rsb.DistSorter.rank(ctx, stream, key_type, max_keys_per_gpu, self.rank, self.world, bytes(ident.cpu().tolist()))

    def sort(self, shard):
        """Sort the union of all ranks' shards; returns this rank's slice of the global order as a tensor that views the
        sorter's output buffer (valid until the next sort)."""
        (ptr, n), = self.sorter.sort([shard.data_ptr()], [shard.numel()])
        if n == 0:
            return shard[:0]
        return self.torch.as_tensor(_DevicePtr(ptr, n, self.kb), device=shard.device)

    def stage_ms(self):
        return self.sorter.stage_ms(0)

    def close(self):
        self.sorter.close()
