#!/usr/bin/env python
"""bench.py -- the contract bench of the B200 radix sort (one JSON line on stdout, rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dataset uniform]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N

A "step" is one complete sort of the workload through the hot path:
  N = 1 : 2^28 uint32 keys, keys-only (BASELINE.json configs[1]), dataset `--dataset` (default: uniform),
          inputs resident in HBM; every step re-sorts the same unsorted input (the sort never writes it);
          1 GiB in / 1 GiB out per pass, far larger than the 126 MB L2.
  N > 1 : 2^33 uint32 keys in total (configs[4]; strong scaling over N = 2, 4, 8), generated on the
          devices per shard; a step = local onesweep + splitter all-reduce + NCCL all-to-all + final sort.
`value` is whole-job Gkeys/s (keys of all ranks / max-over-ranks device time, CUDA events).
`e2e` runs the same metric through the reference-facing call sequence of the C ABI
(rs_sorter_upload -> rs_sorter_sort -> rs_sorter_download, pinned HOST buffers; copies inside the timing).
`roofline` is the dominant kernel (one onesweep pass): algorithmic bytes 2*sizeof(key)*n per launch over the
kernel's event-timed duration, against MEASURED_PEAKS.json's hbm_gbs.
`cpu_baseline` times the reference's own RadixSortCPU (oracle/_ref, unmodified reference headers; falls back to
the C restatement in oracle/) on the host, single thread (the reference is single-threaded), on a bounded sample.
`--impl reference` prints the same line for that CPU implementation alone.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Gkeys/s sorting 2^28 uint32 keys"
METRIC_DIST = "Gkeys/s sorting 2^33 uint32 keys distributed over the GPUs of one box"
UNIT = "Gkeys/s"
LOG2_SINGLE = 28
LOG2_DIST_TOTAL = 33
KEY_TYPE = "u32"
KEY_BYTES = 4
PASSES = 4
CPU_SAMPLE_LOG2 = 24          # cpu_baseline leg of our own arm: bounded sample, 2^24 keys of the same workload
REFERENCE_BUDGET_S = 100.0    # --impl reference: whole 2^28-key sorts until K steps or this many seconds (>= 1 sort)
NVLINK_PEER_READ_GBS = 672.0  # measured both-ways peer-read rate of one GPU (tools/peerbw.cu, profiles/r1f_peerbw_aligned.txt)
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for name, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples and self.samples[0][1].isdigit() else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_gkeys(sample_log2, dataset, iterations, budget_s=None):
    """RadixSortCPU<uint32_t>::sort on the first 2^sample_log2 keys of the workload's input (all of it for 28), averaged
    like _NUM_PERFORMANCE_ITERATIONS (src/CRadixSortTask.cpp:203-221).  Single host thread: the reference is
    single-threaded.  With budget_s, sorts are timed one at a time until `iterations` or the budget is reached."""
    from oracle import Oracle, Reference, have_reference
    n = 1 << sample_log2
    ora = Oracle()
    data = ora.dataset(dataset, KEY_TYPE, n, 0x5EED)
    ref = Reference() if have_reference() else None
    kind = "reference" if ref is not None else "port"

    def sort_ms(k):
        if ref is not None:
            return ref.time_radix_sort_cpu_ms(data, KEY_TYPE, k) * k
        t0 = time.perf_counter()
        for _ in range(k):
            ora.radix_sort_cpu(data, KEY_TYPE)
        return (time.perf_counter() - t0) * 1e3

    if budget_s is None:
        done, total_ms = iterations, sort_ms(iterations)
    else:
        done, total_ms = 0, 0.0
        while done < iterations and (done == 0 or total_ms * 1e-3 * (done + 1) / done < budget_s):
            total_ms += sort_ms(1)
            done += 1
    ms = total_ms / done
    return {"value": n / ms / 1e6, "unit": UNIT, "cores": 1, "kind": kind, "host_cores_available": os.cpu_count(),
            "sample": f"2^{sample_log2} uint32 keys, '{dataset}' dataset (reference generator, "
                      f"{'its fixed seed_seq' if dataset == 'random' else 'seed 0x5EED'}), "
                      f"{done} iteration{'s' if done != 1 else ''} averaged, RadixSortCPU<uint32_t>::sort, 1 thread "
                      f"(the reference is single-threaded; the host has {os.cpu_count()} cores)",
            "ms_per_sort": ms, "iterations": done}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path, on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    dataset = args.dataset
    # warm-up (a small sort), then whole sorts of the N=1 workload itself -- all 2^28 keys, one step = one sort (about
    # half a minute each) -- until K steps or REFERENCE_BUDGET_S; the line says how many were timed.  For N > 1 the
    # workload is 2^33 keys (32 GiB): out of reach of a bounded CPU step, so the same 2^28-key sort is the sample.
    for _ in range(min(args.warmup, 1)):
        cpu_reference_gkeys(20, dataset, 1)
    res = cpu_reference_gkeys(args.log2n, dataset, max(1, args.steps), budget_s=REFERENCE_BUDGET_S)
    n_total = (1 << args.log2n) if args.gpus == 1 else (1 << args.log2n_total)
    note = ("CPU reference timed on the whole workload: " if args.gpus == 1 else
            "CPU reference timed on a bounded sample (2^28 of the 2^33 keys): ") + res["sample"]
    line = {
        "impl": "reference", "metric": METRIC if args.gpus == 1 else METRIC_DIST, "value": res["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": res["iterations"], "steps_requested": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": res["ms_per_sort"], "higher_is_better": True,
        "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus, dataset), "keys_total": n_total, "keys_timed_per_step": 1 << args.log2n,
                   "note": note},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores_available")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_name(gpus, dataset):
    if gpus == 1:
        return f"2^{LOG2_SINGLE} uint32 keys, keys-only, '{dataset}' dataset, 1xB200 (BASELINE.json configs[1])"
    return (f"2^{LOG2_DIST_TOTAL} uint32 keys total, keys-only, '{dataset}' dataset generated per shard on the devices, "
            f"{gpus}xB200: local onesweep + MSD splitter all-reduce + exchange over NVLink + final local sort "
            f"(BASELINE.json configs[4])")


def measure_other_config(rsb, torch, ctx, st, key_type, with_perm, n, steps, peak):
    """Device-resident sort of n keys of another key type / payload (BASELINE.json configs[2], [3] and the added
    key types): Gkeys/s, algorithmic bytes per key and the fraction of the HBM roofline, same timing rules as the
    headline (warm-ups, CUDA events, inputs far larger than L2)."""
    kb = rsb.key_dtype(key_type).itemsize
    tdt = {2: torch.int16, 4: torch.int32, 8: torch.int64}[kb]
    d_in = torch.empty(n, dtype=tdt, device="cuda")
    bufs = [torch.empty(n, dtype=tdt, device="cuda") for _ in range(2)]
    vin = torch.arange(n, dtype=torch.int32, device="cuda") if with_perm else None
    vb = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)] if with_perm else [None, None]
    ctx.fill_dataset(st, key_type, "uniform", d_in.data_ptr(), n, seed=0x5EED)
    sorter = rsb.Sorter(ctx, key_type, n, rsb.RS_NO_KEY_BUFFERS)

    def once():
        return sorter.sort_device(st, d_in.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), n,
                                  vin.data_ptr() if with_perm else None, vb[0].data_ptr() if with_perm else None,
                                  vb[1].data_ptr() if with_perm else None)

    for _ in range(3):
        r = once()
    torch.cuda.synchronize()
    l0 = sorter.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        once()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = sorter.launch_count - l0
    chk = ctx.check(st, key_type, bufs[r].data_ptr(), n)
    if chk["unordered"] != 0:
        raise RuntimeError(f"bench: {key_type} output is not sorted")
    if with_perm:
        res = ctx.check_permutation(st, key_type, d_in.data_ptr(), bufs[r].data_ptr(), vb[r].data_ptr(), n)
        if res != {"mismatches": 0, "unstable": 0}:
            raise RuntimeError(f"bench: {key_type} permutation is not the stable sorting permutation")
    sorter.close()
    passes = kb
    bpk = kb + passes * 2 * (kb + (4 if with_perm else 0))
    gbs = bpk * n / (ms * 1e-3) / 1e9
    return {"key_type": key_type, "index_payload": bool(with_perm), "keys": n, "dataset": "uniform", "ms_per_sort": round(ms, 4),
            "gkeys_s": round(n / ms / 1e6, 3), "passes": passes, "algorithmic_bytes_per_key": bpk,
            "achieved_gbs": round(gbs, 1), "roofline_frac": round(gbs / peak, 4)}, launches


def bench_single(args):
    import numpy as np
    import torch
    rsb = importlib.import_module("radix-sort_b200")
    n = 1 << args.log2n
    torch.cuda.set_device(0)
    ctx = rsb.Context(0)
    st = rsb.Stream(ctx, torch.cuda.current_stream().cuda_stream)
    d_in = torch.empty(n, dtype=torch.int32, device="cuda")
    bufs = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)]
    sorter = rsb.Sorter(ctx, KEY_TYPE, n, rsb.RS_NO_KEY_BUFFERS)

    def sort_once():
        return sorter.sort_device(st, d_in.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), n)

    datasets = ["uniform", "random", "range", "inverted", "zeros"]
    per_dataset = {}
    skipped = {}   # passes skipped on the device because their digit was constant (rates stay per key sorted)
    headline = None
    clocks = None
    launches = 0
    rt_head = None
    sampler = ClockSampler(0)   # samples clocks while all five datasets are sorted (warm-ups + timed steps)
    sampler.start()
    time.sleep(0.3)
    for kind in [args.dataset] + [d for d in datasets if d != args.dataset]:
        ctx.fill_dataset(st, KEY_TYPE, kind, d_in.data_ptr(), n, seed=0x5EED)
        for _ in range(max(args.warmup, 3)):
            r = sort_once()
        torch.cuda.synchronize()
        sorter.reset_runtimes()
        l0 = sorter.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            sort_once()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        chk = ctx.check(st, KEY_TYPE, bufs[r].data_ptr(), n)
        if chk["unordered"] != 0:
            raise RuntimeError(f"bench: output of dataset {kind} is not sorted")
        per_dataset[kind] = round(n / ms / 1e6, 3)
        skipped[kind] = sorter.last_skipped_passes(st)
        if kind == args.dataset:
            headline = ms
            launches = sorter.launch_count - l0
            rt_head = sorter.runtimes()

    # ---- BASELINE.json configs[2] and [3] (key + index, signed and 64-bit keys) and the added key types, same size ----
    peak, peak_src = measured_hbm_peak()
    other_configs, other_launches = [], 0
    if not args.no_other:
        del bufs[:]
        other_steps = max(3, min(args.steps, 10))
        for key_type, with_perm in (("u32", True), ("i32", False), ("u64", False), ("i64", False), ("u64", True), ("f32", False), ("u16", False)):
            res, nl = measure_other_config(rsb, torch, ctx, st, key_type, with_perm, n, other_steps, peak)
            other_configs.append(res)
            other_launches += nl
        bufs = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)]
    clocks = sampler.finish()

    # ---- end to end through the reference-facing calls with pinned HOST buffers ---------------------------
    h_in = rsb.HostBuffer(n, np.uint32)
    ctx.fill_dataset(st, KEY_TYPE, args.dataset, d_in.data_ptr(), n, seed=0x5EED)
    torch.cuda.synchronize()
    h_in.array[:] = d_in.cpu().numpy().view(np.uint32)   # same workload, now living in pinned host memory
    # Three sorters on three streams: while batch i downloads, batch i+1 sorts and batch i+2 uploads (PCIe is full
    # duplex).  Every step still moves its own 1 GiB in and 1 GiB out through the reference-facing calls.
    NLANES = 3
    e2e_steps = max(NLANES, min(args.steps, 9))
    lanes = [(rsb.Sorter(ctx, KEY_TYPE, n), rsb.Stream(ctx), rsb.HostBuffer(n, np.uint32)) for _ in range(NLANES)]

    def e2e_run(steps):
        for i in range(steps):
            srt, stream, hout = lanes[i % NLANES]
            stream.sync()                       # the lane's previous batch (its host buffer) is complete
            srt.upload(stream, h_in, None, n)
            srt.sort(stream, n)
            srt.download_async(stream, hout, None, n)
        for _, stream, _ in lanes:
            stream.sync()

    e2e_run(NLANES)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    # every lane's host result: sorted AND the same multiset as the host input (a stale or partial download fails)
    want = ctx.check(st, KEY_TYPE, d_in.data_ptr(), n)
    for _, _, hout in lanes:
        bufs[0].copy_(torch.from_numpy(hout.array.view(np.int32)))
        torch.cuda.synchronize()
        got = ctx.check(st, KEY_TYPE, bufs[0].data_ptr(), n)
        if got["unordered"] != 0 or any(got[k] != want[k] for k in ("sum", "hash_sum", "hash_xor")):
            raise RuntimeError("bench: e2e output is not the sorted input")
    # the plain sequential call sequence of one batch (what RadixSortGPU::uploadData/calculate/downloadData do)
    srt, stream, hout = lanes[0]
    t0 = time.perf_counter()
    srt.upload(stream, h_in, None, n)
    srt.sort(stream, n)
    srt.download(stream, hout, None, n)
    seq_ms = (time.perf_counter() - t0) * 1e3
    e2e = {"value": n / e2e_ms / 1e6, "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "h2d_bytes_per_step": n * KEY_BYTES, "d2h_bytes_per_step": n * KEY_BYTES,
           "api": "rs_sorter_upload -> rs_sorter_sort -> rs_sorter_download_async on three sorters/streams "
                  "(batch i+2 uploads while batch i+1 sorts and batch i downloads), pinned host buffers, copies issued in 128 MiB pieces",
           "timing": "host wall clock between device synchronisations",
           "verified": "each lane's downloaded buffer: sorted, and multiset fingerprint equal to the host input's",
           "input": "every step uploads the same pinned 1 GiB host buffer (the workload), and downloads 1 GiB",
           "single_batch_sequential_ms": seq_ms, "single_batch_sequential_gkeys_s": n / seq_ms / 1e6}
    for srt, stream, hout in lanes:
        srt.close(); stream.close(); hout.close()

    # ---- roofline of the dominant kernel (one onesweep pass) ----------------------------------------------
    pass_ms = rt_head["time_reorder"]["avg"] / PASSES
    alg_bytes = 2 * KEY_BYTES * n
    achieved = alg_bytes / (pass_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("onesweep_u32_keys_2^28_dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "rs_onesweep_kernel<u32, keys-only> (one pass)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": pass_ms,
                "whole_sort": {"algorithmic_bytes": (KEY_BYTES + PASSES * 2 * KEY_BYTES) * n,
                               "achieved": (KEY_BYTES + PASSES * 2 * KEY_BYTES) * n / (headline * 1e-3) / 1e9,
                               "frac": (KEY_BYTES + PASSES * 2 * KEY_BYTES) * n / (headline * 1e-3) / 1e9 / peak},
                "phases_ms": {"histogram": rt_head["time_histo"]["avg"], "scan": rt_head["time_scan"]["avg"],
                              "onesweep_passes": rt_head["time_reorder"]["avg"], "total": rt_head["time_total"]["avg"]}}
    cpu = cpu_reference_gkeys(CPU_SAMPLE_LOG2, args.dataset, 5) if not args.no_cpu else None
    line = {
        "metric": METRIC, "value": n / headline / 1e6, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": headline, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(1, args.dataset), "keys": n, "passes": PASSES, "digit_bits": 8,
                   "l2": "inputs (1 GiB per buffer) are larger than the 126 MB L2; no flush needed",
                   "datasets_gkeys_s": per_dataset, "datasets_passes_skipped": skipped,
                   "other_configs": other_configs,
                   "verified": "timed loops: output sorted (device check, every dataset and configuration; key+index "
                               "configurations also the stable permutation); exact compares live in tests/",
                   "skip_note": "a pass whose digit is equal for all keys is the identity and is skipped on the device; "
                                "Gkeys/s is keys sorted per second, not bytes moved"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
    }
    if cpu is not None:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores_available")}
        # BASELINE.json configs[0], the reference's own CPU-runnable case (2^20 "Random" uint32), beside it
        c0 = cpu_reference_gkeys(20, "random", 5)
        line["cpu_baseline"]["configs0"] = {"value": c0["value"], "unit": UNIT, "ms_per_sort": c0["ms_per_sort"], "sample": c0["sample"]}
    print(json.dumps(line), flush=True)
    return 0


def bench_distributed(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rsb = importlib.import_module("radix-sort_b200")
    dsort = importlib.import_module("radix-sort_b200.dist")
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    total = 1 << args.log2n_total
    n_local = total // world
    ctx = rsb.Context(local_rank)
    st = rsb.Stream(ctx, torch.cuda.current_stream().cuda_stream)
    cap = n_local + n_local // 8 + (1 << 20)
    exchange = args.exchange
    native = None
    if exchange == "native":
        # the product path: rs_dist_create_rank / rs_dist_sort (C ABI): NCCL, peer mapping, splitters and merge tree inside
        native = dsort.NativeDistSort(rsb, ctx, st, KEY_TYPE, n_local + 1024, dist)
    ops = dsort.CudaOps(rsb, ctx, st, KEY_TYPE, (n_local + n_local // 8 + (1 << 20)) if native is None else 1024)
    ops.final_step = args.final   # "merge": pairwise merge tree over peer memory, "sort": fused pull + histogram + onesweep passes
    if exchange == "p2p":
        # fused NVLink pull + histogram kernel instead of NCCL all-to-all; every rank must be able to map every peer
        ok = 1
        try:
            ops.enable_p2p(dist, cap)
        except Exception as e:  # noqa: BLE001  (CUDA IPC unavailable: fall back collectively, and say so in the line)
            sys.stderr.write(f"[rank {rank}] peer mapping failed, falling back to NCCL all-to-all: {e}\n")
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if ok:
                ops.close()
                ops = dsort.CudaOps(rsb, ctx, st, KEY_TYPE, n_local + n_local // 8 + (1 << 20))
            exchange = "nccl"
    # ---- exact parity on THIS box before anything is timed: 2^22 keys, 4 key types x 5 datasets against numpy,
    #      through the same exchange / final step (tools/dist_check.py); a mismatch fails the bench -------------------
    from tools.dist_check import run_checks
    parity_log = []
    parity_cases, parity_failures = run_checks(rsb, dsort, dist, torch, ctx, st, 1 << 22, exchange, args.final, log=parity_log.append)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"dist_parity_n{world}.txt"), "w") as f:
            f.write("\n".join(parity_log) + f"\n{parity_cases} cases, {parity_failures} failures\n")
    if parity_failures:
        raise RuntimeError(f"bench: distributed sort failed exact parity in {parity_failures} of {parity_cases} cases")
    shard = ops.empty(n_local)
    scratch = {} if native is not None else {k: ops.empty(cap) for k in (("recv", "c") if exchange == "p2p" else ("a", "b", "recv", "c"))}
    ctx.fill_dataset(st, KEY_TYPE, args.dataset, shard.data_ptr(), n_local, seed=0x5EED,
                     global_offset=rank * n_local, total_n=total)
    before = ctx.check(st, KEY_TYPE, shard.data_ptr(), n_local)

    def step(src=None):
        src = shard if src is None else src
        if native is not None:
            return native.sort(src)
        return dsort.distributed_sort(src, ops, dist, KEY_TYPE, scratch=scratch)

    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = e0.elapsed_time(e1) / args.steps
    # kernels of one sort per GPU: memset-free count of the launches rs_dist_sort / distributed_sort enqueue
    levels_ = max(1, (world - 1).bit_length())
    split_rounds = 3   # 32-bit keys: 11 + 11 + 10 bits
    launches = args.steps * ((2 + PASSES) + 3 * split_rounds + 3 + 3 * levels_)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    clocks = sampler.finish() if sampler else None

    def verify(out, before):
        """locally sorted, boundaries ordered across ranks, multiset preserved, exact balance (collectives inside)"""
        after = ctx.check(st, KEY_TYPE, out.data_ptr(), out.numel())
        edge = torch.stack([out[0], out[-1]]).to(torch.int64) & 0xFFFFFFFF if out.numel() else torch.zeros(2, dtype=torch.int64, device="cuda")
        edges = [torch.empty_like(edge) for _ in range(world)]
        dist.all_gather(edges, edge)
        fp = torch.tensor([before["sum"] & ((1 << 62) - 1), before["hash_xor"] & ((1 << 62) - 1), after["sum"] & ((1 << 62) - 1),
                           after["hash_xor"] & ((1 << 62) - 1), after["unordered"], out.numel()], dtype=torch.int64, device="cuda")
        fps = [torch.empty_like(fp) for _ in range(world)]
        dist.all_gather(fps, fp)
        fps = torch.stack(fps).cpu().numpy().astype(np.uint64)
        ok = int(fps[:, 4].sum()) == 0 and int(fps[:, 5].sum()) == total
        mask = np.uint64((1 << 62) - 1)
        ok = ok and (np.bitwise_xor.reduce(fps[:, 1]) == np.bitwise_xor.reduce(fps[:, 3]))
        ok = ok and ((fps[:, 0].sum() & mask) == (fps[:, 2].sum() & mask))
        e = torch.stack(edges).cpu().numpy()
        ok = ok and all(e[r, 1] <= e[r + 1, 0] for r in range(world - 1))
        ok = ok and all(int(fps[r, 5]) == ((r + 1) * total) // world - (r * total) // world for r in range(world))
        return bool(ok)

    # ---- one instrumented step: device time of every stage (CUDA events; the headline above runs without them) ----
    stage_names = ("local_sort", "splitters", "merge_level_1", "merge_levels_local", "closing", "total")
    stages = native.stage_ms() if native is not None else {}
    st_t = torch.tensor([stages.get(k, 0.0) for k in stage_names], dtype=torch.float64, device="cuda")
    dist.all_reduce(st_t, op=dist.ReduceOp.MAX)
    stage_ms = dict(zip(stage_names, [round(float(x), 3) for x in st_t.tolist()])) if native is not None else None

    # ---- the other four datasets at this size (range / inverted arrive already range-partitioned) ------------------
    per_dataset = {args.dataset: round(total / ms_max / 1e6, 2)}
    for kind in ("uniform", "random", "range", "inverted", "zeros"):
        if kind == args.dataset:
            continue
        ctx.fill_dataset(st, KEY_TYPE, kind, shard.data_ptr(), n_local, seed=0x5EED, global_offset=rank * n_local, total_n=total)
        b4 = ctx.check(st, KEY_TYPE, shard.data_ptr(), n_local)
        o = step()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            o = step()
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if not verify(o, b4):
            raise RuntimeError(f"bench: distributed output of dataset {kind} failed verification")
        per_dataset[kind] = round(total / float(tt.item()) / 1e6, 2)
    ctx.fill_dataset(st, KEY_TYPE, args.dataset, shard.data_ptr(), n_local, seed=0x5EED, global_offset=rank * n_local, total_n=total)

    # ---- end to end: shard in pinned host memory -> H2D -> distributed sort -> D2H of this rank's result ----
    def host_mem_available():
        try:
            for ln in open("/proc/meminfo"):
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) * 1024
        except Exception:
            pass
        return 0

    need = (n_local + cap) * KEY_BYTES * world          # pinned bytes over all ranks of this box
    e2e_ok = torch.tensor([1 if host_mem_available() > 3 * need else 0], dtype=torch.int64, device="cuda")
    dist.all_reduce(e2e_ok, op=dist.ReduceOp.MIN)
    e2e = None
    if int(e2e_ok.item()) == 1:
        h_in = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
        h_out = torch.empty(cap, dtype=torch.int32, pin_memory=True)
        h_in.copy_(shard)
        d_shard = ops.empty(n_local)
        e2e_steps = max(1, min(args.steps, 2))

        def e2e_once():
            d_shard.copy_(h_in, non_blocking=True)
            res = step(d_shard)
            h_out[:res.numel()].copy_(res, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return res.numel()

        e2e_once()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e2e_steps):
            n_out = e2e_once()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = {"value": total / e2e_ms / 1e6, "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "h2d_bytes_per_step": n_local * KEY_BYTES * world, "d2h_bytes_per_step": int(n_out) * KEY_BYTES * world,
               "api": "pinned host shard -> H2D -> radix-sort_b200.dist.distributed_sort -> D2H, every rank"}
    else:
        e2e = {"value": None, "unit": UNIT, "skipped": "not enough free host memory for the pinned shard buffers",
               "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    out = step()
    torch.cuda.synchronize()

    # global verification: locally sorted, boundaries ordered across ranks, multiset preserved, exact balance
    ok = verify(out, before)
    if rank == 0:
        if not ok:
            raise RuntimeError("bench: distributed output failed verification")
        peak, peak_src = measured_hbm_peak()
        levels = max(1, (world - 1).bit_length())
        merge_path = exchange == "native" or (exchange == "p2p" and args.final == "merge" and (world & (world - 1)) == 0)
        # algorithmic HBM bytes per key and GPU: the local sort (36) + the final step -- a merge tree reads and writes
        # every key once per level (8 B/key/level; level 1 reads over NVLink, writes locally), a final sort is 36 more
        # (the fused pull kernel replaces the histogram read by the NVLink read)
        sort_bpk = KEY_BYTES + PASSES * 2 * KEY_BYTES
        hbm_bpk = sort_bpk + (2 * KEY_BYTES * levels if merge_path else sort_bpk)
        nvlink_bytes = KEY_BYTES * n_local * (world - 1) / world
        lvl1 = (stage_ms or {}).get("merge_level_1") or None
        line = {
            "metric": METRIC_DIST, "value": total / ms_max / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(world, args.dataset), "keys_total": total, "keys_per_gpu": n_local,
                       "api": "rs_dist_create_rank / rs_dist_sort (C ABI; NCCL + CUDA IPC inside)" if exchange == "native" else "radix-sort_b200.dist (Python orchestration, A/B)",
                       "exchange": (("pairwise merge tree reading the peers' sorted runs over NVLink (CUDA IPC)" if (args.final == "merge" or exchange == "native")
                                     else "fused peer-memory pull + histogram kernel (CUDA IPC over NVLink) + final onesweep passes")
                                    if exchange in ("p2p", "native") else "NCCL all_to_all_single + final local sort"),
                       "l2": "shards are GiBs, far larger than L2",
                       "datasets_gkeys_s": per_dataset,
                       "exact_parity": f"{parity_cases} reduced-size cases (4 key types x 5 datasets, 2^22 keys) equal to numpy on this box "
                                       f"before the timed loop; full size: sortedness + cross-rank order + multiset fingerprint + exact balance"},
            "clocks": clocks,
            "e2e": e2e, "gpu_launches": launches * world,
            "roofline": {"bound": "hbm", "kernel": "whole step per GPU (local sort + " + (f"{levels}-level merge tree)" if merge_path else "final local sort)"),
                         "algorithmic_bytes_per_key": hbm_bpk,
                         "achieved": hbm_bpk * n_local / (ms_max * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                         "frac": hbm_bpk * n_local / (ms_max * 1e-3) / 1e9 / peak,
                         "traffic": None,
                         "stage_ms_max_over_ranks": stage_ms,
                         "limits": {"local_sort": "onesweep passes: shared-memory (LSU) bound, see the N=1 line",
                                    "merge_level_1": "NVLink peer reads", "merge_levels_local": "HBM / LSU (merge-path tile kernel)",
                                    "splitters": "latency of the refinement rounds (small collectives)"},
                         "nvlink": {"bytes_in_per_gpu": nvlink_bytes, "peak_gbs": NVLINK_PEER_READ_GBS,
                                    "peak_source": "measured both-ways peer-read rate of one GPU (tools/peerbw.cu)",
                                    "exchange_ms": lvl1, "achieved_gbs": (nvlink_bytes / (lvl1 * 1e-3) / 1e9) if lvl1 else None,
                                    "frac": (nvlink_bytes / (lvl1 * 1e-3) / 1e9 / NVLINK_PEER_READ_GBS) if lvl1 else None,
                                    "min_exchange_ms": nvlink_bytes / (NVLINK_PEER_READ_GBS * 1e9) * 1e3}},
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dataset", default="uniform", choices=["uniform", "random", "range", "inverted", "zeros"])
    ap.add_argument("--log2n", type=int, default=LOG2_SINGLE, help="single-GPU keys (default 28 = the contract workload)")
    ap.add_argument("--log2n-total", type=int, default=LOG2_DIST_TOTAL, help="multi-GPU total keys (default 33)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (developer runs)")
    ap.add_argument("--no-other", action="store_true", help="skip the other key types / payload configurations (developer runs)")
    ap.add_argument("--final", default="merge", choices=["merge", "sort"],
                    help="multi-GPU final local step with --exchange p2p: merge tree of the P received sorted runs (default) or a "
                         "full local onesweep sort fed by the fused pull + histogram kernel")
    ap.add_argument("--exchange", default="native", choices=["native", "p2p", "nccl"],
                    help="multi-GPU path: native = rs_dist_* behind the C ABI (default; merge tree over peer memory), p2p = the same "
                         "exchange orchestrated from Python (A/B), nccl = all_to_all_single + final sort (A/B)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        if world == 1:
            print(json.dumps({"error": "--gpus N>1 must be launched with torch.distributed.run"}))
            return 2
        return bench_distributed(args)
    return bench_single(args)


if __name__ == "__main__":
    sys.exit(main())
