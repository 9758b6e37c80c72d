"""Multi-GPU parity check: reduced-size EXACT compare of the distributed sort against numpy on rank 0, for every
dataset and key type (4 x 5), plus exact balance of the output shards.  Used two ways:
  * `run_checks(...)` is called by bench.py --gpus N BEFORE its timed loop and raises on any mismatch, so the rc of every
    multi-GPU bench line covers exact parity on that box;
  * stand-alone under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py [log2n_total=22]
      [exchange=native|p2p|nccl] [final=merge|sort]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def run_checks(rsb, dsort, dist, torch, ctx, st, total, mode="native", final="merge", log=None):
    """Returns (cases, failures); every rank takes part (collectives inside)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    failures, cases = 0, 0
    for key_type in ("u32", "i32", "u64", "i64"):
        np_dt = rsb.key_dtype(key_type)
        n_local = total // world + (17 * rank if key_type == "i32" else 0)  # ragged shards for one type
        native = None
        if mode == "native":   # the product path: rs_dist_* behind the C ABI
            native = dsort.NativeDistSort(rsb, ctx, st, key_type, 2 * n_local + 4096, dist)
        ops = dsort.CudaOps(rsb, ctx, st, key_type, 2 * n_local + 4096)
        ops.final_step = final
        if mode == "p2p":
            ops.enable_p2p(dist, 2 * n_local + 4096)
        for kind in ("uniform", "random", "range", "inverted", "zeros"):
            shard = ops.empty(n_local)
            ctx.fill_dataset(st, key_type, kind, shard.data_ptr(), n_local, seed=0x5EED, global_offset=rank * (total // world),
                             total_n=total)
            out = native.sort(shard) if native is not None else dsort.distributed_sort(shard, ops, dist, key_type)
            torch.cuda.synchronize()
            sizes = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([n_local, out.numel()], dtype=torch.int64, device="cuda"))
            sizes = torch.stack(sizes).cpu().numpy()
            m_in, m_out = int(sizes[:, 0].max()), int(sizes[:, 1].max())
            pad_in = torch.zeros(m_in, dtype=shard.dtype, device="cuda"); pad_in[:n_local] = shard
            pad_out = torch.zeros(m_out, dtype=shard.dtype, device="cuda"); pad_out[:out.numel()] = out
            g_in = [torch.empty_like(pad_in) for _ in range(world)]
            g_out = [torch.empty_like(pad_out) for _ in range(world)]
            dist.all_gather(g_in, pad_in)
            dist.all_gather(g_out, pad_out)
            cases += 1
            if rank == 0:
                ins = np.concatenate([g_in[r].cpu().numpy().view(np_dt)[: sizes[r, 0]] for r in range(world)])
                outs = np.concatenate([g_out[r].cpu().numpy().view(np_dt)[: sizes[r, 1]] for r in range(world)])
                n_all = len(ins)
                ok = np.array_equal(outs, np.sort(ins))
                bal = all(int(sizes[r, 1]) == ((r + 1) * n_all) // world - (r * n_all) // world for r in range(world))
                if log is not None:
                    log(f"[{mode}/{final}] {key_type} {kind:9s} total={n_all} world={world}: {'OK' if ok and bal else 'FAIL'} "
                        f"sizes={sizes[:, 1].tolist()}")
                failures += int(not (ok and bal))
        ops.close()
        if native is not None:
            native.close()
    f = torch.tensor([failures], dtype=torch.int64, device="cuda")
    dist.broadcast(f, src=0)
    return cases, int(f.item())


def main():
    import torch
    import torch.distributed as dist
    rsb = importlib.import_module("radix-sort_b200")
    dsort = importlib.import_module("radix-sort_b200.dist")
    args = dict(a.split("=") for a in sys.argv[1:])
    total = 1 << int(args.get("log2n_total", 22))
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = rsb.Context(local_rank)
    st = rsb.Stream(ctx, torch.cuda.current_stream().cuda_stream)
    _, failures = run_checks(rsb, dsort, dist, torch, ctx, st, total, args.get("exchange", "native"), args.get("final", "merge"),
                             log=lambda m: print(m, flush=True))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("dist_check", "PASSED" if failures == 0 else f"FAILED ({failures})")
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
