"""Multi-GPU parity check (run under torchrun on the GPU box): reduced-size exact compare of the distributed
sort against numpy on rank 0, for every dataset and key type, plus exact balance.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py [log2n_total=22]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rsb = importlib.import_module("radix-sort_b200")
dsort = importlib.import_module("radix-sort_b200.dist")


def main():
    args = dict(a.split("=") for a in sys.argv[1:])
    total = 1 << int(args.get("log2n_total", 22))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = rsb.Context(local_rank)
    st = rsb.Stream(ctx, torch.cuda.current_stream().cuda_stream)
    failures = 0
    mode = args.get("exchange", "p2p")  # p2p = fused NVLink pull + histogram kernel, nccl = all_to_all_single
    for key_type in ("u32", "i32", "u64", "i64"):
        np_dt = rsb.key_dtype(key_type)
        n_local = total // world + (17 * rank if key_type == "i32" else 0)  # ragged shards for one type
        ops = dsort.CudaOps(rsb, ctx, st, key_type, 2 * n_local + 4096)
        ops.final_step = args.get("final", "merge")
        if mode == "p2p":
            ops.enable_p2p(dist, 2 * n_local + 4096)
        for kind in ("uniform", "random", "range", "inverted", "zeros"):
            shard = ops.empty(n_local)
            ctx.fill_dataset(st, key_type, kind, shard.data_ptr(), n_local, seed=0x5EED, global_offset=rank * (total // world),
                             total_n=total)
            out = dsort.distributed_sort(shard, ops, dist, key_type)
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
            if rank == 0:
                ins = np.concatenate([g_in[r].cpu().numpy().view(np_dt)[: sizes[r, 0]] for r in range(world)])
                outs = np.concatenate([g_out[r].cpu().numpy().view(np_dt)[: sizes[r, 1]] for r in range(world)])
                n_all = len(ins)
                ok = np.array_equal(outs, np.sort(ins))
                bal = all(int(sizes[r, 1]) == ((r + 1) * n_all) // world - (r * n_all) // world for r in range(world))
                print(f"[{mode}/{ops.final_step}] {key_type} {kind:9s} total={n_all} world={world}: {'OK' if ok and bal else 'FAIL'} sizes={sizes[:, 1].tolist()}",
                      flush=True)
                failures += int(not (ok and bal))
        ops.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("dist_check", "PASSED" if failures == 0 else f"FAILED ({failures})")
    sys.exit(1 if failures else 0)


main()
