"""Developer probe: per-rank phase times of the final (fused exchange) sort.  torchrun, 2+ GPUs."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rsb = importlib.import_module("radix-sort_b200"); dsort = importlib.import_module("radix-sort_b200.dist")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr); dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
total = 1 << int(dict(a.split("=") for a in sys.argv[1:]).get("log2n_total", 31)); n = total // world
ctx = rsb.Context(lr); st = rsb.Stream(ctx, torch.cuda.current_stream().cuda_stream)
ops = dsort.CudaOps(rsb, ctx, st, "u32", n + n // 8 + (1 << 20)); ops.sorter.close()
ops.sorter = rsb.Sorter(ctx, "u32", n + n // 8 + (1 << 20), rsb.RS_NO_KEY_BUFFERS)   # profiling on
cap = n + n // 8 + (1 << 20); ops.enable_p2p(dist, cap)
shard = ops.empty(n); scratch = {k: ops.empty(cap) for k in ("recv", "c")}
ctx.fill_dataset(st, "u32", "uniform", shard.data_ptr(), n, seed=0x5EED, global_offset=rank * n, total_n=total)
for _ in range(3): dsort.distributed_sort(shard, ops, dist, "u32", scratch=scratch)
torch.cuda.synchronize(); ops.sorter.reset_runtimes()
for _ in range(3): dsort.distributed_sort(shard, ops, dist, "u32", scratch=scratch)
torch.cuda.synchronize(); rt = ops.sorter.runtimes()
# runtimes mix local sorts (plain histogram) and final sorts (gather+histogram): report max/min of the hist phase
print(f"[rank {rank}] hist phase min {rt['time_histo']['min']:.3f} max {rt['time_histo']['max']:.3f} ms | reorder avg {rt['time_reorder']['avg']:.3f} | total avg {rt['time_total']['avg']:.3f}", flush=True)
dist.barrier(); dist.destroy_process_group()
