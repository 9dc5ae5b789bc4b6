"""NCCL timings at the size of the c2 gradient buffer: all-reduce vs reduce-scatter + all-gather (in place / out of place).
    torchrun --nproc-per-node N tools/nccl_micro.py"""
import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 67_500_000 // (4 * world) * 4 * world
buf = torch.randn(n, device="cuda")
cnt = n // world
shard = torch.empty(cnt, device="cuda")
small = torch.zeros(1025, device="cuda")
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
res = {
    "all_reduce": t(lambda: dist.all_reduce(buf)),
    "reduce_scatter in place": t(lambda: dist.reduce_scatter_tensor(buf[rank * cnt:(rank + 1) * cnt], buf)),
    "reduce_scatter out of place": t(lambda: dist.reduce_scatter_tensor(shard, buf)),
    "all_gather in place": t(lambda: dist.all_gather_into_tensor(buf, buf[rank * cnt:(rank + 1) * cnt])),
    "all_gather out of place": t(lambda: dist.all_gather_into_tensor(buf, shard)),
    "all_reduce 1025 floats": t(lambda: dist.all_reduce(small)),
}
if rank == 0:
    for k, v in res.items():
        print("world %d  %-28s %8.1f us  (%.0f GB/s algbw)" % (world, k, v, n * 4 / v / 1e3))
dist.destroy_process_group()
