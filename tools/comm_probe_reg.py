"""All-reduce of the 100.7 MB gradient bucket with / without NCCL user-buffer registration (torchrun)."""
import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 25169920
def bench(x, op, tag):
    for _ in range(5): dist.all_reduce(x, op=op)
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): dist.all_reduce(x, op=op)
    e.record(); torch.cuda.synchronize()
    if rank == 0: print(f"allreduce {tag}: {s.elapsed_time(e)/20:.4f} ms", file=sys.stderr)
plain = torch.randn(n, device=dev)
bench(plain, dist.ReduceOp.AVG, "plain AVG")
bench(plain, dist.ReduceOp.SUM, "plain SUM")
try:
    backend = dist.group.WORLD._get_backend(dev)
    pool = torch.cuda.MemPool(backend.mem_allocator)
    with torch.cuda.use_mem_pool(pool):
        reg = torch.randn(n, device=dev)
    backend.register_mem_pool(pool)
    ref = reg.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    chk = reg.clone()
    with torch.cuda.use_mem_pool(pool):
        chk2 = torch.empty(n, device=dev)
    chk2.copy_(chk)
    dist.all_reduce(chk2, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize()
    err = float((chk2 - ref).abs().max())
    if rank == 0: print(f"registered SUM max abs diff vs plain: {err:.3e}", file=sys.stderr)
    bench(reg, dist.ReduceOp.SUM, "registered SUM")
    bench(reg, dist.ReduceOp.AVG, "registered AVG")
except Exception as ex:
    if rank == 0: print("registration failed:", repr(ex)[:300], file=sys.stderr)
dist.destroy_process_group()
