"""Time the gradient all-reduce alone (100.7 MB fp32 bucket and its two spans) under torchrun."""
import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sizes = (25169920,) if os.environ.get("PROBE_SHORT") else (25169920, 16777216, 8392704)
for opname, op in (("AVG", dist.ReduceOp.AVG), ("SUM", dist.ReduceOp.SUM)):
    for n in sizes:
        x = torch.randn(n, device=dev)
        for _ in range(5): dist.all_reduce(x, op=op)
        torch.cuda.synchronize(); dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20): dist.all_reduce(x, op=op)
        e.record(); torch.cuda.synchronize()
        if rank == 0: print(f"allreduce {opname} {n*4/1e6:.1f} MB: {s.elapsed_time(e)/20:.4f} ms", file=sys.stderr)
dist.destroy_process_group()
