"""Why does pinned-host -> device bandwidth per GPU fall at N >= 4?  

Run under torchrun on an N-GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
        tools/h2d_numa_probe.py

Every rank prints the NUMA node of its GPU (sysfs), the CPUs / memory nodes it is allowed to use, and then all ranks at
once copy the e2e leg's 147.6 MB from pinned host memory to their GPU with
  (a) the pinned buffer allocated as bench.py does (wherever the kernel puts it),
  (b) the buffer allocated after binding the thread's memory policy (set_mempolicy, raw syscall: no libnuma in the image)
      and, if the cpuset allows it, its CPU affinity to the GPU's node,
  (c) the buffer deliberately placed on the OTHER node (the worst case, to show the sensitivity).
Output: one JSON line per rank and mode.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from audio_visual_llm_b200 import numa  # noqa: E402


def measure(buf_host, dev, iters=20):
    dst = torch.empty_like(buf_host, device=dev)
    stream = torch.cuda.Stream(device=dev)
    for _ in range(3):
        with torch.cuda.stream(stream):
            dst.copy_(buf_host, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(iters):
            dst.copy_(buf_host, non_blocking=True)
        e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / iters
    dist.barrier()
    return {"ms": round(ms, 4), "GBps": round(buf_host.nbytes / ms / 1e6, 2), "wall_ms": round(wall * 1e3 / iters, 4)}


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl" if world > 1 else "gloo", rank=rank, world_size=world,
                            init_method=None if world > 1 else "tcp://127.0.0.1:29533", device_id=dev if world > 1 else None)
    topo = numa.describe(local)
    topo["rank"] = rank
    print(json.dumps({"topology": topo}), flush=True)
    nbytes = 147621632
    out = {}

    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    out["default"] = dict(measure(host, dev), pages_on_node=numa.pages_on_node(host))
    del host

    node = topo["gpu_numa_node"]
    nodes = topo["nodes"]
    if node is not None and node >= 0 and len(nodes) > 1:
        for label, target in (("local_node", node), ("other_node", [n for n in nodes if n != node][0])):
            host, how = numa.pinned_empty((nbytes,), torch.uint8, target)
            out[label] = dict(measure(host, dev), node=target, bound=how, pinned=host.is_pinned(),
                              pages_on_node=numa.pages_on_node(host))
            numa.release(host)
            del host
    print(json.dumps({"rank": rank, "world": world, "h2d": out}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
