"""NVSwitch multicast (NVLS) probe, run under torchrun on 2+ GPUs of one box.

Question for the next step of the fused dW + all-reduce kernel: can the gradient bucket be bound
to a CUDA multicast object on this pool, so that the comm warps use `multimem.ld_reduce` (sum of every rank's copy, added
inside the switch) and `multimem.st` (one store that lands in every rank's copy) instead of `world` peer loads and
`world` peer stores?  That moves ~100 MB per direction per GPU instead of 176 MB at 8 ranks.

The probe (driver API through cuda-python, nothing from the product library):
  1. CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED on every rank's device
  2. rank 0 creates the multicast object and exports it as a POSIX fd; the fd travels over an AF_UNIX socket
     (SCM_RIGHTS); every rank imports it and adds its device
  3. every rank cuMemCreate's its buffer, binds it to the multicast object, maps the unicast and multicast addresses
  4. every rank also exports its buffer's fd and maps every peer's buffer (the VMM replacement of the CUDA IPC mapping
     the product uses today)
  5. a `multimem.ld_reduce` + `multimem.st` all-reduce kernel (each rank reduces its 1 / world slice) is checked against
     the expected sum and timed on a 100.7 MB buffer
Prints one line per stage on rank 0; exits non-zero on the first failure.
"""
import ctypes
import os
import socket
import subprocess
import sys
import tempfile
import threading
import time

import torch
import torch.distributed as dist

try:
    from cuda.bindings import driver as cu
except ImportError:  # older cuda-python
    from cuda import cuda as cu

KERNEL = r"""
extern "C" __global__ void mc_allreduce(float* mc, long long start_vec, long long nvec) {
  for (long long i = start_vec + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < start_vec + nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float4 r;
    float* p = mc + 4 * i;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(r.x), "f"(r.y), "f"(r.z), "f"(r.w) : "memory");
  }
}
"""

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
POSIX_FD = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR


def ck(res, what):
    err = res[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"rank {rank}: {what} failed: {err}")
    return res[1] if len(res) == 2 else res[1:]


def say(msg):
    if rank == 0:
        print("[mc_probe]", msg, file=sys.stderr, flush=True)


def sock_path(r):
    return os.path.join(tempfile.gettempdir(), f"avc_mc_{os.environ.get('MASTER_PORT', '0')}_{r}.sock")


def serve_fds(fds, nclients):
    """Hand `fds` to `nclients` connecting processes (SCM_RIGHTS); returns the serving thread."""
    path = sock_path(rank)
    if os.path.exists(path):
        os.unlink(path)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    srv.listen(nclients)

    def run():
        for _ in range(nclients):
            conn, _ = srv.accept()
            socket.send_fds(conn, [b"fds"], list(fds))
            conn.close()
        srv.close()
        os.unlink(path)

    t = threading.Thread(target=run, daemon=True)
    t.start()
    return t


def fetch_fds(r, n):
    path = sock_path(r)
    for _ in range(500):
        if os.path.exists(path):
            break
        time.sleep(0.01)
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    for _ in range(500):
        try:
            c.connect(path)
            break
        except (ConnectionRefusedError, FileNotFoundError):
            time.sleep(0.01)
    _, fds, _, _ = socket.recv_fds(c, 16, n)
    c.close()
    return fds


class DevBlock:
    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


def main():
    torch.cuda.set_device(local)
    dev_t = torch.device("cuda", local)
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev_t)
    torch.zeros(1, device=dev_t)  # primary context current on this thread
    ck(cu.cuInit(0), "cuInit")
    dev = ck(cu.cuDeviceGet(local), "cuDeviceGet")
    sup = ck(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev), "attr")
    allsup = torch.tensor([int(sup)], device=dev_t)
    dist.all_reduce(allsup, op=dist.ReduceOp.MIN)
    say(f"1. CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED on all {world} ranks: {int(allsup.item())}")
    if int(allsup.item()) == 0:
        say("RESULT: multicast not supported on this box")
        return 2

    nfloats = 25169920  # the cfg2 gradient bucket
    prop = cu.CUmulticastObjectProp()
    prop.numDevices = world
    prop.handleTypes = POSIX_FD
    prop.flags = 0
    prop.size = nfloats * 4
    gran = ck(cu.cuMulticastGetGranularity(prop, cu.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED),
              "cuMulticastGetGranularity")
    size = (nfloats * 4 + gran - 1) // gran * gran
    prop.size = size
    say(f"   multicast granularity {gran} bytes, object size {size} bytes")

    # ---- 2. multicast object: created by rank 0, imported by the others
    if rank == 0:
        mc = ck(cu.cuMulticastCreate(prop), "cuMulticastCreate")
        fd = ck(cu.cuMemExportToShareableHandle(mc, POSIX_FD, 0), "export multicast fd")
        t = serve_fds([int(fd)], world - 1)
        dist.barrier()
        t.join()
    else:
        dist.barrier()
        (fd,) = fetch_fds(0, 1)
        mc = ck(cu.cuMemImportFromShareableHandle(fd, POSIX_FD), "import multicast fd")
    ck(cu.cuMulticastAddDevice(mc, dev), "cuMulticastAddDevice")
    dist.barrier()
    say("2. multicast object created, shared as a POSIX fd, every device added")

    # ---- 3. local buffer: create, bind, map unicast + multicast
    aprop = cu.CUmemAllocationProp()
    aprop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    aprop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    aprop.location.id = local
    aprop.requestedHandleTypes = POSIX_FD
    mem = ck(cu.cuMemCreate(size, aprop, 0), "cuMemCreate")
    ck(cu.cuMulticastBindMem(mc, 0, mem, 0, size, 0), "cuMulticastBindMem")
    dist.barrier()
    acc = cu.CUmemAccessDesc()
    acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = local
    acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE

    def map_handle(handle, what):
        va = ck(cu.cuMemAddressReserve(size, gran, 0, 0), f"reserve {what}")
        ck(cu.cuMemMap(va, size, 0, handle, 0), f"map {what}")
        ck(cu.cuMemSetAccess(va, size, [acc], 1), f"access {what}")
        return int(va)

    uc = map_handle(mem, "unicast")
    mcva = map_handle(mc, "multicast")
    say("3. local buffer bound to the multicast object; unicast and multicast addresses mapped")

    # ---- 4. peers' buffers through exported fds (VMM peer mapping)
    my_fd = ck(cu.cuMemExportToShareableHandle(mem, POSIX_FD, 0), "export buffer fd")
    t = serve_fds([int(my_fd)], world - 1)
    dist.barrier()
    peers = {}
    for r in range(world):
        if r == rank:
            continue
        (pfd,) = fetch_fds(r, 1)
        ph = ck(cu.cuMemImportFromShareableHandle(pfd, POSIX_FD), f"import rank {r}'s buffer")
        peers[r] = map_handle(ph, f"rank {r}'s buffer")
    t.join()
    dist.barrier()
    say(f"4. every peer's buffer mapped through its POSIX fd ({len(peers)} peers)")

    # ---- 5. multimem all-reduce kernel
    with tempfile.TemporaryDirectory() as td:
        src, cubin = os.path.join(td, "k.cu"), os.path.join(td, "k.cubin")
        open(src, "w").write(KERNEL)
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", src, "-o", cubin], check=True)
        image = open(cubin, "rb").read()
    mod = ck(cu.cuModuleLoadData(image), "cuModuleLoadData")
    fn = ck(cu.cuModuleGetFunction(mod, b"mc_allreduce"), "cuModuleGetFunction")
    local_t = torch.as_tensor(DevBlock(uc, nfloats), device=dev_t)
    local_t.fill_(float(rank + 1))
    peer_sum = None
    if peers:
        r0 = sorted(peers)[0]
        torch.cuda.synchronize()
        dist.barrier()
        import numpy as np

        host = np.zeros(1024, dtype=np.float32)
        ck(cu.cuMemcpyDtoH(host, peers[r0], 4096), "cuMemcpyDtoH through the peer mapping")
        peer_sum = float(host.sum())
        assert peer_sum == 1024.0 * (r0 + 1), (peer_sum, r0)
    torch.cuda.synchronize()
    dist.barrier()
    nvec = nfloats // 4
    per = (nvec + world - 1) // world
    start = rank * per
    mine = max(0, min(per, nvec - start))
    stream = torch.cuda.current_stream().cuda_stream
    args = ((mcva, start, mine), (ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong))

    def launch(blocks):
        ck(cu.cuLaunchKernel(fn, blocks, 1, 1, 256, 1, 1, 0, stream, args, 0), "cuLaunchKernel")

    launch(296)
    torch.cuda.synchronize()
    dist.barrier()
    expect = float(sum(range(1, world + 1)))
    got = local_t[::4099].cpu()
    ok = bool((got == expect).all())
    allok = torch.tensor([int(ok)], device=dev_t)
    dist.all_reduce(allok, op=dist.ReduceOp.MIN)
    say(f"5. multimem.ld_reduce + multimem.st all-reduce: every rank holds {expect} everywhere: {bool(allok.item())}")
    if not allok.item():
        say(f"RESULT: wrong values, e.g. {got[:4].tolist()} on rank 0")
        return 3
    for blocks in (32, 74, 148, 296, 592):
        local_t.fill_(1e-3)
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            launch(blocks)
        e.record()
        torch.cuda.synchronize()
        t_ms = torch.tensor([s.elapsed_time(e) / 10], device=dev_t)
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        say(f"   {blocks:4d} CTAs x 256 threads: {float(t_ms):.4f} ms per 100.7 MB all-reduce "
            f"(algbw {nfloats * 4 / float(t_ms) / 1e6:.0f} GB/s)")
    say("RESULT: multicast works on this box")
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
