"""NUMA placement of the pinned host buffers that feed a GPU.

On a two-socket 8-GPU box a pinned buffer that sits on the other socket's memory is read over the inter-socket link,
and with >= 4 ranks copying at once that link -- not PCIe -- sets the host -> device rate.  Nothing in the image
provides libnuma / numactl, so this module talks to the kernel directly (sysfs for the topology, raw `mbind` /
`sched_setaffinity` for the placement, `/proc/self/numa_maps` to verify) and degrades to "no binding" -- with the reason
in `how` -- when the container forbids those calls.  No reference counterpart: the reference feeds the device from
pageable memory on one GPU (clip_whisper_trainer.py:655-657).

    node = numa.gpu_numa_node(torch.cuda.current_device())
    buf, how = numa.pinned_empty((B, T, D), torch.bfloat16, node)   # page-locked, pages on `node`
"""
import contextlib
import ctypes
import glob
import os
import platform
import re

import torch

_SYSCALLS = {"x86_64": {"mbind": 237, "set_mempolicy": 238, "get_mempolicy": 239, "move_pages": 279},
             "aarch64": {"mbind": 235, "get_mempolicy": 236, "set_mempolicy": 237, "move_pages": 239}}
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2
MPOL_MF_MOVE = 2
_PAGE = os.sysconf("SC_PAGE_SIZE") if hasattr(os, "sysconf") else 4096


def _libc():
    return ctypes.CDLL(None, use_errno=True)


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def nodes():
    """NUMA node ids the kernel exposes (sorted); [] when sysfs has none."""
    out = []
    for path in glob.glob("/sys/devices/system/node/node[0-9]*"):
        m = re.search(r"node(\d+)$", path)
        if m:
            out.append(int(m.group(1)))
    return sorted(out)


def node_cpus(node):
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            return _parse_cpulist(f.read())
    except OSError:
        return set()


def gpu_numa_node(index):
    """NUMA node of CUDA device `index` from sysfs, or None when it is not exposed (single node, VM, -1)."""
    props = torch.cuda.get_device_properties(index)
    try:
        addr = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    except AttributeError:
        return None
    try:
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as f:
            node = int(f.read().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def describe(index):
    """Topology facts for logs: the GPU's node, all nodes with their CPU counts, what this process may use."""
    allowed = sorted(os.sched_getaffinity(0))
    mems = None
    try:
        with open("/proc/self/status") as f:
            for line in f:
                if line.startswith("Mems_allowed_list:"):
                    mems = line.split(":", 1)[1].strip()
    except OSError:
        pass
    ns = nodes()
    return {"gpu": index, "gpu_numa_node": gpu_numa_node(index), "nodes": ns,
            "node_cpus": {n: len(node_cpus(n)) for n in ns},
            "allowed_cpus": len(allowed), "allowed_cpus_per_node": {n: len(node_cpus(n) & set(allowed)) for n in ns},
            "mems_allowed": mems}


def _mbind(ptr, nbytes, node, mode=MPOL_BIND):
    """mbind the pages that hold [ptr, ptr+nbytes) to `node`, moving the ones already touched.  Returns errno (0 = ok)."""
    nr = _SYSCALLS.get(platform.machine(), {}).get("mbind")
    if nr is None:
        return -1
    start = ptr & ~(_PAGE - 1)
    length = ((ptr + nbytes + _PAGE - 1) & ~(_PAGE - 1)) - start
    mask = (ctypes.c_ulong * 2)(0, 0)
    mask[node // 64] = 1 << (node % 64)
    libc = _libc()
    libc.syscall.restype = ctypes.c_long
    rc = libc.syscall(ctypes.c_long(nr), ctypes.c_void_p(start), ctypes.c_ulong(length), ctypes.c_int(mode),
                      ctypes.byref(mask), ctypes.c_ulong(129), ctypes.c_uint(MPOL_MF_MOVE))
    return 0 if rc == 0 else (ctypes.get_errno() or -1)


@contextlib.contextmanager
def cpus_of_node(node):
    """Run the calling thread on `node`'s CPUs (those the cpuset allows) for the duration; yields how many it got."""
    tid = 0
    before = os.sched_getaffinity(tid)
    want = node_cpus(node) & before
    if want:
        os.sched_setaffinity(tid, want)
    try:
        yield len(want)
    finally:
        os.sched_setaffinity(tid, before)


def pages_on_node(tensor):
    """{node: pages} of the mapping that holds `tensor`'s first byte (from /proc/self/numa_maps), or None."""
    ptr = tensor.data_ptr()
    best = None
    try:
        with open("/proc/self/numa_maps") as f:
            for line in f:
                head, _, rest = line.partition(" ")
                try:
                    start = int(head, 16)
                except ValueError:
                    continue
                if start <= ptr and (best is None or start > best[0]):
                    best = (start, rest)
    except OSError:
        return None
    if best is None:
        return None
    return {int(k): int(v) for k, v in re.findall(r"N(\d+)=(\d+)", best[1])}


_REGISTERED = {}   # data_ptr -> tensor: registered buffers stay alive (views may outlive the handle) until release()


def release(tensor):
    """Unregister and drop a buffer made by `pinned_empty` (no views of it may be in use)."""
    ptr = tensor.data_ptr()
    if _REGISTERED.pop(ptr, None) is not None:
        torch.cuda.cudart().cudaHostUnregister(ptr)


def pinned_empty(shape, dtype, node):
    """A page-locked CPU tensor whose pages sit on NUMA node `node`; returns (tensor, how).

    how: "mbind" (pages bound with the mbind syscall), "affinity:<n>" (first touch from <n> CPUs of the node because
    mbind is not permitted), or "unbound:<reason>" (plain `pin_memory()`; `node` None, a single node, or no CPU of the
    node in the cpuset).  The memory is registered with cudaHostRegister, so `.is_pinned()` holds and async copies
    from it are truly asynchronous; it stays registered until `release()`."""
    if node is None or len(nodes()) < 2:
        return torch.empty(shape, dtype=dtype).pin_memory(), "unbound:no NUMA choice"
    t = torch.empty(shape, dtype=dtype)
    ptr, nbytes = t.data_ptr(), t.nbytes
    if nbytes == 0:
        return t.pin_memory(), "unbound:empty"
    err = _mbind(ptr, nbytes, node)
    if err == 0:
        how = "mbind"
        ctypes.memset(ptr, 0, nbytes)
    else:
        with cpus_of_node(node) as n:
            if n == 0:
                return t.pin_memory(), f"unbound:mbind errno {err}, no CPU of node {node} in the cpuset"
            ctypes.memset(ptr, 0, nbytes)   # first touch by THIS thread (torch's fill_ would fan out over a pool)
            how = f"affinity:{n}"
    rc = torch.cuda.cudart().cudaHostRegister(ptr, nbytes, 0)
    if int(rc) != 0:
        return t.pin_memory(), f"unbound:cudaHostRegister error {int(rc)}"
    _REGISTERED[ptr] = t
    return t, how


def pin_like(src, node):
    """Copy of CPU tensor `src` in pinned memory on `node`; returns (tensor, how)."""
    dst, how = pinned_empty(tuple(src.shape), src.dtype, node)
    dst.copy_(src)
    return dst, how
