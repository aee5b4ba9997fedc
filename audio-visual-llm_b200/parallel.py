"""Data parallelism for the connector: one process per GPU, batch sharded by sample, weights replicated.

The path has exactly one collective: all-reduce(mean) of the projector gradients between
backward() and the trainer's clip_grad_norm_ (clip_whisper_trainer.py:454-458).  The reference has no
distributed code at all, so this is new.  All gradients live in ONE flat fp32 bucket (~100 MB at Llama-2-7B width);
the dW GEMM and the bias column-sum kernels write straight into views of the bucket, so there is no flatten /
unflatten copy.  Rank-local work (gather / GEMMs / splice) never communicates.

Two ways to reduce the bucket:
  * `PeerMemory` (default on NCCL process groups of <= 8 ranks): the bucket lives in memory every rank's process maps
    -- CUDA IPC peer mappings, or an NVSwitch multicast object -- and the dW GEMM launch all-reduces it itself
    (`avc_proj_bwd_dw_allreduce`); torch.distributed only carries handles, agreements and barriers during setup.
  * `GradBucket.allreduce()`: one NCCL (or gloo, for the CPU tests) all-reduce after the backward.
"""
from __future__ import annotations

import os
import secrets
import socket
import threading
import time
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def _fd_socket_address(tag: str) -> str:
    """Address of the fd-passing socket.  A leading NUL selects Linux's ABSTRACT socket namespace: no file is created,
    so nothing can be pre-created, unlinked or left behind in a shared /tmp; `tag` carries a random per-job token
    (broadcast through the process group) so that a local user cannot squat the name either."""
    return f"\0avc_{tag}_{os.getuid()}"


def _serve_fd(fd: int, nclients: int, tag: str) -> threading.Thread:
    """Hand a file descriptor to `nclients` local processes over an AF_UNIX socket (SCM_RIGHTS).  Raises OSError if
    the socket cannot be bound (the caller turns that into an all-rank fallback)."""
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    try:
        srv.bind(_fd_socket_address(tag))
        srv.listen(max(1, nclients))
        srv.settimeout(60.0)
    except OSError:
        srv.close()
        raise

    def run():
        try:
            for _ in range(nclients):
                conn, _ = srv.accept()
                socket.send_fds(conn, [b"fd"], [fd])
                conn.close()
        except OSError:
            pass  # a client that never came: the import on that rank fails and every rank falls back together
        finally:
            srv.close()

    t = threading.Thread(target=run, daemon=True)
    t.start()
    return t


def _fetch_fd(tag: str, timeout_s: float = 30.0) -> int:
    addr = _fd_socket_address(tag)
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    deadline = time.time() + timeout_s
    while True:
        try:
            c.connect(addr)
            break
        except (ConnectionRefusedError, FileNotFoundError):
            if time.time() > deadline:
                c.close()
                raise TimeoutError(f"no fd server at {addr!r}")
            time.sleep(0.01)
    _, fds, _, _ = socket.recv_fds(c, 16, 1)
    c.close()
    if not fds:
        raise OSError("fd server closed the connection without sending a descriptor")
    return fds[0]


class PeerMemory:
    """This rank's gradient bucket + flag area in memory that every other rank's process maps (CUDA IPC, peer access
    over NVLink), and the other ranks' blocks mapped here.  It is what `avc_proj_bwd_dw_allreduce` -- the dW GEMM with
    the gradient all-reduce fused into the same kernel -- reads and writes; NCCL is not involved in that path.
    Handles are exchanged once through the process group (`all_gather_object`).

    `next_epoch()` must be called once per fused launch, in the same order on every rank."""

    def __init__(self, nfloats: int, device, process_group=None, timeout_s: float = 20.0, multimem: bool = False):
        """multimem=True: the bucket is bound to an NVSwitch multicast object (`avc_mc_*`), and the fused launch reduces
        with multimem.ld_reduce / multimem.st instead of peer loads / stores (falls back to the peer mapping, on every
        rank alike, when the devices do not support multicast)."""
        from . import _lib as L

        self._L = L
        self.device = torch.device(device)
        self.group = process_group
        ddp = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(process_group) if ddp else 1
        self.rank = dist.get_rank(process_group) if ddp else 0
        if self.world > L.COMM_MAX_WORLD:
            raise L.ConnectorError(f"peer-memory all-reduce supports up to {L.COMM_MAX_WORLD} ranks, got {self.world}")
        self.nfloats = nfloats
        self.mc = None  # AvcMcBucket when the multicast transport is in use
        if multimem:
            ok = torch.tensor([1 if L.mc_supported(self.device.index or 0) else 0], dtype=torch.int32, device=self.device)
            if self.world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
            multimem = bool(int(ok.item()))
        with torch.cuda.device(self.device):
            if multimem:
                self.mc = self._setup_multicast(L, nfloats * 4, process_group)
            self._own = (self.mc.ptr if self.mc is not None else L.comm_alloc(nfloats * 4),
                         L.comm_alloc(L.comm_flag_bytes()))
            # multicast transport: the peers' buckets are reached through the multicast address, only flags are mapped
            mine = (None if self.mc is not None else L.comm_export(self._own[0]), L.comm_export(self._own[1]))
            handles = [mine]
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, mine, group=process_group)
            self.bucket_ptrs, self.flag_ptrs, self._opened = [], [], []
            err = None
            try:
                for r, (hb, hf) in enumerate(handles):
                    if r == self.rank:
                        pb, pf = self._own
                    else:
                        pb = 0
                        if hb is not None:
                            pb = L.comm_open(hb)
                            self._opened.append(pb)
                        pf = L.comm_open(hf)
                        self._opened.append(pf)
                    self.bucket_ptrs.append(pb)
                    self.flag_ptrs.append(pf)
            except L.ConnectorError as e:  # e.g. no peer access between two GPUs
                err = e
            if self.world > 1:
                # every rank must reach the same verdict, or some would wait for flags that never come
                ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
                if int(ok.item()) == 0 and err is None:
                    err = L.ConnectorError("another rank could not map the peer buffers")
            if err is not None:
                for p in self._opened:
                    L.comm_close(p)
                self._free_own()
                self._opened = []
                raise L.ConnectorError(f"peer-memory setup failed on rank {self.rank}: {err}")
        self.flat = L.as_tensor(self._own[0], nfloats, torch.float32, self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        # The bias-sum kernel is launched while the fused GEMM already waits for it; CUDA loads kernels lazily and a
        # load can block behind running kernels, so run it once now (avc_comm_alloc also preloads it).
        with torch.cuda.device(self.device):
            warm = torch.zeros(1, 8, 64, dtype=torch.bfloat16, device=self.device)
            out = torch.empty(64, dtype=torch.float32, device=self.device)
            L.colsum(warm, out, None, L.colsum_workspace(64, self.device))
            torch.cuda.synchronize(self.device)
        self.timeout_ns = int(timeout_s * 1e9)
        self.epoch = 0
        if self.world > 1:
            dist.barrier(group=process_group)  # every rank has mapped every block before the first launch

    def _setup_multicast(self, L, nbytes: int, group):
        """rank 0 creates the multicast object and serves its fd; every rank adds its device, allocates + binds.
        Every stage ends with an agreement (all-reduce MIN of "it worked here"), which is also the barrier the driver
        API needs between the stages; if any rank fails, ALL ranks return None and use the peer mapping instead."""

        def agree(ok: bool) -> bool:
            if self.world == 1:
                return ok
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(int(t.item()))

        def attempt(fn):
            try:
                return fn(), None
            except (L.ConnectorError, OSError, TimeoutError) as e:
                return None, e

        padded, err = attempt(lambda: L.mc_padded_bytes(self.world, nbytes))
        created, fd = None, -1
        if err is None and self.rank == 0:
            created, err = attempt(lambda: L.mc_create(self.world, padded))
        if not agree(err is None):
            return None
        handle = None
        # the socket name carries a random per-job token chosen by rank 0 (nobody can pre-create or guess it)
        token = [secrets.token_hex(8) if self.rank == 0 else None]
        if self.world > 1:
            dist.broadcast_object_list(token, src=dist.get_global_rank(group, 0) if group is not None else 0,
                                       group=group)
        tag = f"mc_{token[0]}"
        server = None
        if self.rank == 0:
            handle, fd = created
            server, err = attempt(lambda: _serve_fd(fd, self.world - 1, tag))
        listening = agree(err is None)       # the socket is listening (or every rank falls back together)
        if self.rank == 0:
            if server is not None and listening:
                server.join(timeout=60)
            os.close(fd)
        elif listening:

            def imp():
                f = _fetch_fd(tag)
                try:
                    return L.mc_import(f)
                finally:
                    os.close(f)

            handle, err = attempt(imp)
        if not listening or not agree(err is None):
            return None
        _, err = attempt(lambda: L.mc_add_device(handle))
        if not agree(err is None):           # every device is in the object before anyone binds memory to it
            return None
        b, err = attempt(lambda: L.mc_bucket_alloc(handle, padded))
        if not agree(err is None):           # every rank's memory is bound before the first multimem access
            if b is not None:
                L.mc_bucket_free(b)
            return None
        return b

    def _free_own(self):
        L = self._L
        if self._own is None:
            return
        if self.mc is not None:
            L.mc_bucket_free(self.mc)
            self.mc = None
        else:
            L.comm_free(self._own[0])
        L.comm_free(self._own[1])
        self._own = None

    def next_epoch(self):
        """Descriptor of the next fused launch (the epoch number is the protocol's only per-step state)."""
        L = self._L
        self.epoch += 1
        c = L.AvcComm()
        c.world, c.rank, c.epoch = self.world, self.rank, self.epoch
        for r in range(self.world):
            c.bucket[r] = self.bucket_ptrs[r] or None
            c.flags[r] = self.flag_ptrs[r]
        c.mc_bucket = self.mc.mc_ptr if self.mc is not None else None
        c.status = self.status.data_ptr()
        c.timeout_ns = self.timeout_ns
        c.bucket_bytes = self.nfloats * 4
        return c

    def check(self) -> None:
        """Raise if a fused launch gave up waiting for a peer (synchronises)."""
        code = int(self.status.item())
        if code != 0:
            raise self._L.ConnectorError(f"fused gradient all-reduce failed on rank {self.rank}: status {code} "
                                         "(1 = a peer did not reach the same launch in time)")

    def close(self) -> None:
        L = self._L
        if self._own is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)  # nobody unmaps / frees while a peer may still touch the memory
        with torch.cuda.device(self.device):
            for p in self._opened:
                L.comm_close(p)
            self.flat = None
            self._free_own()
        self._opened = []


class GradBucket:
    """Flat fp32 gradient buffer with named views; `allreduce()` averages it over the process group.

    peer=True puts the buffer in `PeerMemory` so that the dW GEMM can all-reduce it itself (`self.peer`)."""

    def __init__(self, shapes: Dict[str, Tuple[int, ...]], device, process_group=None, align_elems: int = 64,
                 peer: bool = False, multimem: bool = False):
        self.views: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        offs, total = {}, 0
        for name, shape in shapes.items():
            n = 1
            for d in shape:
                n *= d
            offs[name] = (total, n, shape)
            total += (n + align_elems - 1) // align_elems * align_elems  # keep every view 256-byte aligned
        self.peer: Optional[PeerMemory] = PeerMemory(total, device, process_group, multimem=multimem) if peer else None
        self.flat = self.peer.flat if peer else torch.zeros(total, dtype=torch.float32, device=device)
        self._pads = []
        for name, (o, n, shape) in offs.items():
            self.views[name] = self.flat[o:o + n].view(*shape)
            end = (n + align_elems - 1) // align_elems * align_elems
            if end > n:
                self._pads.append(self.flat[o + n:o + end])
        self.group = process_group
        self.comm_stream: Optional[torch.cuda.Stream] = None

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.views[name]

    def zero_padding(self) -> None:
        """Alignment gaps between views stay zero so that whole-bucket reductions (norms) see only gradients."""
        for pad in self._pads:
            pad.zero_()

    def world_size(self) -> int:
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    def allreduce(self, async_op: bool = False, prescaled: bool = False):
        """flat <- mean over ranks.  No-op for a single process.

        prescaled=True: the producer already divided its gradients by the world size (the dW GEMM / bias-sum
        epilogues take the factor for free), so a plain SUM is issued -- which, unlike AVG, lets NCCL use the
        in-switch NVLS reduction on NVSwitch systems."""
        ws = self.world_size()
        if ws == 1:
            return None
        if dist.get_backend(self.group) == "nccl":
            op = dist.ReduceOp.SUM if prescaled else dist.ReduceOp.AVG
            return dist.all_reduce(self.flat, op=op, group=self.group, async_op=async_op)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=False)  # gloo (CPU tests)
        if not prescaled:
            self.flat.div_(ws)
        return work

    def span(self, first: str, last: str) -> torch.Tensor:
        """Contiguous slice of the flat buffer covering views `first` .. `last` (in bucket order)."""
        names = list(self.views)
        i, j = names.index(first), names.index(last)
        lo = self.views[names[i]].data_ptr() - self.flat.data_ptr()
        hi = self.views[names[j]].data_ptr() - self.flat.data_ptr() + self.views[names[j]].numel() * 4
        return self.flat[lo // 4: hi // 4]

    def allreduce_span(self, first: str, last: str, prescaled: bool = False):
        """Average only views `first` .. `last`; used to overlap the reduction of finished gradients with the
        kernels that still produce the rest."""
        ws = self.world_size()
        if ws == 1:
            return
        t = self.span(first, last)
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.SUM if prescaled else dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            if not prescaled:
                t.div_(ws)

    def attach(self, named_params: Sequence[Tuple[str, torch.nn.Parameter]]) -> None:
        """Point each parameter's .grad at its bucket view (so optimizers / clip_grad_norm_ see the reduced grads)."""
        for name, p in named_params:
            if name in self.views:
                p.grad = self.views[name]


class FusedGradSync:
    """Data-parallel gradient sync for the PUBLIC API (`fused_connector(..., grad_sync=...)`,
    `ClipWhisperModel.enable_data_parallel()`): the four projector parameters' .grad are views of one peer-mapped
    bucket, the autograd backward writes dW / db straight into them and the dW GEMM launch all-reduces the bucket
    itself (`avc_proj_bwd_dw_db_allreduce`) -- no NCCL call, no flatten / unflatten copy, and nothing is returned
    through autograd for those parameters (so gradients are OVERWRITTEN every backward, not accumulated: call the
    optimizer after every backward, as the reference trainer does, clip_whisper_trainer.py:453-464).

    After `backward()` every rank's `.grad` holds the MEAN over ranks.  Falls back to a plain local bucket + one NCCL
    all-reduce (`finish()`) when peer mapping is unavailable."""

    NAMES = ("audio_connector.linear.weight", "video_connector.linear.weight", "audio_connector.linear.bias",
             "video_connector.linear.bias")

    def __init__(self, wa: Optional[torch.Tensor], ba: Optional[torch.Tensor], wv: Optional[torch.Tensor],
                 bv: Optional[torch.Tensor], process_group=None, multimem: Optional[bool] = None):
        from . import _lib as L

        params = dict(zip(self.NAMES, (wa, wv, ba, bv)))
        shapes = {n: tuple(p.shape) for n, p in params.items() if p is not None}
        dev = next(p for p in params.values() if p is not None).device
        ddp = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(process_group) if ddp else 1
        if multimem is None:
            multimem = os.environ.get("AVC_COMM_MULTIMEM", "1") == "1"
        peer = self.world <= L.COMM_MAX_WORLD and (not ddp or dist.get_backend(process_group) == "nccl")
        try:
            self.bucket = GradBucket(shapes, dev, process_group=process_group, peer=peer, multimem=multimem and peer)
        except L.ConnectorError:
            if not peer:
                raise
            self.bucket = GradBucket(shapes, dev, process_group=process_group)
        self.fused = self.bucket.peer is not None
        self.bucket.attach([(n, p) for n, p in params.items() if p is not None])

    def views(self, use_a: bool, use_v: bool):
        g = self.bucket.views
        return (g.get(self.NAMES[0]) if use_a else None, g.get(self.NAMES[2]) if use_a else None,
                g.get(self.NAMES[1]) if use_v else None, g.get(self.NAMES[3]) if use_v else None)

    def next_epoch(self):
        if not self.fused:
            raise RuntimeError("no peer-mapped bucket: use finish() after backward")
        return self.bucket.peer.next_epoch()

    def check(self) -> None:
        if self.fused:
            self.bucket.peer.check()

    def close(self) -> None:
        if self.fused:
            self.bucket.peer.close()


def shard_batch(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous sample range [lo, hi) of this rank (even split, remainder to the low ranks)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balance_ragged(token_counts: Sequence[int], world: int) -> List[List[int]]:
    """Assign samples to ranks so that the per-rank sum of fused tokens is balanced (greedy longest-first).
    Used for ragged batches (cfg4), where balancing by sample count would leave ranks idle."""
    order = sorted(range(len(token_counts)), key=lambda i: -token_counts[i])
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += token_counts[i]
    for lst in out:
        lst.sort()
    return out
