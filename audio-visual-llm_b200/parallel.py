"""Data parallelism for the connector: one process per GPU, batch sharded by sample, weights replicated.

The path has exactly one collective: all-reduce(mean) of the projector gradients between
backward() and the trainer's clip_grad_norm_ (clip_whisper_trainer.py:454-458).  The reference has no
distributed code at all, so this is new.  All gradients live in ONE flat fp32 bucket so the collective is a
single NCCL call (~100 MB at Llama-2-7B width: latency- not link-bound over NVSwitch); the dW GEMM and the bias
column-sum kernels write straight into views of the bucket, so there is no flatten/unflatten copy.
Rank-local work (gather / GEMMs / splice) never communicates.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class GradBucket:
    """Flat fp32 gradient buffer with named views; `allreduce()` averages it over the process group."""

    def __init__(self, shapes: Dict[str, Tuple[int, ...]], device, process_group=None, align_elems: int = 64):
        self.views: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        offs, total = {}, 0
        for name, shape in shapes.items():
            n = 1
            for d in shape:
                n *= d
            offs[name] = (total, n, shape)
            total += (n + align_elems - 1) // align_elems * align_elems  # keep every view 256-byte aligned
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
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
