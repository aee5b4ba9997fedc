"""Optimizer step for the connector parameters, on the device, right after the gradient all-reduce.

Reference (src/clip_whisper/trainer/clip_whisper_trainer.py): parameters whose name contains "bias" get no weight
decay, the rest `weight_decay` (:183-207); AdamW(betas=(0.9, 0.95), eps=1e-8) (:171-207); every step:
backward -> clip_grad_norm_(ALL parameters, max_grad_norm) -> optimizer.step() (:453-464).  The clip is a GLOBAL
norm over connector + LLM/LoRA parameters, so this class exposes the connector's squared-norm contribution
(`grad_sumsq()`, deterministic) and takes the other parameters' contribution back in `step(other_sumsq=...)`; the
clip coefficient is formed inside the AdamW kernel from the device scalar, with no host synchronisation.

The AdamW kernel can also emit the bf16, fusion-scaled copy of each weight into the packed projector operand,
which replaces the per-step weight pack of the forward pass.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch

from . import _lib as L
from .parallel import GradBucket


class ConnectorAdamW:
    def __init__(self, named_params: Iterable[Tuple[str, torch.Tensor]], bucket: Optional[GradBucket] = None, *,
                 lr: float = 2e-5, weight_decay: float = 0.01, betas=(0.9, 0.95), eps: float = 1e-8,
                 max_grad_norm: float = 0.5):
        self.params: Dict[str, torch.Tensor] = {}
        for n, p in named_params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError(f"{n}: connector master parameters must be contiguous fp32")
            self.params[n] = p
        if not self.params:
            raise ValueError("no parameters")
        dev = next(iter(self.params.values())).device
        self.bucket = bucket
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, tuple(betas), eps
        self.max_grad_norm = max_grad_norm
        self.exp_avg = {n: torch.zeros_like(p) for n, p in self.params.items()}
        self.exp_avg_sq = {n: torch.zeros_like(p) for n, p in self.params.items()}
        self.t = 0
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._ws = L.sumsq_workspace(dev)
        self.packed: Dict[str, Tuple[torch.Tensor, float]] = {}  # name -> (bf16 destination view, alpha)

    def decay_for(self, name: str) -> float:
        return 0.0 if "bias" in name else self.weight_decay  # clip_whisper_trainer.py:188

    def attach_packed(self, name: str, dst_view: torch.Tensor, alpha: float) -> None:
        """Have step() write bf16(alpha * param) into `dst_view` (a column slice of the packed projector operand)."""
        if dst_view.dtype != torch.bfloat16 or dst_view.shape != self.params[name].shape or dst_view.stride(1) != 1:
            raise ValueError("packed destination must be a bf16 [rows, cols] view with contiguous columns")
        self.packed[name] = (dst_view, alpha)

    def _grad(self, name: str) -> torch.Tensor:
        if self.bucket is not None and name in self.bucket.views:
            return self.bucket[name]
        g = self.params[name].grad
        if g is None:
            raise RuntimeError(f"{name} has no gradient")
        return g.contiguous()

    def grad_sumsq(self) -> torch.Tensor:
        """Device scalar: sum of squares of all connector gradients (after the all-reduce)."""
        if self.bucket is not None and set(self.bucket.views) == set(self.params):
            # alignment gaps of the bucket are zero-initialised and only ever all-reduced: they stay zero
            L.sumsq(self.bucket.flat, self._sumsq, self._ws)
        else:
            for i, n in enumerate(self.params):
                L.sumsq(self._grad(n), self._sumsq, self._ws, accumulate=i > 0)
        return self._sumsq

    def step(self, other_sumsq: Optional[torch.Tensor] = None, lr: Optional[float] = None) -> None:
        """clip_grad_norm_ (global norm incl. `other_sumsq` from the non-connector parameters) + AdamW."""
        self.t += 1
        clip = None
        if self.max_grad_norm and self.max_grad_norm > 0:
            clip = self.grad_sumsq()
            if other_sumsq is not None:
                clip = clip + other_sumsq.to(clip.device, torch.float32).reshape(1)
        lr = self.lr if lr is None else lr
        # the kernel updates the parameters through raw pointers, which does not bump torch's version counter: any
        # cached bf16 pack of the old weights (forward-only fast path of connector_ops.pack_projector) is stale now
        from .connector_ops import invalidate_pack_cache

        invalidate_pack_cache()
        for n, p in self.params.items():
            dst, alpha = self.packed.get(n, (None, 1.0))
            L.adamw_step(p.data if isinstance(p, torch.nn.Parameter) else p, self._grad(n), self.exp_avg[n],
                         self.exp_avg_sq[n], lr=lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                         weight_decay=self.decay_for(n), step=self.t, clip_sumsq=clip,
                         max_norm=float(self.max_grad_norm or 0.0), packed=dst, packed_alpha=alpha)
