"""Train-time length adaptation of the reference (`_adaptive_projection`, `_adapt_mask`,
clip_whisper_model.py:621-736) as one sparse row-mixing kernel (forward) and its transpose (backward).

  source longer than target:  AdaptiveAvgPool1d windows [floor(i*S/L), ceil((i+1)*S/L)), weight 1/len  (:641-656)
  source shorter than target: F.interpolate(mode="linear", align_corners=True): two taps per row       (:671-676)
"""
from __future__ import annotations

from functools import lru_cache

import torch

from . import _lib as L


@lru_cache(maxsize=64)
def _taps(S: int, Lt: int):
    """CSR (row_ptr, col, weight) of the S -> Lt resampling matrix and of its transpose (host lists)."""
    rows = []
    if S > Lt:
        for i in range(Lt):
            lo = (i * S) // Lt
            hi = -((-(i + 1) * S) // Lt)
            w = 1.0 / (hi - lo)
            rows.append([(s, w) for s in range(lo, hi)])
    else:
        scale = torch.tensor((S - 1) / (Lt - 1) if Lt > 1 else 0.0, dtype=torch.float32)
        for i in range(Lt):
            src = torch.tensor(float(i), dtype=torch.float32) * scale  # fp32 index math, as ATen does
            lo = int(src)
            hi = min(lo + 1, S - 1)
            lam = float(src - lo)
            rows.append([(lo, 1.0 - lam), (hi, lam)] if hi != lo else [(lo, 1.0)])
    cols = [[] for _ in range(S)]
    for i, taps in enumerate(rows):
        for s, w in taps:
            cols[s].append((i, w))

    def csr(lists):
        ptr, col, wt = [0], [], []
        for taps in lists:
            for c, w in taps:
                col.append(c)
                wt.append(w)
            ptr.append(len(col))
        return ptr, col, wt

    return csr(rows), csr(cols)


_dev_cache = {}


def _device_taps(S, Lt, device):
    key = (S, Lt, str(device))
    if key not in _dev_cache:
        out = []
        for ptr, col, wt in _taps(S, Lt):
            out.append((torch.tensor(ptr, dtype=torch.int32, device=device),
                        torch.tensor(col, dtype=torch.int32, device=device),
                        torch.tensor(wt, dtype=torch.float32, device=device)))
        _dev_cache[key] = out
    return _dev_cache[key]


class _ResampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target_len):
        B, S, H = x.shape
        fwd, _ = _device_taps(S, target_len, x.device)
        x = x.contiguous()
        out = torch.empty(B, target_len, H, dtype=x.dtype, device=x.device)
        L.row_resample(x, out, *fwd)
        ctx.dims = (S, target_len)
        return out

    @staticmethod
    def backward(ctx, dy):
        S, Lt = ctx.dims
        _, bwd = _device_taps(S, Lt, dy.device)
        dy = dy.contiguous()
        dx = torch.empty(dy.shape[0], S, dy.shape[2], dtype=dy.dtype, device=dy.device)
        L.row_resample(dy, dx, *bwd)
        return dx, None


def adaptive_projection(tensor: torch.Tensor, target_length: int) -> torch.Tensor:
    """[B, S, H] -> [B, target_length, H] (clip_whisper_model.py:621-707, training branch)."""
    if tensor.shape[1] == target_length:
        return tensor
    if not tensor.is_cuda:
        raise L.ConnectorError("adaptive_projection: tensor is not on a CUDA device (no CPU fallback)")
    if tensor.dtype not in L.DTYPE_CODES:
        raise L.ConnectorError(f"adaptive_projection: dtype {tensor.dtype} unsupported (fp32, bf16 or fp16)")
    return _ResampleFn.apply(tensor, target_length)


def adapt_mask(mask: torch.Tensor, target_length: int) -> torch.Tensor:
    """Slice, or right-pad with ones (clip_whisper_model.py:709-736).  Index bookkeeping, not arithmetic."""
    if mask.shape[1] >= target_length:
        return mask[:, :target_length]
    pad = torch.ones(mask.shape[0], target_length - mask.shape[1], dtype=mask.dtype, device=mask.device)
    return torch.cat([mask, pad], dim=1)
