"""Drop-in for the reference's src/clip_whisper/models/modality_connector.py on the B200 path.

Same class names, constructor signatures, call convention and state-dict keys (`linear.weight` [H, D],
`linear.bias` [H]) as the reference (modality_connector.py:6-44, 383-402), so checkpoints written by the
reference (`audio_connector.pt` / `video_connector.pt`, clip_whisper_model.py:745-746) load unchanged and
`decode.py:211-260` style re-creation works.  The arithmetic runs in libavconnector_b200.so (bf16 operands,
fp32 accumulate, tcgen05) -- there is no PyTorch or CPU fallback.

Master parameters are kept in fp32 whatever `dtype` says; `dtype` selects the OUTPUT dtype (fp32 | bf16).
The reference's `use_fp16` path (fp16 connector weights) maps to bf16 output here: fp16 does not exist on
this path.
"""
from __future__ import annotations

import logging

import torch
import torch.nn as nn

from . import _lib as L
from .connector_ops import linear_project


def _out_dtype(dtype: torch.dtype) -> torch.dtype:
    if dtype == torch.float16:
        logging.warning("fp16 requested: the B200 connector computes in bf16 and emits bf16 instead")
        return torch.bfloat16
    if dtype not in (torch.float32, torch.bfloat16):
        raise L.ConnectorError(f"connector dtype {dtype} unsupported (fp32 or bf16)")
    return dtype


class BaseModalityConnector(nn.Module):
    """Base class for modality connectors (modality_connector.py:6-23)."""

    def __init__(self, input_dim, output_dim, device="cuda", dtype=torch.float32):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.device = device
        self.dtype = _out_dtype(dtype)

    def forward(self, x):
        # the reference casts x to the module dtype here (:18-19); the kernels take bf16 operands, so the cast
        # to bf16 happens inside the op and self.dtype is the dtype of the result
        return self._forward_impl(x)

    def _forward_impl(self, x):
        raise NotImplementedError("Subclasses must implement _forward_impl")


class SimpleModalityConnector(BaseModalityConnector):
    """Linear projection D -> H (modality_connector.py:25-44): xavier-uniform weight, zero bias."""

    def __init__(self, input_dim, output_dim, device="cuda", dtype=torch.float32, max_seq_len=None, **kwargs):
        super().__init__(input_dim, output_dim, device, dtype)
        if input_dim % 8 or output_dim % 8:
            raise ValueError("input_dim and output_dim must be multiples of 8 (128-bit bf16 vectors)")
        self.linear = nn.Linear(input_dim, output_dim)
        nn.init.xavier_uniform_(self.linear.weight)
        nn.init.zeros_(self.linear.bias)
        self.linear = self.linear.to(device=device, dtype=torch.float32)

    def _forward_impl(self, x):
        return linear_project(x, self.linear.weight, self.linear.bias, self.dtype)


class MLPModalityConnector(BaseModalityConnector):
    """Linear -> GELU(erf) -> Linear (LLaVA `mlp2x_gelu`); new on this path (north_star "linear/MLP projector").
    Parameters: fc1.weight [hidden_dim, input_dim], fc1.bias, fc2.weight [output_dim, hidden_dim], fc2.bias."""

    def __init__(self, input_dim, output_dim, device="cuda", dtype=torch.float32, max_seq_len=None, hidden_dim=None,
                 **kwargs):
        super().__init__(input_dim, output_dim, device, dtype)
        hidden_dim = hidden_dim or output_dim
        if input_dim % 8 or output_dim % 8 or hidden_dim % 8:
            raise ValueError("input_dim, hidden_dim and output_dim must be multiples of 8")
        self.hidden_dim = hidden_dim
        self.fc1 = nn.Linear(input_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, output_dim)
        for lin in (self.fc1, self.fc2):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)
        self.fc1 = self.fc1.to(device=device, dtype=torch.float32)
        self.fc2 = self.fc2.to(device=device, dtype=torch.float32)

    def mlp_params(self):
        return (self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)

    def _forward_impl(self, x):
        from .connector_ops import FusePlan, fused_connector

        squeeze = x.dim() == 2
        if squeeze:
            x = x.unsqueeze(0)
        emb, _, _ = fused_connector(x, None, None, None, None, None, FusePlan(modality="audio"),
                                    out_dtype=self.dtype, mlp_audio=self.mlp_params())
        return emb[0] if squeeze else emb


class UnsupportedConnector(BaseModalityConnector):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "only the 'simple' (linear) projector is on the B200 hot path; in the reference 'deep', 'conv' and "
            "'attention' cannot be constructed through the model either (they reject max_seq_len=, "
            "clip_whisper_model.py:1171-1189) and 'adaptive' is outside this path's scope (SURVEY.md 8(a) A2)")


def create_modality_connector(connector_type, input_dim, output_dim, **kwargs):
    """Factory with the reference's signature (modality_connector.py:383-399)."""
    connector_map = {"simple": SimpleModalityConnector, "mlp": MLPModalityConnector}
    if connector_type not in connector_map:
        raise NotImplementedError(
            f"connector type {connector_type!r} is not available on the B200 path ('simple' or 'mlp'); the reference "
            "would fall back to 'deep' here and then fail with a TypeError (SURVEY.md 8(a) A2)")
    return connector_map[connector_type](input_dim, output_dim, **kwargs)


# For backward compatibility (modality_connector.py:402)
ModalityConnector = SimpleModalityConnector
