"""Drop-in for the reference's src/clip_whisper/models/modality_connector.py on the B200 path.

Same class names, constructor signatures, call convention and state-dict keys (`linear.weight` [H, D],
`linear.bias` [H]) as the reference (modality_connector.py:6-44, 383-402), so checkpoints written by the
reference (`audio_connector.pt` / `video_connector.pt`, clip_whisper_model.py:745-746) load unchanged and
`decode.py:211-260` style re-creation works.  The arithmetic runs in libavconnector_b200.so (bf16 operands,
fp32 accumulate, tcgen05) -- there is no PyTorch or CPU fallback.

Master parameters are kept in fp32 whatever `dtype` says; `dtype` selects the OUTPUT dtype (fp32 | bf16 | fp16).
The reference's `use_fp16` path (fp16 connector weights, clip_whisper_model.py:164, modality_connector.py:18-19) emits
fp16 here as well: bf16 tensor-core operands, fp32 accumulate, fp16 pack in the GEMM epilogue.
"""
from __future__ import annotations


import torch
import torch.nn as nn

from . import _lib as L
from .connector_ops import linear_project


def _out_dtype(dtype: torch.dtype) -> torch.dtype:
    if dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise L.ConnectorError(f"connector dtype {dtype} unsupported (fp32, bf16 or fp16)")
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


class _ProjLinear(nn.Linear):
    """nn.Linear whose matmul runs on the tcgen05 projector GEMM (fwd, dW + db, dX); same parameters / state-dict keys."""

    def __init__(self, in_features, out_features, out_dtype=torch.float32):
        super().__init__(in_features, out_features)
        self.out_dtype = out_dtype

    def forward(self, x):
        return linear_project(x, self.weight, self.bias, self.out_dtype)


class _PositionalEncoding(nn.Module):
    """Sinusoidal table added to the sequence (modality_connector.py:301-323)."""

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        import math

        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)

    def forward(self, x):
        return x + self.pe[:x.size(1)].unsqueeze(0)


class _AdaptiveSequencePooling(nn.Module):
    """modality_connector.py:325-380: > 512 frames go through two stride-2 convolutions, then self-attention + norm."""

    def __init__(self, dim):
        super().__init__()
        self.long_adapter = nn.Sequential(nn.Conv1d(dim, dim, kernel_size=3, stride=2, padding=1), nn.GELU(),
                                          nn.Conv1d(dim, dim, kernel_size=3, stride=2, padding=1))
        self.attn = nn.MultiheadAttention(embed_dim=dim, num_heads=8, dropout=0.1, batch_first=True)
        self.norm = nn.LayerNorm(dim)
        for m in self.modules():
            if isinstance(m, (nn.Conv1d, nn.Linear)):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        if x.shape[1] > 512:
            x = self.long_adapter(x.transpose(1, 2)).transpose(1, 2)
        residual = x
        x, _ = self.attn(x, x, x)
        return self.norm(x + residual)


class AdaptiveModalityConnector(BaseModalityConnector):
    """The reference's `adaptive` connector (modality_connector.py:239-299) -- the only type besides `simple` that the
    reference model can construct.  Same sub-modules, parameter names and state-dict keys
    (input_proj, norm1, pos_encoder.pe, adaptive_pool.{long_adapter,attn,norm}, output_proj, norm2).  Its two dense
    projections -- where its FLOPs are -- run on the tcgen05 projector GEMM (forward, dW + db, and dX, which carries
    the gradient back through output_proj into the layers before it); LayerNorm / GELU / Conv1d / self-attention are
    not part of the hot path BASELINE.json names and stay the PyTorch modules the reference uses.  It is a stand-alone
    module (`connector(x)`); the fused gather -> GEMM -> splice path requires `simple` or `mlp`."""

    def __init__(self, input_dim, output_dim, device="cuda", dtype=torch.float32, max_seq_len=1536):
        super().__init__(input_dim, output_dim, device, dtype)
        self.max_seq_len = max_seq_len
        mid_dim = (input_dim + output_dim) // 2
        if input_dim % 8 or output_dim % 8 or mid_dim % 8:
            raise ValueError("input_dim, output_dim and their mean must be multiples of 8 (128-bit bf16 vectors)")
        self.input_proj = _ProjLinear(input_dim, mid_dim, torch.float32)
        self.norm1 = nn.LayerNorm(mid_dim)
        self.act = nn.GELU()
        self.pos_encoder = _PositionalEncoding(mid_dim, max_len=max_seq_len)
        self.adaptive_pool = _AdaptiveSequencePooling(mid_dim)
        self.output_proj = _ProjLinear(mid_dim, output_dim, torch.float32)
        self.norm2 = nn.LayerNorm(output_dim)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        self.to(device=device, dtype=torch.float32)  # fp32 masters, as every projector on this path

    def _forward_impl(self, x):
        x = self.act(self.norm1(self.input_proj(x)))
        x = self.adaptive_pool(self.pos_encoder(x))
        x = self.norm2(self.output_proj(x))
        return x if x.dtype == self.dtype else x.to(self.dtype)


def create_modality_connector(connector_type, input_dim, output_dim, **kwargs):
    """Factory with the reference's signature (modality_connector.py:383-399).

    `simple` and `adaptive` are the two types the reference model can actually build (its `deep` / `conv` / `attention`
    classes reject the `max_seq_len=` keyword the model always passes, clip_whisper_model.py:1171-1189, and unknown
    names fall back to `deep` and then raise the same TypeError); `mlp` is the north_star's GELU projector.  Types the
    reference cannot construct are refused here with the reason instead of the reference's TypeError."""
    connector_map = {"simple": SimpleModalityConnector, "mlp": MLPModalityConnector,
                     "adaptive": AdaptiveModalityConnector}
    if connector_type not in connector_map:
        raise NotImplementedError(
            f"connector type {connector_type!r} is not available on the B200 path ('simple', 'mlp' or 'adaptive'); the "
            "reference cannot construct it through the model either: it falls back to 'deep' / rejects max_seq_len= "
            "and fails with a TypeError (SURVEY.md 8(a) A2)")
    return connector_map[connector_type](input_dim, output_dim, **kwargs)


# For backward compatibility (modality_connector.py:402)
ModalityConnector = SimpleModalityConnector
