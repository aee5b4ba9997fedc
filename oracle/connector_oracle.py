"""TEST INFRASTRUCTURE ONLY -- CPU (torch fp32/fp64) restatement of the reference connector path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
file, and only as the checker / the timed CPU baseline.  The product (audio-visual-llm_b200/) never does.

Parity pin: the reference has no tests or golden vectors for this path, so this
restatement is pinned against outputs of the reference itself, executed in the build container by
oracle/make_golden.py (fixtures in tests/golden/*.npz, checked by tests/test_oracle_golden.py).

Two tiers:
  * reference_* functions follow the reference line by line (citations are relative to
    /root/reference/src/clip_whisper/models/);
  * connector_forward() is the fp32 restatement of the B200 path (gather -> one GEMM over [a ; v] ->
    masked bias -> splice), including the north_star extensions (stride-k stacking, concat fusion, GELU-MLP,
    placeholder splice, real masks).  tests/test_oracle_golden.py shows it reduces to the reference tier
    when every knob is at its parity setting.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn.functional as F

MAX_PROMPT_LEN = 32  # clip_whisper_model.py:469


# ------------------------------------------------------------------------------------- reference tier
def reference_connector(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """SimpleModalityConnector.forward: cast to module dtype, then nn.Linear (modality_connector.py:16-20, 43-44)."""
    return F.linear(x.to(weight.dtype), weight, bias)


def reference_pad_or_truncate(feats: torch.Tensor, target_len: int) -> torch.Tensor:
    """_pad_or_truncate on [B, t, H] (clip_whisper_model.py:349-366): slice or right-pad with zeros."""
    t = feats.shape[1]
    if t == target_len:
        return feats
    if t > target_len:
        return feats[:, :target_len, :]
    pad = torch.zeros(feats.shape[0], target_len - t, feats.shape[2], dtype=feats.dtype)
    return torch.cat([feats, pad], dim=1)


def reference_cls_select(clip_hidden: torch.Tensor, batch: int, frames: int) -> torch.Tensor:
    """last_hidden_state[:, 0].view(B, F, -1) (clip_whisper_model.py:1141-1142)."""
    return clip_hidden[:, 0].reshape(batch, frames, -1)


def reference_encode(audio_feats, video_feats, wa, ba, wv, bv, *, modality="both", fusion_scale=0.5,
                     max_seq_len=256, prompt_ids=None, embed_table=None, out_dtype=None):
    """ClipWhisperModel.encode (clip_whisper_model.py:407-462) on tower outputs.

    Returns (inputs_embeds [B, P+T, H], attention_mask int64 ones [B, P+T])."""
    a = v = None
    if modality in ("audio", "both") and audio_feats is not None:
        a = reference_connector(audio_feats, wa, ba)  # :1105
    if modality in ("video", "both") and video_feats is not None:
        v = reference_connector(video_feats, wv, bv)  # :1145
    if a is not None and v is not None:
        max_len = min(max_seq_len, max(a.shape[1], v.shape[1]))  # :426-427
        a = reference_pad_or_truncate(a, max_len)  # :430  (pad AFTER projection: padded rows carry no bias)
        v = reference_pad_or_truncate(v, max_len)  # :431
        out = fusion_scale * a + (1 - fusion_scale) * v  # :434
    elif a is not None:
        out = a  # :436-439 (no max_seq_len cap)
    elif v is not None:
        out = v  # :440-443
    else:
        raise ValueError("No valid inputs provided - both audio and video are None")  # :445
    if prompt_ids is not None:
        ids = prompt_ids[:, :MAX_PROMPT_LEN]  # :481-482
        out = torch.cat([F.embedding(ids, embed_table), out], dim=1)  # :484-485, :450
    if out_dtype is not None and out.dtype != out_dtype:
        out = out.to(out_dtype)  # :454-457
    mask = torch.ones(out.shape[0], out.shape[1], dtype=torch.long)  # :460
    return out, mask


def reference_labels_eval(labels: torch.Tensor, pad_token_id: int, seq_len: int) -> torch.Tensor:
    """forward(), eval branch (clip_whisper_model.py:569-570, 586-598): pad -> -100, truncate / right-pad -100."""
    labels = labels.clone()
    labels[labels == pad_token_id] = -100
    if labels.shape[1] > seq_len:
        return labels[:, :seq_len]
    if labels.shape[1] < seq_len:
        pad = torch.full((labels.shape[0], seq_len - labels.shape[1]), -100, dtype=labels.dtype)
        return torch.cat([labels, pad], dim=1)
    return labels


def reference_adaptive_projection(x: torch.Tensor, target_len: int) -> torch.Tensor:
    """_adaptive_projection, training branch (clip_whisper_model.py:633-676): adaptive avg-pool when shrinking,
    linear interpolation with align_corners=True when growing, written out as explicit index arithmetic."""
    B, S, H = x.shape
    if S == target_len:
        return x
    if S > target_len:
        # AdaptiveAvgPool1d: window i = [floor(i*S/L), ceil((i+1)*S/L))
        rows = []
        for i in range(target_len):
            lo = (i * S) // target_len
            hi = -((-(i + 1) * S) // target_len)
            rows.append(x[:, lo:hi].mean(dim=1))
        return torch.stack(rows, dim=1)
    # F.interpolate(mode="linear", align_corners=True): src = i * (S-1)/(L-1)
    out = torch.empty(B, target_len, H, dtype=x.dtype)
    scale = (S - 1) / (target_len - 1) if target_len > 1 else 0.0
    for i in range(target_len):
        src = torch.tensor(i, dtype=torch.float32) * torch.tensor(scale, dtype=torch.float32)
        lo = int(src)
        hi = min(lo + 1, S - 1)
        lam = src - lo
        out[:, i] = x[:, lo] * (1 - lam) + x[:, hi] * lam
    return out


def reference_adapt_mask(mask: torch.Tensor, target_len: int) -> torch.Tensor:
    """_adapt_mask (clip_whisper_model.py:709-736)."""
    if mask.shape[1] >= target_len:
        return mask[:, :target_len]
    return torch.cat([mask, torch.ones(mask.shape[0], target_len - mask.shape[1], dtype=mask.dtype)], dim=1)


def reference_forward_inputs(audio_feats, video_feats, wa, ba, wv, bv, *, labels=None, pad_token_id=0,
                             training=False, **kw):
    """What forward() hands to the LLM (clip_whisper_model.py:489-607): (inputs_embeds, attention_mask, labels)."""
    emb, mask = reference_encode(audio_feats, video_feats, wa, ba, wv, bv, **kw)
    if labels is None:
        return emb, mask, None
    lab = labels.clone()
    lab[lab == pad_token_id] = -100  # :569-570
    if lab.shape[1] != emb.shape[1]:
        if training:  # :577-585
            emb = reference_adaptive_projection(emb, lab.shape[1])
            mask = reference_adapt_mask(mask, lab.shape[1])
        else:  # :586-598
            lab = reference_labels_eval(labels, pad_token_id, emb.shape[1])
    return emb, mask, lab


# ------------------------------------------------------------------------------ B200-path restatement
@dataclass
class ConnectorSpec:
    """Knobs of the fused connector.  Defaults are the reference-parity settings."""
    modality: str = "both"          # audio | video | both            (configs/clip_whisper.yaml:21)
    fusion: str = "sum"             # sum (reference :434) | concat   (new)
    fusion_scale: float = 0.5       # (configs/clip_whisper.yaml:30)
    max_seq_len: int = 256          # cap on fused tokens in `both` mode (clip_whisper_model.py:427)
    audio_stride: int = 1           # k_a frames stacked per token     (new; 1 = reference)
    video_stride: int = 1           # k_v
    audio_repeat: int = 1           # each stacked token used r times (token j -> stack j // r)   (new; 1 = reference)
    video_repeat: int = 1
    mask_mode: int = 0              # 0 all ones (reference :460) | 1 valid tokens only
    label_mode: int = 0             # 0 reference eval rule | 1 also -100 on placeholders / pads
    act: int = 0                    # 0 linear projector | 1 erf-GELU after the (first) projection
    extra: dict = field(default_factory=dict)


def stack_frames(x: torch.Tensor, k: int, ntok: int, valid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, T, D] -> [B, ntok, k*D]: token j = frames k*j .. k*j+k-1, zero at / past the valid length."""
    B, T, D = x.shape
    if valid is not None:
        keep = torch.arange(T).unsqueeze(0) < valid.clamp(0, T).unsqueeze(1)
        x = x * keep.unsqueeze(-1).to(x.dtype)
    need = ntok * k
    if need > T:
        x = torch.cat([x, torch.zeros(B, need - T, D, dtype=x.dtype)], dim=1)
    return x[:, :need].reshape(B, ntok, k * D)


def token_counts(spec: ConnectorSpec, Ta: Optional[int], Tv: Optional[int]) -> int:
    """Fused tokens per sample: ceil(T/k) per stream, max over streams, capped only in `both` mode."""
    na = -(-Ta // spec.audio_stride) * spec.audio_repeat if Ta is not None else None
    nv = -(-Tv // spec.video_stride) * spec.video_repeat if Tv is not None else None
    if na is not None and nv is not None:
        return min(spec.max_seq_len, max(na, nv))
    return na if na is not None else nv


def connector_tokens(audio_feats, video_feats, wa, ba, wv, bv, spec: ConnectorSpec, audio_valid=None,
                     video_valid=None):
    """Projected AV tokens [B, N, H] (fp32/fp64 math on whatever dtype comes in) + row flags [B, N] (bit0 audio,
    bit1 video token present).  out[j] = [a_j ; v_j] . [sa*Wa | sv*Wv]^T + sa*ba*f0 + sv*bv*f1  (SURVEY A7)."""
    use_a = spec.modality in ("audio", "both") and audio_feats is not None
    use_v = spec.modality in ("video", "both") and video_feats is not None
    if not (use_a or use_v):
        raise ValueError("No valid inputs provided - both audio and video are None")
    N = token_counts(spec, audio_feats.shape[1] if use_a else None, video_feats.shape[1] if use_v else None)
    if use_a and use_v and spec.fusion == "sum":
        sa, sv = spec.fusion_scale, 1 - spec.fusion_scale
    else:
        sa = sv = 1.0
    B = (audio_feats if use_a else video_feats).shape[0]
    out = None
    flags = torch.zeros(B, N, dtype=torch.uint8)
    j = torch.arange(N).unsqueeze(0)
    for bit, (use, x, w, b, k, r, valid, s) in enumerate([
            (use_a, audio_feats, wa, ba, spec.audio_stride, spec.audio_repeat, audio_valid, sa),
            (use_v, video_feats, wv, bv, spec.video_stride, spec.video_repeat, video_valid, sv)]):
        if not use:
            continue
        T = x.shape[1]
        n_valid = torch.full((B,), T) if valid is None else valid.clamp(0, T).to(torch.long)
        present = ((j // r) * k) < n_valid.unsqueeze(1)  # token has at least one real frame
        stacked = stack_frames(x.to(w.dtype), k, -(-N // r), valid).repeat_interleave(r, dim=1)[:, :N]
        y = s * (stacked @ w.t() + b * present.unsqueeze(-1).to(w.dtype))
        out = y if out is None else out + y
        flags |= present.to(torch.uint8) << bit
    if spec.act == 1:
        out = F.gelu(out)  # erf form, as nn.GELU() (modality_connector.py:60)
    return out, flags


def splice_tokens(tokens: torch.Tensor, input_ids: torch.Tensor, placeholder_id: int, embed_table, pad_id: int,
                  spec: ConnectorSpec, ntok: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None):
    """inputs_embeds[b, p] = tokens[b, rank(p)] at placeholders, embed_table[id] elsewhere; masks as ConnectorSpec.

    With input_ids = [prompt ids (<= 32) | PLACEHOLDER x T] this is the reference's torch.cat layout
    (clip_whisper_model.py:448-451)."""
    B, S = input_ids.shape
    H = tokens.shape[-1]
    emb = torch.zeros(B, S, H, dtype=tokens.dtype)
    mask = torch.ones(B, S, dtype=torch.long)
    lab = torch.full((B, S), -100, dtype=torch.long)
    for b in range(B):
        n = tokens.shape[1] if ntok is None else int(ntok[b])
        is_ph = input_ids[b] == placeholder_id
        rank = torch.cumsum(is_ph.to(torch.long), 0) - 1
        has = is_ph & (rank < n)
        emb[b, has] = tokens[b, rank[has]]
        text = ~is_ph
        if embed_table is not None:
            ok = text & (input_ids[b] >= 0) & (input_ids[b] < embed_table.shape[0])
            emb[b, ok] = embed_table[input_ids[b, ok]].to(tokens.dtype)
        if spec.mask_mode == 1:
            mask[b] = torch.where(is_ph, has, input_ids[b] != pad_id).to(torch.long)
        if labels is not None:
            L = min(labels.shape[1], S)
            lab[b, :L] = labels[b, :L]
        elif spec.label_mode == 1:
            lab[b] = input_ids[b]
        lab[b][lab[b] == pad_id] = -100
        if spec.label_mode == 1:
            lab[b][is_ph | (input_ids[b] == pad_id)] = -100
    return emb, mask, lab


def connector_forward(audio_feats, video_feats, wa, ba, wv, bv, spec: ConnectorSpec, *, prompt_ids=None,
                      embed_table=None, labels=None, pad_id=0, placeholder_id=-1):
    """Reference layout `[prompt | AV]` through the placeholder splice: the B200 path's semantics in fp32."""
    tokens, flags = connector_tokens(audio_feats, video_feats, wa, ba, wv, bv, spec)
    B, N, _ = tokens.shape
    ph = torch.full((B, N), placeholder_id, dtype=torch.long)
    ids = ph if prompt_ids is None else torch.cat([prompt_ids[:, :MAX_PROMPT_LEN], ph], dim=1)
    emb, mask, lab = splice_tokens(tokens, ids, placeholder_id, embed_table, pad_id, spec, labels=labels)
    return emb, mask, (lab if labels is not None or spec.label_mode == 1 else None), flags


def connector_grads(audio_feats, video_feats, wa, ba, wv, bv, spec: ConnectorSpec, upstream_tokens: torch.Tensor):
    """dW / db of both projectors for d(loss)/d(tokens) = upstream_tokens [B, N, H] (autograd of connector_tokens)."""
    params = [p.detach().clone().requires_grad_(True) for p in (wa, ba, wv, bv)]
    tokens, _ = connector_tokens(audio_feats, video_feats, *params, spec)
    (tokens * upstream_tokens.to(tokens.dtype)).sum().backward()
    return [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]


def connector_tokens_mlp(audio_feats, video_feats, mlp_a, mlp_v, spec: ConnectorSpec, audio_valid=None,
                         video_valid=None):
    """Two-layer GELU projector per modality (new; LLaVA mlp2x_gelu with the erf GELU of modality_connector.py:60):
    tokens = sum_s scale_s * present_s * (gelu(x_s W1_s^T + b1_s) W2_s^T + b2_s), stacking / presence as
    connector_tokens().  mlp_* = (fc1.weight, fc1.bias, fc2.weight, fc2.bias)."""
    use_a = spec.modality in ("audio", "both") and audio_feats is not None
    use_v = spec.modality in ("video", "both") and video_feats is not None
    N = token_counts(spec, audio_feats.shape[1] if use_a else None, video_feats.shape[1] if use_v else None)
    sa, sv = (spec.fusion_scale, 1 - spec.fusion_scale) if (use_a and use_v and spec.fusion == "sum") else (1.0, 1.0)
    B = (audio_feats if use_a else video_feats).shape[0]
    j = torch.arange(N).unsqueeze(0)
    out = None
    for use, x, p, k, valid, s in [(use_a, audio_feats, mlp_a, spec.audio_stride, audio_valid, sa),
                                   (use_v, video_feats, mlp_v, spec.video_stride, video_valid, sv)]:
        if not use:
            continue
        w1, b1, w2, b2 = p
        T = x.shape[1]
        n_valid = torch.full((B,), T) if valid is None else valid.clamp(0, T).to(torch.long)
        present = ((j * k) < n_valid.unsqueeze(1)).unsqueeze(-1).to(w1.dtype)
        h = F.gelu(stack_frames(x.to(w1.dtype), k, N, valid) @ w1.t() + b1)
        y = s * present * (h @ w2.t() + b2)
        out = y if out is None else out + y
    return out
