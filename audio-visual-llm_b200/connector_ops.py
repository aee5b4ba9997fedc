"""Host-side orchestration of the connector kernels (gather -> projector GEMM -> splice, and the backward).

Everything numeric happens inside libavconnector_b200.so (sm_100a CUDA); torch is used for device memory,
streams and autograd bookkeeping only.  There is no eager-PyTorch or CPU fallback: inputs that are not on an
sm_100 device raise.

Reference semantics replaced (paths relative to /root/reference/src/clip_whisper/models/):
  modality_connector.py:16-20,43-44     connector = cast + nn.Linear
  clip_whisper_model.py:320-374,424-434  pad/truncate AFTER projection + weighted-sum fusion
  clip_whisper_model.py:448-451,460      prompt-embedding concat + int64 ones mask
  clip_whisper_model.py:569-570,586-598  pad -> -100, truncate / right-pad labels
  clip_whisper_model.py:1141-1142        CLS-row select (folded into the gather through the frame stride)
The weighted sum after two projections is computed as ONE GEMM over [a_j ; v_j] with
W = [sa*Wa | sv*Wv] and bias sa*ba*1[audio token] + sv*bv*1[video token].
"""
from __future__ import annotations

import os
import weakref
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L

MAX_PROMPT_LEN = 32  # clip_whisper_model.py:469
# bf16 (training configuration), fp16 (the reference's use_fp16 mode, clip_whisper_model.py:164), fp32 (its default)
_SUPPORTED_OUT = (torch.float32, torch.bfloat16, torch.float16)


def _bias_in_gemm() -> bool:
    """db comes out of the dW GEMM launch (extra work items of its tile schedule) unless AVC_BIAS_IN_GEMM=0 selects
    the stand-alone column-sum kernel."""
    return os.environ.get("AVC_BIAS_IN_GEMM", "1") != "0"


@dataclass(frozen=True)
class FusePlan:
    """Connector knobs.  Defaults = reference behaviour (index-aligned, k=1, sum fusion, ones mask)."""
    modality: str = "both"        # audio | video | both                 (configs/clip_whisper.yaml:21)
    fusion: str = "sum"           # sum (clip_whisper_model.py:434) | concat (new)
    fusion_scale: float = 0.5     # configs/clip_whisper.yaml:30
    max_seq_len: int = 256        # caps fused tokens in `both` mode only  (clip_whisper_model.py:427)
    audio_stride: int = 1         # k_a frames stacked per token (new; 1 = reference)
    video_stride: int = 1         # k_v
    audio_repeat: int = 1         # r_a: every stacked audio token is used r_a times (token j -> stack j // r_a)
    video_repeat: int = 1         # r_v: 2 pairs 25 fps video with 50 Hz audio at stride 1 (rate alignment)
    mask_mode: int = 0            # 0 all ones (clip_whisper_model.py:460) | 1 valid tokens only
    label_mode: int = 0           # 0 reference eval rule | 1 also -100 on placeholders / pad ids

    def __post_init__(self):
        if self.modality not in ("audio", "video", "both"):
            raise ValueError(f"modality must be audio|video|both, got {self.modality!r}")
        if self.fusion not in ("sum", "concat"):
            raise ValueError(f"fusion must be sum|concat, got {self.fusion!r}")
        if self.audio_stride < 1 or self.video_stride < 1:
            raise ValueError("strides must be >= 1")
        if self.audio_repeat < 1 or self.video_repeat < 1:
            raise ValueError("repeats must be >= 1")

    def scales(self, use_a: bool, use_v: bool) -> Tuple[float, float]:
        if use_a and use_v and self.fusion == "sum":
            return float(self.fusion_scale), float(1 - self.fusion_scale)
        return 1.0, 1.0

    def tokens(self, Ta: Optional[int], Tv: Optional[int]) -> int:
        """Fused tokens for Ta audio / Tv video frames: ceil(T/k) per stream, max, capped in `both` mode."""
        na = -(-Ta // self.audio_stride) * self.audio_repeat if Ta is not None else None
        nv = -(-Tv // self.video_stride) * self.video_repeat if Tv is not None else None
        if na is not None and nv is not None:
            return min(self.max_seq_len, max(na, nv))
        return na if na is not None else nv


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise L.ConnectorError(f"{what} is on {t.device}; the B200 connector has no CPU fallback "
                               "(the reference's cuda-else-cpu default is removed on this path)")


def to_bf16_features(x: torch.Tensor) -> torch.Tensor:
    """[B, T, D] features of any float dtype -> bf16 with D contiguous (cast kernel; strided CLS views allowed)."""
    _require_cuda(x, "features")
    if x.dim() != 3:
        raise ValueError(f"features must be [batch, frames, dim], got {tuple(x.shape)}")
    if x.dtype == torch.bfloat16 and x.stride(2) == 1:
        return x
    if x.dtype not in L.DTYPE_CODES or x.stride(2) != 1 or (x.stride(1) * x.element_size()) % 16 or \
            x.data_ptr() % 16:
        x = x.float().contiguous()
    B, T, D = x.shape
    if D % 8:
        raise ValueError(f"feature dim {D} must be a multiple of 8 (128-bit bf16 vectors)")
    out = torch.empty(B, T, D, dtype=torch.bfloat16, device=x.device)
    if B * T == 0:
        return out
    if x.stride(0) == T * x.stride(1):  # uniform row pitch (dense, or a CLS view of [B*F, 1+Np, D])
        src2d = torch.as_strided(x, (B * T, D), (x.stride(1), 1))
        L.cast_bf16(src2d, out.view(B * T, D))
    else:
        for b in range(B):
            L.cast_bf16(x[b], out[b])
    return out


_PACK_CACHE: dict = {}
_PACK_CACHE_MAX = 16


def invalidate_pack_cache() -> None:
    """Drop every cached bf16 weight pack.  The cache is validated by tensor identity + torch's version counter, which
    an update through raw pointers (`ConnectorAdamW.step`, apex / DeepSpeed flat-parameter optimizers, any write via
    `.data` from a custom kernel) does not bump: such optimizers must call this after they change the weights.
    `ConnectorAdamW.step` and `ClipWhisperModel.train()/eval()` do."""
    _PACK_CACHE.clear()


def pack_projector(weights: Sequence[torch.Tensor], scales: Sequence[float], cache: bool = False) -> torch.Tensor:
    """[H, K_s] fp32 master weights -> one bf16 [H, sum K_s] matrix with the fusion scales folded in.

    Forward-only calls (`cache=True`: no gradient is recorded for the weights -- decode / generate,
    clip_whisper_model.py:1301-1340) reuse the packed
    copy while the parameters are unchanged (same storage and torch version counter), so a latency-bound encode is
    GEMM + splice only.  Training re-packs every step: the optimizer has changed the weights."""
    key = None
    if cache:  # the caller is not recording gradients for these weights
        # keyed by tensor identity; a hit also needs the same live objects (weak references: an address can be
        # reused by another tensor) at the same version counter
        key = tuple(id(w) for w in weights) + tuple(float(s) for s in scales)
        hit = _PACK_CACHE.get(key)
        if hit is not None:
            refs, versions, packed = hit
            if all(r() is w for r, w in zip(refs, weights)) and versions == [w._version for w in weights]:
                return packed
            del _PACK_CACHE[key]
    H = weights[0].shape[0]
    K = sum(w.shape[1] for w in weights)
    packed = torch.empty(H, K, dtype=torch.bfloat16, device=weights[0].device)
    col = 0
    for w, s in zip(weights, scales):
        if w.dtype != torch.float32 or w.stride(1) != 1:
            raise ValueError("projector master weights must be fp32 with contiguous columns")
        L.pack_weight(w, packed[:, col:col + w.shape[1]], s)
        col += w.shape[1]
    if key is not None:
        if len(_PACK_CACHE) >= _PACK_CACHE_MAX:
            _PACK_CACHE.pop(next(iter(_PACK_CACHE)))
        _PACK_CACHE[key] = ([weakref.ref(w) for w in weights], [w._version for w in weights], packed)
    return packed


def _as_int32(x, device) -> Optional[torch.Tensor]:
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int32).contiguous()
    return torch.tensor(list(x), dtype=torch.int32, device=device)


def _to_bf16_rows(t: torch.Tensor) -> torch.Tensor:
    """2-D [rows, cols] bf16 / fp16 / fp32 -> bf16 (the dW GEMM's operand type); no copy if it already is."""
    if t.dtype == torch.bfloat16:
        return t
    out = torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
    if t.numel():
        L.cast_bf16(t, out)
    return out


def _grad_like(g: torch.Tensor, ref_dtype: torch.dtype) -> torch.Tensor:
    return g if g.dtype == ref_dtype else g.to(ref_dtype)


class _LinearProjectFn(torch.autograd.Function):
    """y = x . W^T + b on the tcgen05 projector GEMM (one modality; SimpleModalityConnector._forward_impl)."""

    @staticmethod
    def forward(ctx, x, weight, bias, out_dtype, cache_pack):
        xb = to_bf16_features(x if x.dim() == 3 else x.unsqueeze(0))
        B, T, D = xb.shape
        H = weight.shape[0]
        wp = pack_projector([weight], [1.0], cache=cache_pack)
        y = torch.empty(B, T, H, dtype=out_dtype, device=x.device)
        if B * T:
            L.proj_fwd([xb], [wp], y, bias0=bias)
        ctx.save_for_backward(xb, weight)
        ctx.shape = (H, D)
        ctx.squeeze = x.dim() == 2
        ctx.x_dtype = x.dtype
        return y[0] if ctx.squeeze else y

    @staticmethod
    def backward(ctx, dy):
        xb, weight = ctx.saved_tensors
        H, D = ctx.shape
        B, T = xb.shape[0], xb.shape[1]
        if ctx.squeeze:
            dy = dy.unsqueeze(0)
        dev = dy.device
        dy2 = dy.reshape(B * T, H)
        if dy2.stride(1) != 1 or dy2.stride(0) != H:
            dy2 = dy2.contiguous()
        dyb = _to_bf16_rows(dy2).view(B, T, H)
        dw = db = dx = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw = torch.empty(H, D, dtype=torch.float32, device=dev)
            db = torch.empty(H, dtype=torch.float32, device=dev)
            if B * T == 0:
                dw.zero_()
                db.zero_()
            elif _bias_in_gemm():
                L.proj_bwd_dw(dyb, [xb], [dw], [1.0], bias=(L.present_operand(B, T, dev), db, None, 1.0, 1.0))
            else:
                L.proj_bwd_dw(dyb, [xb], [dw], [1.0])
                L.colsum(dyb, db, None, L.colsum_workspace(H, dev))
        if ctx.needs_input_grad[0]:
            # unfrozen tower (freeze_encoders=False, clip_whisper_model.py:1096,1136): dX = dY . W on the same TN GEMM
            gdt = ctx.x_dtype if ctx.x_dtype in (torch.float32, torch.bfloat16) else torch.bfloat16
            dx = torch.empty(B, T, D, dtype=gdt, device=dev)
            if B * T:
                wt = torch.empty(D, H, dtype=torch.bfloat16, device=dev)
                L.pack_weight_t(weight, wt, 1.0)
                L.proj_bwd_dx([dyb], [wt], dx)
            dx = _grad_like(dx[0] if ctx.squeeze else dx, ctx.x_dtype)
        return dx, dw, db, None, None


def linear_project(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    if out_dtype not in _SUPPORTED_OUT:
        raise L.ConnectorError(f"connector output dtype {out_dtype} unsupported on the B200 path (fp32, bf16 or fp16)")
    _require_cuda(x, "connector input")
    recording = torch.is_grad_enabled() and (weight.requires_grad or bias.requires_grad)
    return _LinearProjectFn.apply(x, weight, bias, out_dtype, not recording)


class _FusedConnectorFn(torch.autograd.Function):
    """gather -> projector GEMM -> splice (+ masks); backward: splice-bwd -> dW GEMM (+ db work items) [-> dX]."""

    @staticmethod
    def forward(ctx, wa, ba, wv, bv, audio_in, video_in, st):
        dev = st["device"]
        audio, video = st["audio"], st["video"]  # bf16 copies / views of audio_in, video_in
        plan: FusePlan = st["plan"]
        use_a, use_v = audio is not None, video is not None
        sa, sv = plan.scales(use_a, use_v)
        B, N, M = st["batch"], st["ntok"], st["rows"]
        H = (wa if use_a else wv).shape[0]
        ka, kv = plan.audio_stride, plan.video_stride
        Ka = ka * audio.shape[2] if use_a else 0
        Kv = kv * video.shape[2] if use_v else 0
        if use_a and wa.shape[1] != Ka:
            raise ValueError(f"audio connector expects input_dim {wa.shape[1]} but stacked audio width is {Ka}")
        if use_v and wv.shape[1] != Kv:
            raise ValueError(f"video connector expects input_dim {wv.shape[1]} but stacked video width is {Kv}")
        K = Ka + Kv
        ws = ([wa] if use_a else []) + ([wv] if use_v else [])
        wp = pack_projector(ws, ([sa] if use_a else []) + ([sv] if use_v else []),
                            cache=st["cache_pack"])
        out_dtype = st["out_dtype"]
        ids = st["input_ids"]
        S = ids.shape[1]
        emb = torch.empty(B, S, H, dtype=out_dtype, device=dev)
        # `[prompt | AV]` layout built by fused_connector: the GEMM epilogue writes the projected rows straight into
        # the AV region of inputs_embeds (scatter output) and the splice kernel only adds text rows and masks
        in_place = bool(st["uniform_layout"] and M)
        Y = emb[:, S - N:, :] if in_place else torch.empty(M, H, dtype=out_dtype, device=dev)
        if use_a and use_v:
            b0, b1, s0, s1 = ba, bv, sa, sv
        elif use_a:
            b0, b1, s0, s1 = ba, None, sa, 0.0
        else:  # video only: its "present" flag is bit 1
            b0, b1, s0, s1 = None, bv, 0.0, sv
        direct = (st["tok_offset"] is None and plan.audio_repeat == 1 and plan.video_repeat == 1
                  and _stack_is_free_view(audio, ka, N) and _stack_is_free_view(video, kv, N))
        if direct:
            # Dense streams whose frame counts divide by the stride: stacking k frames is a free reshape
            # [B, T, D] -> [B*T/k, k*D], so the GEMM reads the tower outputs in place through one TMA descriptor
            # per modality (two K segments) and the gathered operand is never materialised.
            xs = ([audio.view(B * N, Ka)] if use_a else []) + ([video.view(B * N, Kv)] if use_v else [])
            wsegs = ([wp[:, :Ka]] if use_a else []) + ([wp[:, Ka:]] if use_v else [])
            A, flags = None, None
            if M:
                L.proj_fwd(xs, wsegs, Y, bias0=b0, bias1=b1, bias_scale0=s0, bias_scale1=s1)
        else:
            # 1. gather (align + stack + concat, zero padding, CLS stride) into the packed A matrix
            A = torch.empty(M, K, dtype=torch.bfloat16, device=dev)
            flags = torch.empty(M, dtype=torch.uint8, device=dev)
            xs = ([A[:, :Ka]] if use_a else []) + ([A[:, Ka:]] if use_v else [])
            if M:
                L.gather_fwd(audio, video, ka, kv, B, N, A, flags, st["tok_offset"], st["audio_valid"],
                             st["video_valid"], plan.audio_repeat, plan.video_repeat)
                # 2. projector: one GEMM over [a ; v]
                L.proj_fwd([A], [wp], Y, bias0=b0, bias1=b1, bias_scale0=s0, bias_scale1=s1, row_flags=flags)
        # 3. splice into the LLM input-embedding sequence + masks
        mask = torch.empty(B, S, dtype=torch.int64, device=dev)
        want_labels = st["labels"] is not None or plan.label_mode == 1
        labels_out = torch.empty(B, S, dtype=torch.int64, device=dev) if want_labels else None
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        sp = L.make_splice(ids, st["placeholder_id"], st["pad_id"], H, tokens_per_sample=N,
                           tok_offset=st["tok_offset"], embed_table=st["embed_table"], attention_mask=mask,
                           mask_mode=plan.mask_mode, label_mode=plan.label_mode, labels_in=st["labels"],
                           labels_out=labels_out, status=status, elem_size=emb.element_size(),
                           av_rows_in_place=in_place)
        L.splice_fwd(sp, None if (in_place or not M) else Y, emb)
        ctx.save_for_backward(*xs, *ws)
        ctx.nx = len(xs)
        ctx.flags = flags
        ctx.sp = sp
        ctx.meta = (use_a, use_v, sa, sv, Ka, Kv, H, M, out_dtype)
        # uniform `[prompt | AV]` layout built by fused_connector itself: backward may read d(inputs_embeds) in place
        ctx.uniform = (B, N, S - N) if (st["uniform_layout"] and out_dtype == torch.bfloat16) else None
        ctx.grad_sync = st.get("grad_sync")
        # what the input gradients (unfrozen towers) need: the gather geometry and the raw inputs' layout
        ctx.dx = dict(direct=direct, B=B, N=N, plan=plan, tok_offset=st["tok_offset"],
                      audio_valid=st["audio_valid"], video_valid=st["video_valid"],
                      a_meta=None if audio_in is None else (tuple(audio_in.shape), audio_in.dtype),
                      v_meta=None if video_in is None else (tuple(video_in.shape), video_in.dtype))
        st["status"] = status
        st["row_flags"] = flags
        st["direct"] = direct
        ctx.mark_non_differentiable(mask)
        if labels_out is not None:
            ctx.mark_non_differentiable(labels_out)
            return emb, mask, labels_out
        return emb, mask

    @staticmethod
    def backward(ctx, d_emb, *unused):
        saved = list(ctx.saved_tensors)
        xs, ws = saved[:ctx.nx], saved[ctx.nx:]
        flags = ctx.flags
        use_a, use_v, sa, sv, Ka, Kv, H, M, out_dtype = ctx.meta
        dev = d_emb.device
        want_w = any(ctx.needs_input_grad[:4])
        want_x = ctx.needs_input_grad[4] or ctx.needs_input_grad[5]
        if d_emb.dtype != out_dtype or not d_emb.is_contiguous():
            d_emb = d_emb.to(out_dtype).contiguous()
        sync = ctx.grad_sync if want_w else None
        if sync is not None:
            dwa, dba, dwv, dbv = sync.views(use_a, use_v)
        else:
            dwa = torch.empty(H, Ka, dtype=torch.float32, device=dev) if (use_a and want_w) else None
            dwv = torch.empty(H, Kv, dtype=torch.float32, device=dev) if (use_v and want_w) else None
            dba = torch.empty(H, dtype=torch.float32, device=dev) if (use_a and want_w) else None
            dbv = torch.empty(H, dtype=torch.float32, device=dev) if (use_v and want_w) else None
        if M == 0:
            if sync is not None:
                raise L.ConnectorError("data-parallel gradient sync needs at least one fused token on every rank")
            for t in (dwa, dwv, dba, dbv):
                if t is not None:
                    t.zero_()
            return dwa, dba, dwv, dbv, None, None, None
        # ---- dY: d(inputs_embeds) read in place (`[prompt | AV]`, bf16) or gathered back into packed rows
        if ctx.uniform is not None:
            B, N, P = ctx.uniform
            dy, base = d_emb, P
            x3 = [x.view(B, N, x.shape[1]) for x in xs]
            present_shape = (B, N)
        else:
            dY = torch.empty(M, H, dtype=out_dtype, device=dev)
            L.splice_bwd(ctx.sp, d_emb, dY)
            dy, base = _to_bf16_rows(dY), 0
            x3 = xs
            present_shape = (1, M)
        inv = 1.0 / sync.world if sync is not None else 1.0
        if want_w:
            dws = ([dwa] if use_a else []) + ([dwv] if use_v else [])
            al = ([sa * inv] if use_a else []) + ([sv * inv] if use_v else [])
            if _bias_in_gemm() or sync is not None:
                bias = (L.present_operand(*present_shape, dev, row_flags=flags), dba, dbv, sa * inv, sv * inv)
                if sync is not None:
                    L.proj_bwd_dw_allreduce(dy, x3, dws, al, sync.next_epoch(), dy_row_base=base, bias=bias)
                else:
                    L.proj_bwd_dw(dy, x3, dws, al, dy_row_base=base, bias=bias)
            else:
                L.proj_bwd_dw(dy, x3, dws, al, dy_row_base=base)
                # flag bit0 = audio token present, bit1 = video token present (all present when flags is None)
                cs = dict(dy_row_base=base, sum_rows=present_shape[1]) if ctx.uniform is not None else {}
                L.colsum(dy, dba, dbv, L.colsum_workspace(H, dev), row_flags=flags, alpha0=sa, alpha1=sv, **cs)
        d_audio = d_video = None
        if want_x:
            d_audio, d_video = _input_grads(ctx, dy, base, ws, dev)
        if sync is not None:
            return None, None, None, None, d_audio, d_video, None  # the gradients live in the attached bucket views
        return dwa, dba, dwv, dbv, d_audio, d_video, None


def _input_grads(ctx, dy, base, ws, dev):
    """d(audio), d(video) for unfrozen towers: dA_s = dY . (scale_s W_s) per modality on the TN GEMM, then the
    transpose of the gather (a free reshape when the stack was one)."""
    use_a, use_v, sa, sv, Ka, Kv, H, M, _ = ctx.meta
    g = ctx.dx
    plan: FusePlan = g["plan"]
    B, N = g["B"], g["N"]
    if ctx.uniform is not None:
        dy_op = dy[:, base:, :]                     # [B, N, H] view of d(inputs_embeds)
    else:
        dy_op = dy                                  # packed [M, H]
    outs = []
    wi = 0
    for use, need, meta, scale, k, rep, valid, col in (
            (use_a, ctx.needs_input_grad[4], g["a_meta"], sa, plan.audio_stride, plan.audio_repeat, g["audio_valid"], 0),
            (use_v, ctx.needs_input_grad[5], g["v_meta"], sv, plan.video_stride, plan.video_repeat, g["video_valid"], Ka)):
        if not use:
            outs.append(None)
            continue
        w = ws[wi]
        wi += 1
        if not need:
            outs.append(None)
            continue
        (Bx, T, D), in_dtype = meta
        Ks = w.shape[1]
        wt = torch.empty(Ks, H, dtype=torch.bfloat16, device=dev)
        L.pack_weight_t(w, wt, scale)
        gdt = in_dtype if in_dtype in (torch.float32, torch.bfloat16) else torch.bfloat16
        if g["direct"]:
            dx = torch.empty(B, N, Ks, dtype=gdt, device=dev)
            L.proj_bwd_dx([dy_op], [wt], dx if ctx.uniform is not None else dx.view(B * N, Ks))
            dx = dx.view(Bx, T, D)
        else:
            dA = torch.empty(M, Ks, dtype=torch.bfloat16, device=dev)
            L.proj_bwd_dx([dy_op], [wt], dA if dy_op.dim() == 2 else dA.view(B, N, Ks))
            dx = torch.empty(Bx, T, D, dtype=gdt, device=dev)
            L.gather_bwd(dA, 0, dx, k, rep, B, N, tok_offset=g["tok_offset"], valid=valid)
        outs.append(_grad_like(dx, in_dtype))
    return outs[0], outs[1]


def _stack_is_free_view(x: Optional[torch.Tensor], k: int, ntok: int) -> bool:
    """True if stacking k frames of [B, T, D] is the reshape [B*T/k, k*D] with exactly `ntok` tokens per sample."""
    if x is None:
        return True
    B, T, D = x.shape
    return (T % k == 0 and T // k == ntok and x.stride(2) == 1 and x.stride(1) == D and x.stride(0) == T * D
            and x.data_ptr() % 16 == 0)


def ragged_token_offsets(plan: FusePlan, batch: int, Ta: Optional[int], Tv: Optional[int], audio_lengths,
                         video_lengths, device, total_tokens: Optional[int] = None):
    """Per-sample fused-token counts -> exclusive prefix offsets [B+1] int32 on the device (+ clamped valid lengths).

    Host lists / CPU tensors: the arithmetic runs on the host and one small H2D copy carries the offsets.  Device
    tensors: the counts and their prefix sum are formed on the device (integer index bookkeeping, no feature data);
    the only host synchronisation left is reading the total row count, which `total_tokens` (e.g. the number of
    placeholders the data pipeline put into input_ids) removes.  Returns (tok_offset, audio_valid, video_valid, rows).
    """
    def on_device(x):
        return isinstance(x, torch.Tensor) and x.is_cuda

    if on_device(audio_lengths) or on_device(video_lengths):
        cnt = None
        av = vv = None
        if Ta is not None and audio_lengths is not None:
            av = torch.as_tensor(audio_lengths, device=device).to(torch.int32).clamp(0, Ta)
        if Tv is not None and video_lengths is not None:
            vv = torch.as_tensor(video_lengths, device=device).to(torch.int32).clamp(0, Tv)
        for T, valid, k, rep in ((Ta, av, plan.audio_stride, plan.audio_repeat),
                                 (Tv, vv, plan.video_stride, plan.video_repeat)):
            if T is None:
                continue
            ln = valid if valid is not None else torch.full((batch,), T, dtype=torch.int32, device=device)
            n = torch.div(ln + (k - 1), k, rounding_mode="floor") * rep
            cnt = n if cnt is None else torch.maximum(cnt, n)
        if Ta is not None and Tv is not None:
            cnt = cnt.clamp(max=plan.max_seq_len)
        offs = torch.zeros(batch + 1, dtype=torch.int32, device=device)
        offs[1:] = torch.cumsum(cnt, 0)
        rows = int(total_tokens) if total_tokens is not None else int(offs[-1].item())
        return offs, av, vv, rows
    la = None if Ta is None or audio_lengths is None else [min(max(int(x), 0), Ta) for x in _host_list(audio_lengths)]
    lv = None if Tv is None or video_lengths is None else [min(max(int(x), 0), Tv) for x in _host_list(video_lengths)]
    offs = [0]
    for b in range(batch):
        offs.append(offs[-1] + plan.tokens(la[b] if la is not None else Ta, lv[b] if lv is not None else Tv))
    if total_tokens is not None and int(total_tokens) != offs[-1]:
        raise ValueError(f"total_tokens={total_tokens} but the lengths give {offs[-1]} fused tokens")
    return (torch.tensor(offs, dtype=torch.int32, device=device), _as_int32(la, device), _as_int32(lv, device),
            offs[-1])


def fused_connector(audio: Optional[torch.Tensor], video: Optional[torch.Tensor], wa, ba, wv, bv, plan: FusePlan, *,
                    input_ids: Optional[torch.Tensor] = None, prompt_ids: Optional[torch.Tensor] = None,
                    placeholder_id: int = -1, embed_table: Optional[torch.Tensor] = None,
                    labels: Optional[torch.Tensor] = None, pad_id: int = 0, out_dtype: torch.dtype = torch.bfloat16,
                    audio_lengths=None, video_lengths=None, total_tokens: Optional[int] = None, check: bool = False,
                    mlp_audio=None, mlp_video=None, grad_sync=None):
    """Tower features -> (inputs_embeds [B, S, H], attention_mask int64 [B, S], labels int64 [B, S] | None).

    audio [B, Ta, Da] / video [B, Tv, Dv] (video may be a strided CLS view of CLIP's last_hidden_state); features that
    require grad (unfrozen towers, freeze_encoders=False) receive input gradients from the backward.
    mlp_audio / mlp_video = (fc1.weight, fc1.bias, fc2.weight, fc2.bias) select the two-layer GELU projector
    instead of the linear one (wa, ba, wv, bv are then ignored).
    Layout: `input_ids` with `placeholder_id` runs marks where the AV tokens go; without it the reference layout
    `[prompt_ids[:, :32] | AV tokens]` is built (clip_whisper_model.py:448-451).
    audio_lengths / video_lengths (host ints or int32 tensors, valid frames per sample) switch to ragged packing:
    sample b contributes ntok_b = tokens(len_a[b], len_v[b]) rows and needs exactly ntok_b placeholders
    (`total_tokens` = their sum, if the caller knows it, saves the host read of the device-side prefix sum).
    grad_sync (`parallel.FusedGradSync`): data parallel -- the backward writes dW / db into the peer-mapped bucket the
    parameters' .grad point at and all-reduces them inside the dW GEMM launch (no NCCL call, no .grad returned).
    """
    if out_dtype not in _SUPPORTED_OUT:
        raise L.ConnectorError(f"LLM dtype {out_dtype} unsupported on the B200 path (fp32, bf16 or fp16)")
    use_a = plan.modality in ("audio", "both") and audio is not None
    use_v = plan.modality in ("video", "both") and video is not None
    if not (use_a or use_v):
        raise ValueError("No valid inputs provided - both audio and video are None")  # clip_whisper_model.py:445
    a = to_bf16_features(audio) if use_a else None
    v = to_bf16_features(video) if use_v else None
    dev = (a if use_a else v).device
    B = (a if use_a else v).shape[0]
    if use_a and use_v and a.shape[0] != v.shape[0]:
        raise ValueError("audio and video batch sizes differ")
    N = plan.tokens(a.shape[1] if use_a else None, v.shape[1] if use_v else None)
    tok_offset = audio_valid = video_valid = None
    rows = B * N
    ragged = audio_lengths is not None or video_lengths is not None
    if ragged:
        tok_offset, audio_valid, video_valid, rows = ragged_token_offsets(
            plan, B, a.shape[1] if use_a else None, v.shape[1] if use_v else None, audio_lengths, video_lengths, dev,
            total_tokens)
    uniform_layout = input_ids is None
    if input_ids is None:
        ph = torch.full((B, N), placeholder_id, dtype=torch.int64, device=dev)
        if ragged:
            raise ValueError("ragged lengths need explicit input_ids with one placeholder per fused token")
        if prompt_ids is not None:
            _require_cuda(prompt_ids, "prompt ids")
            input_ids = torch.cat([prompt_ids[:, :MAX_PROMPT_LEN].to(torch.int64), ph], dim=1)
        else:
            input_ids = ph
    else:
        _require_cuda(input_ids, "input_ids")
        input_ids = input_ids.to(torch.int64).contiguous()
    if embed_table is not None:
        if embed_table.dtype != out_dtype or not embed_table.is_contiguous():
            raise ValueError(f"embedding table must be contiguous {out_dtype} (the LLM dtype)")
    if labels is not None:
        _require_cuda(labels, "labels")
        labels = labels.to(torch.int64).contiguous()
    st = dict(device=dev, audio=a, video=v, plan=plan, batch=B, ntok=N, rows=rows, tok_offset=tok_offset,
              audio_valid=audio_valid, video_valid=video_valid, out_dtype=out_dtype, input_ids=input_ids,
              placeholder_id=placeholder_id, pad_id=pad_id, embed_table=embed_table, labels=labels,
              uniform_layout=uniform_layout, grad_sync=grad_sync)
    trainable = [p for p in (wa, ba, wv, bv, *(mlp_audio or ()), *(mlp_video or ())) if isinstance(p, torch.Tensor)]
    st["cache_pack"] = not (torch.is_grad_enabled() and any(p.requires_grad for p in trainable))
    dummy = torch.zeros(0, device=dev)
    x_grad = torch.is_grad_enabled() and ((use_a and audio.requires_grad) or (use_v and video.requires_grad))
    if mlp_audio is not None or mlp_video is not None:
        # Linear -> GELU -> Linear per modality: (fc1.weight, fc1.bias, fc2.weight, fc2.bias)
        from .mlp_ops import FusedMLPConnectorFn

        if x_grad:
            raise NotImplementedError("the MLP projector does not produce input gradients (unfrozen towers); use the "
                                      "linear projector (connector_type='simple') with freeze_encoders=False")
        if grad_sync is not None:
            raise NotImplementedError("grad_sync covers the linear projector's gradient bucket only")
        pa = tuple(mlp_audio) if use_a else (dummy,) * 4
        pv = tuple(mlp_video) if use_v else (dummy,) * 4
        out = FusedMLPConnectorFn.apply(*pa, *pv, st)
    else:
        out = _FusedConnectorFn.apply(wa if use_a else dummy, ba if use_a else dummy, wv if use_v else dummy,
                                      bv if use_v else dummy, audio if use_a else None, video if use_v else None, st)
    if check and int(st["status"].item()) != 0:
        raise L.ConnectorError("placeholder count does not match the number of fused tokens for some sample")
    emb, mask = out[0], out[1]
    return emb, mask, (out[2] if len(out) > 2 else None)


def _host_list(x):
    if isinstance(x, torch.Tensor):
        return x.tolist()
    return list(x)
