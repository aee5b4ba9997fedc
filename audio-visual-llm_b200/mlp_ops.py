"""Two-layer GELU projector (Linear -> erf-GELU -> Linear, LLaVA `mlp2x_gelu` style) on the connector kernels.

north_star names a "linear/MLP projector ... with the bias and GELU in the epilogue"; the reference's only
constructible projector is the single nn.Linear, so this is an extension: same gather,
same splice, same fusion algebra, with
    H_s = mask_s * gelu(x_s . W1_s^T + b1_s)                      per modality s (mask: token present)
    Y   = [H_a ; H_v] . [sa W2_a | sv W2_v]^T + sa b2_a mask_a + sv b2_v mask_v
Training keeps the pre-activation Z (needed by gelu') and runs GELU as a separate bf16 pass; forward-only
inference uses the GELU epilogue of the GEMM (act = 1) when every token is present.
Backward: dW2 = dY^T H, db2; dH = dY . W2 (TN GEMM on a transposed bf16 pack of W2); dZ = dH * gelu'(Z);
dW1 = dZ^T x, db1.  No gradient flows into the tower features (frozen encoders).
"""
from __future__ import annotations

import torch

from . import _lib as L
from .connector_ops import FusePlan, _stack_is_free_view, pack_projector


class FusedMLPConnectorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w1a, b1a, w2a, b2a, w1v, b1v, w2v, b2v, st):
        dev = st["device"]
        audio, video = st["audio"], st["video"]
        plan: FusePlan = st["plan"]
        use_a, use_v = audio is not None, video is not None
        sa, sv = plan.scales(use_a, use_v)
        B, N, M = st["batch"], st["ntok"], st["rows"]
        ka, kv = plan.audio_stride, plan.video_stride
        Ka = ka * audio.shape[2] if use_a else 0
        Kv = kv * video.shape[2] if use_v else 0
        mods = []  # (x_s, w1, b1, w2, b2, alpha, flag_bit)
        H = (w2a if use_a else w2v).shape[0]
        out_dtype = st["out_dtype"]
        bf = torch.bfloat16
        direct = (st["tok_offset"] is None and plan.audio_repeat == 1 and plan.video_repeat == 1
                  and _stack_is_free_view(audio, ka, N) and _stack_is_free_view(video, kv, N))
        flags = None
        if direct:
            xa = audio.view(M, Ka) if use_a else None
            xv = video.view(M, Kv) if use_v else None
        else:
            A = torch.empty(M, Ka + Kv, dtype=bf, device=dev)
            flags = torch.empty(M, dtype=torch.uint8, device=dev)
            if M:
                L.gather_fwd(audio, video, ka, kv, B, N, A, flags, st["tok_offset"], st["audio_valid"],
                             st["video_valid"], plan.audio_repeat, plan.video_repeat)
            xa = A[:, :Ka] if use_a else None
            xv = A[:, Ka:] if use_v else None
        if use_a:
            mods.append((xa, w1a, b1a, w2a, b2a, sa, 1))
        if use_v:
            mods.append((xv, w1v, b1v, w2v, b2v, sv, 2))
        for x, w1, _, w2, _, _, _ in mods:
            if w1.shape[1] != x.shape[1] or w2.shape[1] != w1.shape[0] or w2.shape[0] != H:
                raise ValueError("MLP projector shapes do not chain: fc1 [Hd, K], fc2 [H, Hd]")
        need_grad = not st["cache_pack"]  # gradients are being recorded for the projector parameters
        # ---- layer 1 (+ GELU) per modality
        Zs, Hs = [], []
        for x, w1, b1, _, _, _, bit in mods:
            Hd = w1.shape[0]
            w1p = pack_projector([w1], [1.0], cache=not need_grad)
            h = torch.empty(M, Hd, dtype=bf, device=dev)
            if not need_grad and flags is None:
                if M:
                    L.proj_fwd([x], [w1p], h, bias0=b1, act=1)  # GELU in the GEMM epilogue
                z = None
            else:
                z = torch.empty(M, Hd, dtype=bf, device=dev)
                if M:
                    L.proj_fwd([x], [w1p], z, bias0=b1)
                    L.gelu_fwd(z, h, flags, bit)
            Zs.append(z)
            Hs.append(h)
        # ---- layer 2: one GEMM over [H_a ; H_v], epilogue scatters into inputs_embeds when the layout is uniform
        ids = st["input_ids"]
        S = ids.shape[1]
        emb = torch.empty(B, S, H, dtype=out_dtype, device=dev)
        in_place = bool(st["uniform_layout"] and M)
        Y = emb[:, S - N:, :] if in_place else torch.empty(M, H, dtype=out_dtype, device=dev)
        w2p = pack_projector([m[3] for m in mods], [m[5] for m in mods], cache=not need_grad)
        wsegs, col = [], 0
        for m in mods:
            wsegs.append(w2p[:, col:col + m[3].shape[1]])
            col += m[3].shape[1]
        if use_a and use_v:
            bb0, bb1, s0, s1 = b2a, b2v, sa, sv
        elif use_a:
            bb0, bb1, s0, s1 = b2a, None, sa, 0.0
        else:
            bb0, bb1, s0, s1 = None, b2v, 0.0, sv
        if M:
            L.proj_fwd(Hs, wsegs, Y, bias0=bb0, bias1=bb1, bias_scale0=s0, bias_scale1=s1, row_flags=flags)
        mask = torch.empty(B, S, dtype=torch.int64, device=dev)
        want_labels = st["labels"] is not None or plan.label_mode == 1
        labels_out = torch.empty(B, S, dtype=torch.int64, device=dev) if want_labels else None
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        sp = L.make_splice(ids, st["placeholder_id"], st["pad_id"], H, tokens_per_sample=N,
                           tok_offset=st["tok_offset"], embed_table=st["embed_table"], attention_mask=mask,
                           mask_mode=plan.mask_mode, label_mode=plan.label_mode, labels_in=st["labels"],
                           labels_out=labels_out, status=status, elem_size=4 if out_dtype == torch.float32 else 2,
                           av_rows_in_place=in_place)
        L.splice_fwd(sp, None if (in_place or not M) else Y, emb)
        if need_grad:
            ctx.save_for_backward(*[m[0] for m in mods], *Zs, *Hs, *[m[3] for m in mods])
        ctx.nmods = len(mods)
        ctx.flags, ctx.sp = flags, sp
        ctx.meta = (use_a, use_v, [m[5] for m in mods], [m[6] for m in mods], H, M, out_dtype)
        st["status"] = status
        st["row_flags"] = flags
        ctx.mark_non_differentiable(mask)
        if labels_out is not None:
            ctx.mark_non_differentiable(labels_out)
            return emb, mask, labels_out
        return emb, mask

    @staticmethod
    def backward(ctx, d_emb, *unused):
        n = ctx.nmods
        saved = ctx.saved_tensors
        xs, Zs, Hs, w2s = saved[:n], saved[n:2 * n], saved[2 * n:3 * n], saved[3 * n:4 * n]
        use_a, use_v, alphas, bits, H, M, out_dtype = ctx.meta
        flags = ctx.flags
        dev = d_emb.device
        bf = torch.bfloat16
        grads = {}
        if d_emb.dtype != out_dtype or not d_emb.is_contiguous():
            d_emb = d_emb.to(out_dtype).contiguous()
        dY = torch.empty(M, H, dtype=out_dtype, device=dev)
        if M:
            L.splice_bwd(ctx.sp, d_emb, dY)
        if out_dtype == torch.float32:
            dYb = torch.empty(M, H, dtype=bf, device=dev)
            if M:
                L.pack_weight(dY, dYb, 1.0)
            dY = dYb
        ws = L.colsum_workspace(max(H, max(z.shape[1] for z in Zs)), dev)
        # ---- layer 2 gradients
        dw2 = [torch.empty(H, h.shape[1], dtype=torch.float32, device=dev) for h in Hs]
        db2 = [torch.empty(H, dtype=torch.float32, device=dev) for _ in Hs]
        if M:
            L.proj_bwd_dw(dY, list(Hs), dw2, alphas)
            out0 = db2[bits.index(1)] if 1 in bits else None
            out1 = db2[bits.index(2)] if 2 in bits else None
            a0 = alphas[bits.index(1)] if 1 in bits else 1.0
            a1 = alphas[bits.index(2)] if 2 in bits else 1.0
            L.colsum(dY, out0, out1, ws, row_flags=flags, alpha0=a0, alpha1=a1)
        else:
            for t in dw2 + db2:
                t.zero_()
        # ---- through GELU into layer 1
        dw1, db1 = [], []
        for x, z, w2, alpha, bit in zip(xs, Zs, w2s, alphas, bits):
            Hd = z.shape[1]
            g1 = torch.empty(Hd, x.shape[1], dtype=torch.float32, device=dev)
            gb = torch.empty(Hd, dtype=torch.float32, device=dev)
            if M:
                w2t = torch.empty(Hd, H, dtype=bf, device=dev)
                L.pack_weight_t(w2, w2t, alpha)                 # (alpha W2)^T
                dH = torch.empty(M, Hd, dtype=bf, device=dev)
                L.proj_fwd([dY], [w2t], dH)                     # dH = dY . (alpha W2)
                dZ = torch.empty(M, Hd, dtype=bf, device=dev)
                L.gelu_bwd(dH, z, dZ, flags, bit)
                L.proj_bwd_dw(dZ, [x], [g1], [1.0])
                L.colsum(dZ, gb, None, ws)
            else:
                g1.zero_()
                gb.zero_()
            dw1.append(g1)
            db1.append(gb)
        out = []
        i = 0
        for use in (use_a, use_v):
            if use:
                out += [dw1[i], db1[i], dw2[i], db2[i]]
                i += 1
            else:
                out += [None, None, None, None]
        return (*out, None)
