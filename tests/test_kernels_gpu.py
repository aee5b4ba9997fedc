"""Kernel-level parity tests, called through the C ABI (ctypes) on a B200.

Gather / splice / masks are compared bit-exactly with index arithmetic done in torch on the CPU;
the projector GEMMs are compared with an fp32 (fp64-accumulated) matmul on the same bf16-rounded inputs
(tolerances written at each assert; north_star: max-rel <= 1e-2, cosine >= 0.9999).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.double().cpu(), ref.double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def cosine(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.double().cpu().flatten(), ref.double().cpu().flatten()
    return float(torch.dot(got, ref) / (got.norm() * ref.norm()).clamp_min(1e-30))


def bf16_randn(gen, *shape):
    return torch.randn(*shape, generator=gen, dtype=torch.float32).to(torch.bfloat16)


# ----------------------------------------------------------------------------------------- gather
def gather_expected(audio, video, ka, kv, ntok, audio_valid=None, video_valid=None, tok_offset=None):
    """CPU index arithmetic: row (b, j) = [audio[b, ka*j : ka*j+ka] ; video[b, kv*j : kv*j+kv]], zero past valid."""
    B = (audio if audio is not None else video).shape[0]
    feats = [(audio, ka, audio_valid), (video, kv, video_valid)]
    K = sum(k * f.shape[2] for f, k, _ in feats if f is not None)
    if tok_offset is None:
        offs = [b * ntok for b in range(B + 1)]
    else:
        offs = tok_offset.tolist()
    M = offs[-1]
    out = torch.zeros(M, K, dtype=torch.bfloat16)
    flags = torch.zeros(M, dtype=torch.uint8)
    for b in range(B):
        for j in range(offs[b + 1] - offs[b]):
            col = 0
            for bit, (f, k, valid) in enumerate(feats):
                if f is None:
                    continue
                D = f.shape[2]
                n = f.shape[1] if valid is None else min(int(valid[b]), f.shape[1])
                for i in range(k):
                    t = k * j + i
                    if t < n:
                        out[offs[b] + j, col + i * D: col + (i + 1) * D] = f[b, t]
                if k * j < n:
                    flags[offs[b] + j] |= 1 << bit
                col += k * D
    return out, flags


@pytest.mark.parametrize("case", ["parity_k1", "stride_4_2", "audio_only_k3", "video_only", "valid_lens", "ragged"])
def test_gather_bit_exact(avc, cuda_dev, case):
    L = avc._lib
    g = torch.Generator().manual_seed(7)
    B = 3
    audio = bf16_randn(g, B, 40, 64)
    video = bf16_randn(g, B, 20, 48)
    ka = kv = 1
    av = vv = toff = None
    ntok = 40
    if case == "stride_4_2":
        ka, kv, ntok = 4, 2, 10
    elif case == "audio_only_k3":
        video, ka, ntok = None, 3, 14  # 40 frames / 3 -> 14 tokens, last one partially zero
    elif case == "video_only":
        audio, ntok = None, 20
    elif case == "valid_lens":
        ka, kv, ntok = 2, 1, 20
        av = torch.tensor([40, 17, 0], dtype=torch.int32)
        vv = torch.tensor([20, 9, 3], dtype=torch.int32)
    elif case == "ragged":
        ka, kv = 4, 2
        av = torch.tensor([40, 22, 8], dtype=torch.int32)
        vv = torch.tensor([20, 11, 4], dtype=torch.int32)
        toff = torch.tensor([0, 10, 16, 18], dtype=torch.int32)
    exp, exp_flags = gather_expected(audio, video, ka, kv, ntok, av, vv, toff)
    dev = cuda_dev
    a_d = None if audio is None else audio.to(dev)
    v_d = None if video is None else video.to(dev)
    out = torch.full(exp.shape, 7.0, dtype=torch.bfloat16, device=dev)
    flags = torch.full((exp.shape[0],), 255, dtype=torch.uint8, device=dev)
    L.gather_fwd(a_d, v_d, ka, kv, B, ntok, out, flags, None if toff is None else toff.to(dev),
                 None if av is None else av.to(dev), None if vv is None else vv.to(dev))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu().view(torch.int16), exp.view(torch.int16))
    assert torch.equal(flags.cpu(), exp_flags)


def test_gather_cls_row_select(avc, cuda_dev):
    """frame_stride = (1+Np)*Dv folds last_hidden_state[:, 0] (clip_whisper_model.py:1141-1142) into the gather."""
    L = avc._lib
    g = torch.Generator().manual_seed(8)
    B, F, Np1, Dv = 2, 6, 5, 32
    hidden = bf16_randn(g, B * F, Np1, Dv)
    cls = hidden[:, 0].reshape(B, F, Dv)
    exp, _ = gather_expected(None, cls, 1, 1, F)
    h_d = hidden.to(cuda_dev)
    view = h_d.view(B, F, Np1 * Dv)[:, :, :Dv]  # strides (F*Np1*Dv, Np1*Dv, 1)
    out = torch.empty(B * F, Dv, dtype=torch.bfloat16, device=cuda_dev)
    L.gather_fwd(None, view, 1, 1, B, F, out)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu().view(torch.int16), exp.view(torch.int16))


# ------------------------------------------------------------------------------------ projector fwd
def proj_expected(a_segs, w_segs, bias0, bias1, flags0, flags1, act):
    acc = None
    for a, w in zip(a_segs, w_segs):
        t = a.double() @ w.double().transpose(-1, -2)
        acc = t if acc is None else acc + t
    if bias0 is not None:
        acc = acc + flags0.double().unsqueeze(-1) * bias0.double()
    if bias1 is not None:
        acc = acc + flags1.double().unsqueeze(-1) * bias1.double()
    if act == 1:
        acc = torch.nn.functional.gelu(acc)  # erf form, as nn.GELU() (modality_connector.py:60)
    return acc


@pytest.mark.parametrize("B,R,K,N", [(1, 128, 64, 256), (1, 128, 256, 256), (2, 150, 192, 320), (1, 1000, 520, 2048),
                                     (3, 37, 72, 40)])
@pytest.mark.parametrize("out_fp32", [False, True])
def test_proj_fwd_single_segment(avc, cuda_dev, B, R, K, N, out_fp32):
    L = avc._lib
    g = torch.Generator().manual_seed(B * 1000 + R + K + N)
    a = bf16_randn(g, B, R, K)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g) * 0.5
    ref = proj_expected([a], [w], bias, None, torch.ones(B, R), None, 0)
    y = torch.full((B, R, N), float("nan"), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=cuda_dev)
    L.proj_fwd([a.to(cuda_dev)], [w.to(cuda_dev)], y, bias0=bias.to(cuda_dev))
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    # fp32 accumulate over bf16 inputs: fp32 output is near-exact, bf16 output carries one rounding (2^-9)
    tol = 2e-5 if out_fp32 else 6e-3
    assert rel_err(y, ref) <= tol, rel_err(y, ref)
    assert cosine(y, ref) >= 0.99999


@pytest.mark.parametrize("act", [0, 1])
def test_proj_fwd_two_segments_bias_mask(avc, cuda_dev, act):
    """sum fusion as one GEMM over [a ; v] with the video bias masked on padded rows."""
    L = avc._lib
    g = torch.Generator().manual_seed(11)
    B, R, Ka, Kv, N = 2, 200, 192, 128, 512
    a = bf16_randn(g, B, R, Ka)
    v = bf16_randn(g, B, R, Kv)
    Tv = 90
    v[:, Tv:] = 0
    wa = (torch.randn(N, Ka, generator=g) / Ka ** 0.5).to(torch.bfloat16)
    wv = (torch.randn(N, Kv, generator=g) / Kv ** 0.5).to(torch.bfloat16)
    b0 = torch.randn(N, generator=g)
    b1 = torch.randn(N, generator=g)
    rows = torch.arange(R).expand(B, R)
    ref = proj_expected([a, v], [wa, wv], b0, b1, rows < R, rows < Tv, act)
    y = torch.empty(B, R, N, dtype=torch.bfloat16, device=cuda_dev)
    L.proj_fwd([a.to(cuda_dev), v.to(cuda_dev)], [wa.to(cuda_dev), wv.to(cuda_dev)], y, bias0=b0.to(cuda_dev),
               bias1=b1.to(cuda_dev), flag_rows0=R, flag_rows1=Tv, act=act)
    torch.cuda.synchronize()
    assert rel_err(y, ref) <= 6e-3, rel_err(y, ref)
    assert cosine(y, ref) >= 0.99999
    # packed rows + explicit row flags give the same bits
    flags = ((rows < R).to(torch.uint8) | ((rows < Tv).to(torch.uint8) << 1)).reshape(-1).to(cuda_dev)
    y2 = torch.empty(B * R, N, dtype=torch.bfloat16, device=cuda_dev)
    L.proj_fwd([a.reshape(B * R, Ka).to(cuda_dev), v.reshape(B * R, Kv).to(cuda_dev)],
               [wa.to(cuda_dev), wv.to(cuda_dev)], y2, bias0=b0.to(cuda_dev), bias1=b1.to(cuda_dev), row_flags=flags,
               act=act)
    torch.cuda.synchronize()
    assert torch.equal(y2.view(B, R, N), y)


def test_proj_fwd_column_slices_of_gathered_matrix(avc, cuda_dev):
    """A segments may be column slices of one gathered [M, K] matrix (row stride K)."""
    L = avc._lib
    g = torch.Generator().manual_seed(12)
    M, Ka, Kv, N = 300, 256, 128, 256
    A = bf16_randn(g, M, Ka + Kv)
    W = (torch.randn(N, Ka + Kv, generator=g) / 20).to(torch.bfloat16)
    ref = A.double() @ W.double().T
    A_d, W_d = A.to(cuda_dev), W.to(cuda_dev)
    y = torch.empty(M, N, dtype=torch.float32, device=cuda_dev)
    L.proj_fwd([A_d[:, :Ka], A_d[:, Ka:]], [W_d[:, :Ka], W_d[:, Ka:]], y)
    torch.cuda.synchronize()
    assert rel_err(y, ref) <= 2e-5
    y1 = torch.empty(M, N, dtype=torch.float32, device=cuda_dev)
    L.proj_fwd([A_d], [W_d], y1)
    torch.cuda.synchronize()
    assert rel_err(y1, ref) <= 2e-5


@pytest.mark.parametrize("cta_group,mt", [("1", "1"), ("2", "1"), ("2", "2")])
@pytest.mark.parametrize("workers", ["3", "5"])
def test_gemm_multi_round_schedule_with_tail_split(avc, cuda_dev, monkeypatch, cta_group, mt, workers):
    """Persistent schedule with several rounds per worker and a partial last round (cut into sub-tiles), for the
    single-CTA kernel and the CTA-pair kernel with 256- and 512-row tiles, forward (TN, bf16 + fp32 out) and dW (NT)."""
    L = avc._lib
    monkeypatch.setenv("AVC_GEMM_CTA_GROUP", cta_group)
    monkeypatch.setenv("AVC_GEMM_MT", mt)  # 256- and 512-row pair tiles, for both GEMM modes
    monkeypatch.setenv("AVC_GEMM_MAX_WORKERS", workers)
    g = torch.Generator().manual_seed(77)
    B, R, K, N = 2, 300, 136, 1000
    a = bf16_randn(g, B, R, K)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    ref = proj_expected([a], [w], bias, None, torch.ones(B, R), None, 0)
    for dt, tol in [(torch.float32, 2e-5), (torch.bfloat16, 6e-3)]:
        y = torch.full((B, R, N), float("nan"), dtype=dt, device=cuda_dev)
        L.proj_fwd([a.to(cuda_dev)], [w.to(cuda_dev)], y, bias0=bias.to(cuda_dev))
        torch.cuda.synchronize()
        assert torch.isfinite(y).all()
        assert rel_err(y, ref) <= tol, rel_err(y, ref)
    dy = bf16_randn(g, B, R, 520)
    xs = [bf16_randn(g, B, R, 712), bf16_randn(g, B, R, 200)]
    refs = [torch.einsum("brh,brk->hk", dy.double(), x.double()) for x in xs]
    dws = [torch.full((520, x.shape[2]), float("nan"), dtype=torch.float32, device=cuda_dev) for x in xs]
    L.proj_bwd_dw(dy.to(cuda_dev), [x.to(cuda_dev) for x in xs], dws, [1.0, 1.0])
    torch.cuda.synchronize()
    for dw, r in zip(dws, refs):
        assert torch.isfinite(dw).all()
        assert rel_err(dw, r) <= 2e-5, rel_err(dw, r)


# ------------------------------------------------------------------------------------ projector bwd
@pytest.mark.parametrize("B,R,H,Ka,Kv,base", [(1, 64, 128, 256, 0, 0), (2, 150, 320, 192, 72, 0), (2, 100, 256, 256, 128, 16),
                                              (1, 1000, 512, 520, 0, 0)])
def test_proj_bwd_dw(avc, cuda_dev, B, R, H, Ka, Kv, base):
    L = avc._lib
    g = torch.Generator().manual_seed(R + H + Ka)
    dy_full = bf16_randn(g, B, base + R, H)
    xs = [bf16_randn(g, B, R, Ka)] + ([bf16_randn(g, B, R, Kv)] if Kv else [])
    alpha = [0.5, 0.25][: len(xs)]
    dy = dy_full[:, base:]
    refs = [al * torch.einsum("brh,brk->hk", dy.double(), x.double()) for al, x in zip(alpha, xs)]
    dws = [torch.full((H, x.shape[2]), float("nan"), dtype=torch.float32, device=cuda_dev) for x in xs]
    L.proj_bwd_dw(dy_full.to(cuda_dev), [x.to(cuda_dev) for x in xs], dws, alpha, dy_row_base=base)
    torch.cuda.synchronize()
    for dw, ref in zip(dws, refs):
        assert torch.isfinite(dw).all()
        assert rel_err(dw, ref) <= 2e-5, rel_err(dw, ref)


def test_proj_bwd_dw_ragged_segment_rows(avc, cuda_dev):
    """video segment shorter than the audio one (parity mode, k=1): missing rows contribute zero."""
    L = avc._lib
    g = torch.Generator().manual_seed(5)
    B, Ra, Rv, H, Ka, Kv = 2, 130, 70, 256, 128, 64
    dy = bf16_randn(g, B, Ra, H)
    xa, xv = bf16_randn(g, B, Ra, Ka), bf16_randn(g, B, Rv, Kv)
    ref_a = torch.einsum("brh,brk->hk", dy.double(), xa.double())
    ref_v = torch.einsum("brh,brk->hk", dy[:, :Rv].double(), xv.double())
    dwa = torch.empty(H, Ka, dtype=torch.float32, device=cuda_dev)
    dwv = torch.empty(H, Kv, dtype=torch.float32, device=cuda_dev)
    L.proj_bwd_dw(dy.to(cuda_dev), [xa.to(cuda_dev), xv.to(cuda_dev)], [dwa, dwv], [1.0, 1.0])
    torch.cuda.synchronize()
    assert rel_err(dwa, ref_a) <= 2e-5
    assert rel_err(dwv, ref_v) <= 2e-5


def test_colsum_bias_grad(avc, cuda_dev):
    L = avc._lib
    g = torch.Generator().manual_seed(3)
    B, R, H, Tv = 3, 211, 328, 77
    dy = bf16_randn(g, B, R, H)
    ref0 = 0.5 * dy.double().sum((0, 1))
    ref1 = 0.25 * dy[:, :Tv].double().sum((0, 1))
    o0 = torch.empty(H, dtype=torch.float32, device=cuda_dev)
    o1 = torch.empty(H, dtype=torch.float32, device=cuda_dev)
    ws = L.colsum_workspace(H, cuda_dev)
    L.colsum(dy.to(cuda_dev), o0, o1, ws, flag_rows0=R, flag_rows1=Tv, alpha0=0.5, alpha1=0.25)
    torch.cuda.synchronize()
    assert rel_err(o0, ref0) <= 1e-5
    assert rel_err(o1, ref1) <= 1e-5
    # deterministic: a second run gives identical bits
    o0b = torch.empty_like(o0)
    L.colsum(dy.to(cuda_dev), o0b, None, ws, flag_rows0=R, flag_rows1=Tv, alpha0=0.5, alpha1=0.25)
    torch.cuda.synchronize()
    assert torch.equal(o0, o0b)


def test_pack_weight_bit_exact(avc, cuda_dev):
    L = avc._lib
    g = torch.Generator().manual_seed(4)
    H, Ka, Kv = 96, 64, 40
    wa, wv = torch.randn(H, Ka, generator=g), torch.randn(H, Kv, generator=g)
    packed = torch.zeros(H, Ka + Kv, dtype=torch.bfloat16, device=cuda_dev)
    L.pack_weight(wa.to(cuda_dev), packed[:, :Ka], 0.5)
    L.pack_weight(wv.to(cuda_dev), packed[:, Ka:], 0.3)
    torch.cuda.synchronize()
    exp = torch.cat([(wa * 0.5).to(torch.bfloat16), (wv * torch.tensor(0.3, dtype=torch.float32)).to(torch.bfloat16)], 1)
    assert torch.equal(packed.cpu().view(torch.int16), exp.view(torch.int16))


# ----------------------------------------------------------------------------------------- splice
def splice_expected(ids, ph, pad, y, offs, table, mask_mode, label_mode, labels_in):
    B, S = ids.shape
    H = y.shape[1]
    emb = torch.zeros(B, S, H, dtype=torch.bfloat16)
    mask = torch.ones(B, S, dtype=torch.int64)
    labels = torch.full((B, S), -100, dtype=torch.int64)
    row_of = torch.full((B, S), -1, dtype=torch.int64)
    for b in range(B):
        rank, ntok = 0, offs[b + 1] - offs[b]
        for p in range(S):
            t = int(ids[b, p])
            is_ph = t == ph
            has = False
            if is_ph:
                if rank < ntok:
                    emb[b, p] = y[offs[b] + rank]
                    row_of[b, p] = offs[b] + rank
                    has = True
                rank += 1
            elif table is not None and 0 <= t < table.shape[0]:
                emb[b, p] = table[t]
            if mask_mode == 1:
                mask[b, p] = int(has) if is_ph else int(t != pad)
            lv = -100
            if labels_in is not None and p < labels_in.shape[1]:
                lv = int(labels_in[b, p])
            elif labels_in is None and label_mode == 1:
                lv = t
            if lv == pad:
                lv = -100
            if label_mode == 1 and (is_ph or t == pad):
                lv = -100
            labels[b, p] = lv
    return emb, mask, labels, row_of


@pytest.mark.parametrize("layout", ["prompt_then_av", "ragged_interleaved"])
def test_splice_fwd_bwd_bit_exact(avc, cuda_dev, layout):
    L = avc._lib
    g = torch.Generator().manual_seed(21)
    H, V, PH, PAD = 136, 50, 49, 0
    if layout == "prompt_then_av":
        B, P, T = 3, 5, 70
        S = P + T
        ids = torch.cat([torch.randint(1, V - 1, (B, P), generator=g), torch.full((B, T), PH)], 1)
        offs = [0, T, 2 * T, 3 * T]
        toff = None
        mask_mode = label_mode = 0
        labels_in = torch.randint(0, V - 1, (B, 40), generator=g)  # shorter than S: right-pad with -100
    else:
        B, S = 4, 83
        ns = [20, 0, 33, 7]
        ids = torch.full((B, S), PAD, dtype=torch.int64)
        for b, n in enumerate(ns):
            row = torch.randint(1, V - 1, (S,), generator=g)
            start = 3 + 2 * b
            row[start:start + n] = PH
            row[start + n + 10:] = PAD
            ids[b] = row
        offs = [0]
        for n in ns:
            offs.append(offs[-1] + n)
        toff = torch.tensor(offs, dtype=torch.int32)
        mask_mode = label_mode = 1
        labels_in = None
    M = offs[-1]
    y = bf16_randn(g, M, H)
    table = bf16_randn(g, V, H)
    emb_e, mask_e, labels_e, row_of = splice_expected(ids, PH, PAD, y, offs, table, mask_mode, label_mode, labels_in)
    dev = cuda_dev
    ids_d = ids.to(dev)
    emb = torch.full((B, S, H), 3.0, dtype=torch.bfloat16, device=dev)
    mask = torch.full((B, S), -7, dtype=torch.int64, device=dev)
    labels = torch.full((B, S), -7, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    sp = L.make_splice(ids_d, PH, PAD, H, tokens_per_sample=0 if toff is not None else offs[1],
                       tok_offset=None if toff is None else toff.to(dev), embed_table=table.to(dev),
                       attention_mask=mask, mask_mode=mask_mode, label_mode=label_mode,
                       labels_in=None if labels_in is None else labels_in.to(dev), labels_out=labels, status=status)
    y_d = y.to(dev)
    L.splice_fwd(sp, y_d, emb)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    assert torch.equal(emb.cpu().view(torch.int16), emb_e.view(torch.int16))
    assert torch.equal(mask.cpu(), mask_e)
    assert torch.equal(labels.cpu(), labels_e)
    # backward: dY rows are the placeholder rows of d(inputs_embeds)
    d_emb = bf16_randn(g, B, S, H)
    dy_e = torch.zeros(M, H, dtype=torch.bfloat16)
    sel = row_of >= 0
    dy_e[row_of[sel]] = d_emb[sel]
    dy = torch.full((M, H), 9.0, dtype=torch.bfloat16, device=dev)
    L.splice_bwd(sp, d_emb.to(dev), dy)
    torch.cuda.synchronize()
    assert torch.equal(dy.cpu().view(torch.int16), dy_e.view(torch.int16))


def test_splice_reports_placeholder_mismatch(avc, cuda_dev):
    L = avc._lib
    ids = torch.tensor([[1, 9, 9, 9, 2]], dtype=torch.int64, device=cuda_dev)
    status = torch.zeros(1, dtype=torch.int32, device=cuda_dev)
    emb = torch.zeros(1, 5, 8, dtype=torch.bfloat16, device=cuda_dev)
    y = torch.zeros(2, 8, dtype=torch.bfloat16, device=cuda_dev)
    sp = L.make_splice(ids, 9, 0, 8, tokens_per_sample=2, status=status)
    L.splice_fwd(sp, y, emb)
    torch.cuda.synchronize()
    assert int(status.item()) == 1


def test_errors_are_loud(avc, cuda_dev):
    L = avc._lib
    a = torch.zeros(4, 12, dtype=torch.bfloat16, device=cuda_dev)  # K = 12: not a multiple of 8
    w = torch.zeros(8, 12, dtype=torch.bfloat16, device=cuda_dev)
    y = torch.zeros(4, 8, dtype=torch.bfloat16, device=cuda_dev)
    with pytest.raises(L.ConnectorError):
        L.proj_fwd([a], [w], y)
