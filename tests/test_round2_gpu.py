"""Round-2 parity tests on a B200 (through the C ABI): bias gradients from the dW launch, fp16 output, input gradients
(unfrozen towers), the TMA bulk splice at piece boundaries, ragged offsets formed on the device, checkpoint round
trips, weight-pack cache invalidation, the `adaptive` connector, the public-API gradient bucket, and the full-size
engine check bench.py runs.

Tolerances: index / byte work bit-exact; projected rows and gradients max-rel <= 1e-2, cosine >= 0.9999 against the
fp32 reference (north_star); kernel-level fp32 outputs against fp64 recomputation of the same bf16 operands <= 2e-5."""
import os
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn

from oracle import connector_oracle as O

pytestmark = pytest.mark.gpu
REL_TOL, COS_TOL = 1e-2, 0.9999


def rel_err(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def cosine(got, ref):
    got, ref = got.detach().double().cpu().flatten(), ref.detach().double().cpu().flatten()
    return float(torch.dot(got, ref) / (got.norm() * ref.norm()).clamp_min(1e-30))


def assert_close(got, ref, what, rel=REL_TOL):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    r, c = rel_err(got, ref), cosine(got, ref)
    assert r <= rel and c >= COS_TOL, f"{what}: max-rel {r:.3e}, cosine {c:.6f}"


def _rand_params(g, H, Ka, Kv):
    return (torch.randn(H, Ka, generator=g) / Ka ** 0.5, torch.randn(H, generator=g) * 0.1,
            torch.randn(H, Kv, generator=g) / Kv ** 0.5, torch.randn(H, generator=g) * 0.1)


# ------------------------------------------------------------------------------------------ db inside the dW launch
@pytest.mark.parametrize("B,N,H,Ka,Kv,base,max_workers", [
    (3, 100, 256, 128, 64, 0, 0),       # fewer tiles than workers
    (2, 375, 1024, 512, 256, 16, 0),    # row base (prompt rows skipped), several M blocks
    (4, 130, 1280, 768, 0, 5, 3),       # one segment, forced multi-round schedule, H not a multiple of 512
])
def test_dw_db_one_launch_matches_fp64_and_colsum(avc, cuda_dev, monkeypatch, B, N, H, Ka, Kv, base, max_workers):
    L = avc._lib
    if max_workers:
        monkeypatch.setenv("AVC_GEMM_MAX_WORKERS", str(max_workers))
    g = torch.Generator().manual_seed(B * 1000 + N)
    dev = cuda_dev
    S = base + N + 3
    dy = torch.randn(B, S, H, generator=g).to(dev, torch.bfloat16)
    xs = [torch.randn(B, N, Ka, generator=g).to(dev, torch.bfloat16)]
    if Kv:
        xs.append(torch.randn(B, N, Kv, generator=g).to(dev, torch.bfloat16))
    # token-present operand: audio on all rows, video only on a prefix of every sample's rows (zero-padded stream)
    nv = N // 2 + 1
    present = torch.zeros(B, N, L.BIAS_COLS, dtype=torch.bfloat16, device=dev)
    present[:, :, 0] = 1
    present[:, :nv, 1] = 1
    present[:, :, 2:] = 7.0  # columns >= 2 are ignored
    dws = [torch.full((H, x.shape[2]), float("nan"), device=dev) for x in xs]
    db0 = torch.full((H,), float("nan"), device=dev)
    db1 = torch.full((H,), float("nan"), device=dev)
    al = [0.7, 1.3][:len(xs)]
    L.proj_bwd_dw(dy, xs, dws, al, dy_row_base=base, bias=(present, db0, db1, 0.7, 1.3))
    torch.cuda.synchronize()
    dy64 = dy[:, base:base + N].double()
    for x, dw, a in zip(xs, dws, al):
        ref = a * torch.einsum("bnh,bnk->hk", dy64, x.double())
        assert rel_err(dw, ref) <= 2e-5
    ref0 = 0.7 * dy64.sum((0, 1))
    ref1 = 1.3 * dy64[:, :nv].sum((0, 1))
    assert rel_err(db0, ref0) <= 2e-5 and rel_err(db1, ref1) <= 2e-5
    # the stand-alone column-sum kernel computes the same sums (different summation tree: fp32 re-association only)
    c0, c1 = torch.empty(H, device=dev), torch.empty(H, device=dev)
    L.colsum(dy, c0, c1, L.colsum_workspace(H, dev), flag_rows0=N, flag_rows1=nv, alpha0=0.7, alpha1=1.3,
             dy_row_base=base, sum_rows=N)
    torch.cuda.synchronize()
    assert rel_err(db0, c0) <= 2e-5 and rel_err(db1, c1) <= 2e-5
    # deterministic: a second launch gives the same bits
    d2, e0, e1 = [torch.empty_like(t) for t in dws], torch.empty_like(db0), torch.empty_like(db1)
    L.proj_bwd_dw(dy, xs, d2, al, dy_row_base=base, bias=(present, e0, e1, 0.7, 1.3))
    torch.cuda.synchronize()
    assert all(torch.equal(a_, b_) for a_, b_ in zip(dws, d2)) and torch.equal(db0, e0) and torch.equal(db1, e1)


@pytest.mark.parametrize("forced", [0, 2, 3, 8])
def test_dw_split_reduction_is_deterministic_and_matches_fp64(avc, cuda_dev, monkeypatch, forced):
    """Few tiles (2 x 3 + 2 bias items on 74 CTA pairs): the row reduction is cut into slices whose partial tiles are
    added in slice order by the last finisher.  Forced slice counts and the planner's own choice; dW, db against
    fp64; two launches give the same bits; the workspace counters are left zero; same values as the unsplit launch."""
    L = avc._lib
    if forced:
        monkeypatch.setenv("AVC_GEMM_KSPLIT", str(forced))
    g = torch.Generator().manual_seed(17)
    dev = cuda_dev
    B, N, P, H, Ka, Kv = 5, 530, 3, 1000, 512, 200   # H, Kv not multiples of the tile: row / column guards
    dy = torch.randn(B, P + N, H, generator=g).to(dev, torch.bfloat16)
    xs = [torch.randn(B, N, Ka, generator=g).to(dev, torch.bfloat16), torch.randn(B, N, Kv, generator=g).to(dev, torch.bfloat16)]
    present = L.present_operand(B, N, dev)
    splits, nbytes = L.dw_plan(dy, xs, True)
    if forced:
        assert splits == forced and nbytes > 0
    else:
        assert splits > 1, "8 work items on 74 workers: the planner must split"
    outs = []
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    for rep in range(2):
        dws = [torch.full((H, x.shape[2]), float("nan"), device=dev) for x in xs]
        db0, db1 = torch.full((H,), float("nan"), device=dev), torch.full((H,), float("nan"), device=dev)
        L.proj_bwd_dw(dy, xs, dws, [0.5, 2.0], dy_row_base=P, bias=(present, db0, db1, 0.5, 2.0), workspace=ws)
        torch.cuda.synchronize()
        outs.append((dws, db0, db1))
    dy64 = dy[:, P:].double()
    for x, dw, a in zip(xs, outs[0][0], [0.5, 2.0]):
        assert rel_err(dw, a * torch.einsum("bnh,bnk->hk", dy64, x.double())) <= 2e-5
    assert rel_err(outs[0][1], 0.5 * dy64.sum((0, 1))) <= 2e-5 and rel_err(outs[0][2], 2.0 * dy64.sum((0, 1))) <= 2e-5
    for a_, b_ in zip(outs[0][0] + [outs[0][1], outs[0][2]], outs[1][0] + [outs[1][1], outs[1][2]]):
        assert torch.equal(a_, b_), "slice-order summation: same bits every launch"
    counters = ws[:8 * 8 * 4].view(torch.int32)   # (6 tiles + 2 bias items) x 8 warp slots
    assert int(counters.abs().max()) == 0, "every launch leaves its arrival counters zero"
    # unsplit launch: same values up to fp32 re-association
    monkeypatch.setenv("AVC_GEMM_KSPLIT", "1")
    d1 = [torch.empty_like(t) for t in outs[0][0]]
    e0, e1 = torch.empty_like(outs[0][1]), torch.empty_like(outs[0][2])
    L.proj_bwd_dw(dy, xs, d1, [0.5, 2.0], dy_row_base=P, bias=(present, e0, e1, 0.5, 2.0), split=False)
    torch.cuda.synchronize()
    for a_, b_ in zip(outs[0][0] + [outs[0][1], outs[0][2]], d1 + [e0, e1]):
        assert rel_err(a_, b_) <= 2e-5


def test_dw_db_with_row_flags_operand(avc, cuda_dev):
    """Packed rows + uint8 row flags (what the gather writes) -> present operand -> db of each stream."""
    L = avc._lib
    g = torch.Generator().manual_seed(77)
    dev = cuda_dev
    M, H, K = 700, 512, 256
    dy = torch.randn(M, H, generator=g).to(dev, torch.bfloat16)
    x = torch.randn(M, K, generator=g).to(dev, torch.bfloat16)
    flags = torch.randint(0, 4, (M,), generator=g).to(torch.uint8).to(dev)
    present = L.present_operand(1, M, dev, row_flags=flags)
    dw, d0, d1 = torch.empty(H, K, device=dev), torch.empty(H, device=dev), torch.empty(H, device=dev)
    L.proj_bwd_dw(dy, [x], [dw], [1.0], bias=(present, d0, d1, 1.0, 1.0))
    torch.cuda.synchronize()
    f = flags.long()
    assert rel_err(d0, (dy.double() * (f & 1)[:, None]).sum(0)) <= 2e-5
    assert rel_err(d1, (dy.double() * ((f >> 1) & 1)[:, None]).sum(0)) <= 2e-5
    assert rel_err(dw, dy.double().t() @ x.double()) <= 2e-5


def test_bias_in_gemm_equals_colsum_path_through_the_public_api(avc, cuda_dev, monkeypatch):
    g = torch.Generator().manual_seed(3)
    B, Ta, Tv, Da, Dv, H, P, V = 2, 48, 20, 64, 32, 256, 4, 30
    a, v = torch.randn(B, Ta, Da, generator=g), torch.randn(B, Tv, Dv, generator=g)
    wa, ba, wv, bv = _rand_params(g, H, Da, Dv)
    prompt = torch.randint(1, V, (B, P), generator=g)
    table = torch.randn(V, H, generator=g).to(cuda_dev, torch.bfloat16)
    up = torch.randn(B, P + Ta, H, generator=g).to(cuda_dev)
    plan = avc.FusePlan(fusion_scale=0.25, max_seq_len=64)  # video shorter: zero-padded rows must not see bv
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("AVC_BIAS_IN_GEMM", mode)
        params = [t.to(cuda_dev).requires_grad_(True) for t in (wa, ba, wv, bv)]
        emb, _, _ = avc.fused_connector(a.to(cuda_dev), v.to(cuda_dev), *params, plan, prompt_ids=prompt.to(cuda_dev),
                                        embed_table=table, out_dtype=torch.bfloat16)
        (emb.float() * up).sum().backward()
        torch.cuda.synchronize()
        res[mode] = [p.grad.clone() for p in params]
    for x, y, n in zip(res["1"], res["0"], ["dWa", "dba", "dWv", "dbv"]):
        assert rel_err(x, y) <= 2e-5, n
    spec = O.ConnectorSpec(fusion_scale=0.25, max_seq_len=64)
    for x, y, n in zip(res["1"], O.connector_grads(a, v, wa, ba, wv, bv, spec, up[:, P:].cpu()), ["dWa", "dba", "dWv", "dbv"]):
        assert_close(x, y, n)


# ------------------------------------------------------------------------------------------ fp16 (the reference's use_fp16)
def test_proj_fwd_fp16_output(avc, cuda_dev):
    L = avc._lib
    g = torch.Generator().manual_seed(11)
    M, K, H = 300, 192, 320
    x = torch.randn(M, K, generator=g).to(cuda_dev, torch.bfloat16)
    w = (torch.randn(H, K, generator=g) / K ** 0.5).to(cuda_dev, torch.bfloat16)
    b = torch.randn(H, generator=g).to(cuda_dev)
    y = torch.full((M, H), float("nan"), dtype=torch.float16, device=cuda_dev)
    L.proj_fwd([x], [w], y, bias0=b)
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t() + b.double()
    assert y.dtype == torch.float16
    assert rel_err(y, ref) <= 1e-3          # fp16 rounding of the result (2^-11 relative per element)


def test_cast_bf16_kernel(avc, cuda_dev):
    L = avc._lib
    g = torch.Generator().manual_seed(12)
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        src = torch.randn(37, 136, generator=g).to(cuda_dev, dt)
        wide = torch.zeros(37, 200, dtype=torch.bfloat16, device=cuda_dev)
        L.cast_bf16(src[:, 8:136], wide[:, 16:144], 0.5)
        torch.cuda.synchronize()
        assert torch.equal(wide[:, 16:144], (src[:, 8:136].float() * 0.5).to(torch.bfloat16))
        assert float(wide[:, :16].abs().max()) == 0 and float(wide[:, 144:].abs().max()) == 0


@pytest.mark.parametrize("uniform", [True, False])
def test_fused_connector_fp16_llm(avc, cuda_dev, uniform):
    """use_fp16=True in the reference = fp16 LLM, fp16 inputs_embeds (clip_whisper_model.py:164, 454-458)."""
    g = torch.Generator().manual_seed(21)
    B, Ta, Tv, Da, Dv, H, P, V = 2, 40, 20, 64, 32, 128, 5, 40
    a, v = torch.randn(B, Ta, Da, generator=g), torch.randn(B, Tv, Dv, generator=g)
    wa, ba, wv, bv = _rand_params(g, H, 2 * Da, Dv)
    table = torch.randn(V, H, generator=g)
    prompt = torch.randint(1, V, (B, P), generator=g)
    labels = torch.randint(0, V, (B, 12), generator=g)
    spec = O.ConnectorSpec(fusion="concat", audio_stride=2, video_stride=1, max_seq_len=64)
    emb_r, mask_r, lab_r, _ = O.connector_forward(a, v, wa, ba, wv, bv, spec, prompt_ids=prompt, embed_table=table,
                                                  labels=labels)
    N = emb_r.shape[1] - P
    up = torch.randn(B, P + N, H, generator=g)
    grads_r = O.connector_grads(a, v, wa, ba, wv, bv, spec, up[:, P:])
    dev = cuda_dev
    params = [t.to(dev).requires_grad_(True) for t in (wa, ba, wv, bv)]
    plan = avc.FusePlan(fusion="concat", audio_stride=2, video_stride=1, max_seq_len=64)
    kw = dict(embed_table=table.to(dev, torch.float16), labels=labels.to(dev), out_dtype=torch.float16, check=True)
    if uniform:
        emb, mask, lab = avc.fused_connector(a.to(dev).half(), v.to(dev).half(), *params, plan,
                                             prompt_ids=prompt.to(dev), **kw)
    else:
        ids = torch.cat([prompt, torch.full((B, N), V + 5)], 1).to(dev)
        emb, mask, lab = avc.fused_connector(a.to(dev).half(), v.to(dev).half(), *params, plan, input_ids=ids,
                                             placeholder_id=V + 5, **kw)
    assert emb.dtype == torch.float16
    (emb.float() * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    assert_close(emb[:, P:], emb_r[:, P:], "AV rows")
    assert torch.equal(emb[:, :P].cpu(), emb_r[:, :P].to(torch.float16))
    assert torch.equal(mask.cpu(), mask_r) and torch.equal(lab.cpu(), lab_r)
    for p, gr, n in zip(params, grads_r, ["dWa", "dba", "dWv", "dbv"]):
        assert_close(p.grad, gr, n)


def test_modality_connector_fp16_and_adaptive_projection_fp16(avc, cuda_dev):
    g = torch.Generator().manual_seed(22)
    conn = avc.ModalityConnector(64, 96, device="cuda:0", dtype=torch.float16)
    x = torch.randn(2, 30, 64, generator=g)
    y = conn(x.to(cuda_dev).half())
    assert y.dtype == torch.float16
    ref = O.reference_connector(x, conn.linear.weight.detach().cpu(), conn.linear.bias.detach().cpu())
    assert_close(y, ref, "fp16 connector")
    t = torch.randn(2, 28, 64, generator=g)
    out = avc.adaptive_projection(t.to(cuda_dev).half(), 10)
    assert out.dtype == torch.float16
    assert_close(out, O.reference_adaptive_projection(t.half().float(), 10), "fp16 pool", rel=2e-3)


# ------------------------------------------------------------------------------------------ input gradients (dX)
def test_linear_connector_input_gradient(avc, cuda_dev):
    g = torch.Generator().manual_seed(31)
    conn = avc.ModalityConnector(64, 96, device="cuda:0")
    with torch.no_grad():
        conn.linear.bias.copy_(torch.randn(96, generator=g))
    x = torch.randn(3, 50, 64, generator=g)
    up = torch.randn(3, 50, 96, generator=g)
    xc = x.clone().requires_grad_(True)
    w, b = conn.linear.weight.detach().cpu().clone().requires_grad_(True), conn.linear.bias.detach().cpu().clone().requires_grad_(True)
    (O.reference_connector(xc, w, b) * up).sum().backward()
    xd = x.to(cuda_dev).requires_grad_(True)
    (conn(xd) * up.to(cuda_dev)).sum().backward()
    torch.cuda.synchronize()
    assert xd.grad.dtype == torch.float32 and xd.grad.shape == x.shape
    assert_close(xd.grad, xc.grad, "dX")
    assert_close(conn.linear.weight.grad, w.grad, "dW")
    assert_close(conn.linear.bias.grad, b.grad, "db")


@pytest.mark.parametrize("case", ["direct", "gather_pad", "gather_ragged_repeat"])
def test_fused_connector_input_gradients(avc, cuda_dev, case):
    """freeze_encoders=False: d(audio), d(video) through the dX GEMM and the gather's transpose vs CPU autograd."""
    g = torch.Generator().manual_seed(32)
    B, Da, Dv, H, P, V = 3, 64, 32, 128, 4, 40
    kw, okw = {}, {}
    if case == "direct":
        Ta, Tv, ka, kv, ra, rv = 40, 20, 4, 2, 1, 1
    elif case == "gather_pad":
        Ta, Tv, ka, kv, ra, rv = 41, 13, 4, 2, 1, 1   # ragged tails + video shorter (zero-padded)
    else:
        Ta, Tv, ka, kv, ra, rv = 30, 15, 1, 1, 1, 2   # rate alignment at stride 1 (every video frame used twice) + lengths
        kw = dict(audio_lengths=[30, 11, 2], video_lengths=[15, 6, 1])
        okw = dict(audio_valid=torch.tensor([30, 11, 2]), video_valid=torch.tensor([15, 6, 1]))
    a, v = torch.randn(B, Ta, Da, generator=g), torch.randn(B, Tv, Dv, generator=g)
    wa, ba, wv, bv = _rand_params(g, H, ka * Da, kv * Dv)
    spec = O.ConnectorSpec(fusion="sum", fusion_scale=0.4, audio_stride=ka, video_stride=kv, audio_repeat=ra,
                           video_repeat=rv, max_seq_len=64)
    ac, vc = a.clone().requires_grad_(True), v.clone().requires_grad_(True)
    pc = [t.clone().requires_grad_(True) for t in (wa, ba, wv, bv)]
    tok, _ = O.connector_tokens(ac, vc, *pc, spec, **okw)
    N = tok.shape[1]
    up = torch.randn(B, N, H, generator=g)
    dev = cuda_dev
    plan = avc.FusePlan(fusion="sum", fusion_scale=0.4, audio_stride=ka, video_stride=kv, audio_repeat=ra,
                        video_repeat=rv, max_seq_len=64)
    ad, vd = a.to(dev).requires_grad_(True), v.to(dev).requires_grad_(True)
    params = [t.to(dev).requires_grad_(True) for t in (wa, ba, wv, bv)]
    if kw:
        counts = [plan.tokens(la, lv) for la, lv in zip(kw["audio_lengths"], kw["video_lengths"])]
        ids = torch.zeros(B, N + 2, dtype=torch.int64)
        for b_, c in enumerate(counts):
            ids[b_, 1:1 + c] = V + 1
        emb, _, _ = avc.fused_connector(ad, vd, *params, plan, input_ids=ids.to(dev), placeholder_id=V + 1,
                                        out_dtype=torch.bfloat16, check=True, **kw)
        keep = torch.zeros(B, N, dtype=torch.bool)
        for b_, c in enumerate(counts):
            keep[b_, :c] = True
        up = up * keep[:, :, None]
        (tok * up).sum().backward()
        (emb[:, 1:1 + N].float() * up.to(dev)).sum().backward()
    else:
        (tok * up).sum().backward()
        emb, _, _ = avc.fused_connector(ad, vd, *params, plan, out_dtype=torch.bfloat16)
        (emb.float() * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    assert ad.grad.shape == a.shape and vd.grad.shape == v.shape and ad.grad.dtype == torch.float32
    assert_close(ad.grad, ac.grad, "d(audio)")
    assert_close(vd.grad, vc.grad, "d(video)")
    # frames no token reads get an exact zero
    dead = (ac.grad == 0).all(-1)
    if bool(dead.any()):
        assert float(ad.grad.cpu()[dead].abs().max()) == 0.0
    for p, q, n in zip(params, pc, ["dWa", "dba", "dWv", "dbv"]):
        assert_close(p.grad, q.grad, n)


def test_gather_bwd_kernel_bit_exact(avc, cuda_dev):
    """The gather's transpose is index work + fp32 adds of at most `repeat` bf16 values: compare with a host loop."""
    L = avc._lib
    g = torch.Generator().manual_seed(33)
    B, T, D, k, rep, N = 3, 11, 16, 2, 2, 9
    valid = torch.tensor([11, 5, 0], dtype=torch.int32)
    K = k * D + 8
    dA = torch.randn(B * N, K, generator=g).to(torch.bfloat16)
    ref = torch.zeros(B, T, D)
    for b in range(B):
        for t in range(T):
            if t >= int(valid[b]):
                continue
            s, slot = t // k, t % k
            for j in range(s * rep, min((s + 1) * rep, N)):
                ref[b, t] += dA[b * N + j, 8 + slot * D: 8 + (slot + 1) * D].float()
    for dt in (torch.float32, torch.bfloat16):
        out = torch.full((B, T, D), float("nan"), dtype=dt, device=cuda_dev)
        L.gather_bwd(dA.to(cuda_dev), 8, out, k, rep, B, N, valid=valid.to(cuda_dev))
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), ref.to(dt))


# ------------------------------------------------------------------------------------------ splice: TMA bulk pieces
@pytest.mark.parametrize("H,dtype", [(2048, torch.bfloat16), (2600, torch.bfloat16), (4096, torch.float32),
                                     (264, torch.float16)])
def test_splice_bulk_rows_across_piece_boundaries(avc, cuda_dev, H, dtype):
    """Rows of exactly one 4 KB piece, a ragged last piece, four pieces (fp32), and rows smaller than a piece; ragged
    placeholders, missing table rows (zeros), forward and backward bit-exact."""
    L = avc._lib
    g = torch.Generator().manual_seed(H)
    B, S, V, PH = 5, 75, 50, 50
    counts = [33, 0, 64, 7, 40]
    ids = torch.randint(1, V, (B, S), generator=g)
    for b, c in enumerate(counts):
        pos = torch.randperm(S, generator=g)[:c].sort().values
        ids[b, pos] = PH
    free = (ids[0] != PH).nonzero().flatten()
    ids[0, int(free[3])] = 777  # out-of-range id at a text position: zero row
    offs = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32)
    M = int(offs[-1])
    y = torch.randn(M, H, generator=g).to(dtype)
    table = torch.randn(V, H, generator=g).to(dtype)
    dev = cuda_dev
    emb = torch.full((B, S, H), float("nan"), dtype=dtype, device=dev)
    mask = torch.empty(B, S, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    sp = L.make_splice(ids.to(dev), PH, 0, H, tok_offset=offs.to(dev), embed_table=table.to(dev), attention_mask=mask,
                       status=status, elem_size=y.element_size())
    L.splice_fwd(sp, y.to(dev), emb)
    torch.cuda.synchronize()
    ref = torch.zeros(B, S, H, dtype=dtype)
    for b in range(B):
        r = 0
        for p in range(S):
            t = int(ids[b, p])
            if t == PH:
                ref[b, p] = y[int(offs[b]) + r]
                r += 1
            elif 0 <= t < V:
                ref[b, p] = table[t]
    assert int(status.item()) == 0
    assert torch.equal(emb.cpu().view(torch.uint8), ref.view(torch.uint8))
    dy = torch.full((M, H), float("nan"), dtype=dtype, device=dev)
    d_emb = torch.randn(B, S, H, generator=g).to(dtype)
    L.splice_bwd(sp, d_emb.to(dev), dy)
    torch.cuda.synchronize()
    assert torch.equal(dy.cpu().view(torch.uint8), d_emb[ids == PH].view(torch.uint8))


# ------------------------------------------------------------------------------------------ host logic on the device
def test_ragged_offsets_on_device_match_host(avc, cuda_dev):
    from audio_visual_llm_b200.connector_ops import ragged_token_offsets

    plan = avc.FusePlan(audio_stride=4, video_stride=2, max_seq_len=9, video_repeat=1)
    la, lv = [40, 0, 17, 3, 99], [20, 5, 0, 9, 12]
    host = ragged_token_offsets(plan, 5, 40, 20, la, lv, cuda_dev)
    devv = ragged_token_offsets(plan, 5, 40, 20, torch.tensor(la, device=cuda_dev), torch.tensor(lv, device=cuda_dev),
                                cuda_dev)
    hint = ragged_token_offsets(plan, 5, 40, 20, torch.tensor(la, device=cuda_dev), torch.tensor(lv, device=cuda_dev),
                                cuda_dev, total_tokens=host[3])
    for other in (devv, hint):
        assert torch.equal(host[0], other[0]) and host[3] == other[3]
        assert torch.equal(host[1], other[1]) and torch.equal(host[2], other[2])
    with pytest.raises(ValueError):
        ragged_token_offsets(plan, 5, 40, 20, la, lv, cuda_dev, total_tokens=host[3] + 1)


def test_pack_cache_is_invalidated_by_the_raw_pointer_optimizer(avc, cuda_dev):
    """eval -> ConnectorAdamW.step (updates through raw pointers: no version bump) -> eval must see the new weights."""
    from audio_visual_llm_b200.trainer_step import ConnectorAdamW

    g = torch.Generator().manual_seed(5)
    conn = avc.ModalityConnector(64, 96, device="cuda:0")
    x = torch.randn(2, 20, 64, generator=g).to(cuda_dev)
    with torch.no_grad():
        y0 = conn(x).clone()
        assert torch.equal(conn(x), y0)  # cached pack
    y = conn(x)
    y.float().pow(2).sum().backward()
    opt = ConnectorAdamW([("linear.weight", conn.linear.weight), ("linear.bias", conn.linear.bias)], lr=0.05,
                         max_grad_norm=0.0)
    v0 = conn.linear.weight._version
    opt.step()
    assert conn.linear.weight._version == v0, "the kernel writes through raw pointers (this is what the cache must survive)"
    with torch.no_grad():
        y1 = conn(x)
    torch.cuda.synchronize()
    ref = O.reference_connector(x.cpu(), conn.linear.weight.detach().cpu(), conn.linear.bias.detach().cpu())
    assert not torch.equal(y1, y0)
    assert_close(y1, ref, "eval after the optimizer step")


# ------------------------------------------------------------------------------------------ model-level
class _Tok:
    pad_token_id = 0


class _LLM(nn.Module):
    def __init__(self, V, H, dtype):
        super().__init__()
        self.embed = nn.Embedding(V, H).to(dtype)
        self.head = nn.Linear(H, 4).to(dtype)
        self.lora_A = nn.Parameter(torch.zeros(2, 2, dtype=dtype))
        self.seen = None

    def get_input_embeddings(self):
        return self.embed

    def forward(self, inputs_embeds=None, attention_mask=None, labels=None, return_dict=True):
        self.seen = inputs_embeds
        out = self.head(inputs_embeds)
        return SimpleNamespace(loss=out.float().pow(2).mean(), logits=out)


class _Whisper(nn.Module):
    """A 'tower' with a trainable layer, so that freeze_encoders=False has something to train."""

    def __init__(self, d):
        super().__init__()
        self.config = SimpleNamespace(d_model=d)
        self.proj = nn.Linear(80, d)
        outer = self

        class Enc(nn.Module):
            def forward(self, audio, attention_mask=None, output_hidden_states=True, return_dict=True):
                return SimpleNamespace(last_hidden_state=outer.proj(audio.transpose(1, 2)))

        self.encoder = Enc()


def _model(avc, dev, dtype=torch.float32, **kw):
    return avc.ClipWhisperModel(device="cuda:0", modality="audio", _provided_tokenizer=_Tok(),
                                _provided_llm=_LLM(30, 64, dtype).to(dev), _provided_whisper=_Whisper(32).to(dev), **kw)


def test_freeze_flags_and_unfrozen_tower_gradients(avc, cuda_dev):
    m = _model(avc, cuda_dev, freeze_encoders=False, freeze_llm=True, use_lora=True)
    trainable = {n for n, p in m.llm.named_parameters() if p.requires_grad}
    assert trainable == {"lora_A"}, "freeze_llm freezes the LLM except LoRA weights (clip_whisper_model.py:1005-1014)"
    assert all(p.requires_grad for p in m.whisper.parameters())
    m.train()
    audio = torch.randn(2, 80, 24, device=cuda_dev)
    out = m(audio=audio, labels=torch.randint(1, 30, (2, 24), device=cuda_dev))
    out["loss"].backward()
    torch.cuda.synchronize()
    gw = m.whisper.proj.weight.grad
    assert gw is not None and float(gw.abs().max()) > 0, "the connector's dX reaches the unfrozen tower"
    # reference on the CPU: same tower, same connector, same LLM head, fp32
    import copy
    m.llm.seen = None  # a non-leaf tensor cannot be deep-copied
    tower, llm = copy.deepcopy(m.whisper).cpu().float(), copy.deepcopy(m.llm).cpu().float()
    for p in tower.parameters():
        p.grad = None
    feats = tower.proj(audio.cpu().transpose(1, 2))
    emb = O.reference_connector(feats, m.audio_connector.linear.weight.detach().cpu(), m.audio_connector.linear.bias.detach().cpu())
    llm.head(emb).float().pow(2).mean().backward()
    assert_close(gw, tower.proj.weight.grad, "tower gradient through dX", rel=2e-2)
    frozen = _model(avc, cuda_dev)  # default: towers frozen, connector under no input grad
    assert not any(p.requires_grad for p in frozen.whisper.parameters())


def test_use_fp16_gives_fp16_embeds(avc, cuda_dev):
    m = _model(avc, cuda_dev, dtype=torch.float16, use_fp16=True)
    assert m.dtype == torch.float16
    m.eval()
    audio = torch.randn(2, 80, 16, device=cuda_dev).half()
    emb, mask = m.encode(audio=audio)
    assert emb.dtype == torch.float16 and mask.dtype == torch.int64
    feats = m.whisper.proj(audio.float().transpose(1, 2)).detach().cpu()
    ref = O.reference_connector(feats, m.audio_connector.linear.weight.detach().cpu(), m.audio_connector.linear.bias.detach().cpu())
    assert_close(emb, ref, "fp16 inputs_embeds")


def test_prompt_rows_are_differentiable_when_the_embedding_trains(avc, cuda_dev):
    m = _model(avc, cuda_dev)
    m.train()
    audio = torch.randn(2, 80, 8, device=cuda_dev)
    prompt = torch.randint(1, 30, (2, 5), device=cuda_dev)
    m(audio=audio, prompt=prompt, labels=torch.randint(1, 30, (2, 13), device=cuda_dev))["loss"].backward()
    g = m.llm.embed.weight.grad
    assert g is not None and float(g[prompt.flatten()].abs().max()) > 0, "clip_whisper_model.py:484-485 is differentiable"


def test_checkpoint_round_trips(avc, cuda_dev, tmp_path):
    """save_connectors / load_connectors (clip_whisper_model.py:745-746, 856-857) and a reference-format
    model_state_dict filtered by key as decode.py:223-260 does."""
    m1, m2, m3 = _model(avc, cuda_dev), _model(avc, cuda_dev), _model(avc, cuda_dev)
    m1.save_connectors(tmp_path)
    for name in ("audio_connector.pt", "video_connector.pt"):
        sd = torch.load(tmp_path / name)
        assert set(sd) == {"linear.weight", "linear.bias"}, "the reference's keys"
    m2.load_connectors(tmp_path)
    # a trainer checkpoint: {"model_state_dict": {... "audio_connector.linear.weight": ...}} in fp16 (use_fp16 run)
    ckpt = {"model_state_dict": {k: v.half() for k, v in m1.state_dict().items()}}
    torch.save(ckpt, tmp_path / "checkpoint.pt")
    sd = torch.load(tmp_path / "checkpoint.pt")["model_state_dict"]
    for prefix in ("audio_connector.", "video_connector."):
        part = {k[len(prefix):]: v.float() for k, v in sd.items() if k.startswith(prefix)}
        getattr(m3, prefix[:-1]).load_state_dict(part)
    audio = torch.randn(2, 80, 12, device=cuda_dev)
    with torch.no_grad():
        e1, e2, e3 = (m.eval().encode(audio=audio)[0] for m in (m1, m2, m3))
    for n in ("weight", "bias"):
        assert torch.equal(getattr(m1.audio_connector.linear, n), getattr(m2.audio_connector.linear, n))
    # m1 / m2 share the LLM-independent connector: identical bits; m3 went through an fp16 checkpoint
    w1 = m1.whisper.proj(audio.transpose(1, 2))
    with torch.no_grad():
        y1, y2, y3 = (m.audio_connector(w1) for m in (m1, m2, m3))
    assert torch.equal(y1, y2)
    assert rel_err(y3, y1) <= 1e-2  # fp16-rounded weights re-rounded to bf16 operands


def test_adaptive_connector_runs_on_the_projector_gemm(avc, cuda_dev):
    """The reference's `adaptive` type (modality_connector.py:239-299): same state-dict keys; its two dense projections
    on the tcgen05 GEMM incl. dX (the gradient reaches input_proj through output_proj's input)."""
    import torch.nn.functional as F

    torch.manual_seed(0)
    conn = avc.create_modality_connector("adaptive", 64, 128, device="cuda:0", max_seq_len=600).eval()
    keys = set(conn.state_dict())
    for k in ("input_proj.weight", "norm1.weight", "pos_encoder.pe", "adaptive_pool.long_adapter.0.weight",
              "adaptive_pool.attn.in_proj_weight", "adaptive_pool.norm.bias", "output_proj.bias", "norm2.weight"):
        assert k in keys, k
    x = torch.randn(2, 40, 64, device=cuda_dev)
    up = torch.randn(2, 40, 128, device=cuda_dev)
    y = conn(x)
    (y * up).sum().backward()
    g_in, g_out = conn.input_proj.weight.grad.clone(), conn.output_proj.weight.grad.clone()

    def eager(xx):
        h = conn.act(conn.norm1(F.linear(xx, conn.input_proj.weight, conn.input_proj.bias)))
        h = conn.adaptive_pool(conn.pos_encoder(h))
        return conn.norm2(F.linear(h, conn.output_proj.weight, conn.output_proj.bias))

    conn.zero_grad()
    ref = eager(x)
    (ref * up).sum().backward()
    assert_close(y, ref, "adaptive forward")
    assert_close(g_out, conn.output_proj.weight.grad, "output_proj dW", rel=2e-2)
    assert_close(g_in, conn.input_proj.weight.grad, "input_proj dW (through dX)", rel=3e-2)
    long = conn(torch.randn(1, 600, 64, device=cuda_dev))
    assert long.shape == (1, 150, 128), "> 512 frames: two stride-2 convolutions (modality_connector.py:359-366)"


def test_public_api_gradient_bucket_world1(avc, cuda_dev):
    """parallel.FusedGradSync without a process group: the backward writes dW / db into the peer bucket the
    parameters' .grad alias and runs the fused (world = 1) all-reduce launch; same gradients as the plain path."""
    from audio_visual_llm_b200.parallel import FusedGradSync

    g = torch.Generator().manual_seed(9)
    B, Ta, Tv, Da, Dv, H = 2, 32, 16, 64, 64, 256
    a, v = torch.randn(B, Ta, Da, generator=g).to(cuda_dev), torch.randn(B, Tv, Dv, generator=g).to(cuda_dev)
    base = _rand_params(g, H, 2 * Da, Dv)
    plan = avc.FusePlan(fusion="concat", audio_stride=2, video_stride=1, max_seq_len=64)
    up = torch.randn(B, 16, H, generator=g).to(cuda_dev)
    plain = [t.to(cuda_dev).requires_grad_(True) for t in base]
    emb, _, _ = avc.fused_connector(a, v, *plain, plan, out_dtype=torch.bfloat16)
    (emb.float() * up).sum().backward()
    synced = [t.to(cuda_dev).requires_grad_(True) for t in base]
    sync = FusedGradSync(*synced, multimem=False)
    try:
        assert sync.fused and all(p.grad is not None for p in synced)
        for _ in range(2):  # epochs: the flags are never reset
            sync.bucket.flat.fill_(float("nan"))
            emb2, _, _ = avc.fused_connector(a, v, *synced, plan, out_dtype=torch.bfloat16, grad_sync=sync)
            (emb2.float() * up).sum().backward()
            torch.cuda.synchronize()
            sync.check()
            for p, q, view, n in zip(synced, plain, sync.views(True, True), ["dWa", "dba", "dWv", "dbv"]):
                assert p.grad.data_ptr() == view.data_ptr(), "the parameter's .grad IS the bucket view"
                assert rel_err(p.grad, q.grad) <= 2e-5, n
    finally:
        sync.close()


# ------------------------------------------------------------------------------------------ what bench.py times
@pytest.mark.parametrize("config", ["cfg2", "cfg1", "cfg4", "cfg2p", "cfg3", "cfg3k4"])
def test_engine_at_full_size_passes_the_bench_self_check(avc, cuda_dev, config):
    """engine.ConnectorStep at the FULL BASELINE shapes against the fp64 recomputation bench.py runs after its timed
    region: 64 sampled output rows, 32 x 32 sampled dW entries per stream, the whole db."""
    import bench

    eng, _, _ = bench.make_engine(avc, bench.CONFIGS[config], cuda_dev, 1234)
    res = bench.self_check(torch, None, eng, 1, cuda_dev)
    assert res["ok"], res
    assert res["output_max_rel"] <= 1e-2 and res["output_cosine"] >= 0.9999
    assert res["dw_max_rel"] <= 1e-3 and res["db_max_rel"] <= 1e-3
    assert int(eng.status.item()) == 0
    assert eng.direct == (config in ("cfg2", "cfg3", "cfg3k4"))   # dense streams whose frames divide by the stride
    if config == "cfg4":
        assert eng.ragged and eng.M == sum(eng.counts) and min(eng.counts) >= 100 and max(eng.counts) <= 400


def test_graphed_encoder_replays_the_eager_bits(avc, cuda_dev):
    """Forward-only fast path: one CUDA-graph launch per encode (decode / generate); new inputs, same graph."""
    from audio_visual_llm_b200.engine import GraphedEncoder

    g = torch.Generator().manual_seed(61)
    B, Ta, Tv, Da, Dv, H, P, V = 1, 40, 20, 64, 32, 128, 4, 50
    wa, ba, wv, bv = (t.to(cuda_dev) for t in _rand_params(g, H, 4 * Da, 2 * Dv))
    table = torch.randn(V, H, generator=g).to(cuda_dev, torch.bfloat16)
    plan = avc.FusePlan(fusion="concat", audio_stride=4, video_stride=2, max_seq_len=64)

    def enc(a, v, ids):
        return avc.fused_connector(a, v, wa, ba, wv, bv, plan, prompt_ids=ids, embed_table=table,
                                   out_dtype=torch.bfloat16)

    graphed = GraphedEncoder(enc)
    for trial in range(3):
        a = torch.randn(B, Ta, Da, generator=g).to(cuda_dev, torch.bfloat16)
        v = torch.randn(B, Tv, Dv, generator=g).to(cuda_dev, torch.bfloat16)
        ids = torch.randint(1, V, (B, P), generator=g).to(cuda_dev)
        with torch.no_grad():
            emb, mask, _ = enc(a, v, ids)
        gemb, gmask, _ = graphed(a, v, ids)
        torch.cuda.synchronize()
        assert torch.equal(gemb, emb) and torch.equal(gmask, mask), trial
    assert len(graphed._graphs) == 1, "one capture serves every call with the same input signature"


def test_adaptive_connector_matches_the_executing_reference(avc, cuda_dev):
    """tests/golden/adaptive_connector.npz: the reference's own `adaptive` connector (its factory, eval mode) -- weights,
    inputs, outputs and the gradients of its two dense projections.  Our module must load that state dict as is and
    reproduce the outputs (bf16 tensor-core operands, fp32 accumulate: REL_TOL) for a short input and for one longer
    than 512 frames (the two stride-2 convolutions)."""
    import numpy as np
    from pathlib import Path

    z = np.load(Path(__file__).resolve().parent / "golden" / "adaptive_connector.npz")
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    conn = avc.create_modality_connector("adaptive", 32, 48, device="cuda:0", max_seq_len=640).eval()
    missing = conn.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x = torch.from_numpy(z["in.x"]).to(cuda_dev)
    up = torch.from_numpy(z["in.upstream"]).to(cuda_dev)
    y = conn(x)
    (y * up).sum().backward()
    torch.cuda.synchronize()
    assert_close(y, torch.from_numpy(z["out.y"]), "adaptive forward vs reference", rel=2e-2)
    for name in ("input_proj", "output_proj"):
        lin = getattr(conn, name)
        assert_close(lin.weight.grad, torch.from_numpy(z[f"out.{name}.weight.grad"]), f"{name} dW vs reference", rel=3e-2)
        assert_close(lin.bias.grad, torch.from_numpy(z[f"out.{name}.bias.grad"]), f"{name} db vs reference", rel=3e-2)
    with torch.no_grad():
        y_long = conn(torch.from_numpy(z["in.x_long"]).to(cuda_dev))
    assert_close(y_long, torch.from_numpy(z["out.y_long"]), "adaptive forward, 520 frames", rel=2e-2)
