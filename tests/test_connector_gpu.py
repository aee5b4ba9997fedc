"""Model-level parity on a B200: the drop-in ClipWhisperModel / ModalityConnector against
(1) outputs of the executing reference (tests/golden/ref_*.npz) and (2) the CPU oracle for the extensions.

Tolerances (north_star): indices / masks / labels / copied text rows bit-exact; projected embeddings and
gradients max-relative error <= 1e-2 and cosine >= 0.9999 against the fp32 reference (bf16 operands, fp32
accumulate)."""
import ast
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import connector_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = sorted(p.stem[len("ref_"):] for p in GOLDEN.glob("ref_*.npz") if "cfg1" not in p.stem)
REL_TOL, COS_TOL = 1e-2, 0.9999


def rel_err(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def cosine(got, ref):
    got, ref = got.detach().double().cpu().flatten(), ref.detach().double().cpu().flatten()
    if float(got.norm()) == 0.0 and float(ref.norm()) == 0.0:
        return 1.0   # two exactly-zero tensors (e.g. the video gradients at fusion_scale = 1) agree
    return float(torch.dot(got, ref) / (got.norm() * ref.norm()).clamp_min(1e-30))


def assert_close(got, ref, what):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    r, c = rel_err(got, ref), cosine(got, ref)
    assert r <= REL_TOL and c >= COS_TOL, f"{what}: max-rel {r:.3e}, cosine {c:.6f}"


class StubWhisper(nn.Module):
    def __init__(self, d_model):
        super().__init__()
        self.dummy = nn.Parameter(torch.zeros(1))
        self.config = SimpleNamespace(d_model=d_model)
        self.features = None
        outer = self

        class Enc(nn.Module):
            def forward(self, audio, attention_mask=None, output_hidden_states=True, return_dict=True):
                return SimpleNamespace(last_hidden_state=outer.features)

        self.encoder = Enc()


class StubClip(nn.Module):
    def __init__(self, hidden_size):
        super().__init__()
        self.dummy = nn.Parameter(torch.zeros(1))
        self.config = SimpleNamespace(hidden_size=hidden_size)
        self.hidden = None

    def forward(self, flat, return_dict=True):
        return SimpleNamespace(last_hidden_state=self.hidden)


class StubLLM(nn.Module):
    def __init__(self, table):
        super().__init__()
        self.embed = nn.Embedding.from_pretrained(table.clone(), freeze=True)
        self.calls = []
        self.upstream = None

    def get_input_embeddings(self):
        return self.embed

    def forward(self, inputs_embeds=None, attention_mask=None, labels=None, return_dict=True):
        self.calls.append(dict(inputs_embeds=inputs_embeds, attention_mask=attention_mask, labels=labels))
        loss = (inputs_embeds.float() * self.upstream.to(inputs_embeds.device)).sum()
        return SimpleNamespace(loss=loss, logits=inputs_embeds)


def build_model(pkg, dev, d, cfg, llm_dtype, Da=32, Dv=16, **kw):
    table = d["in.embed_table"].to(dev, llm_dtype)
    m = pkg.ClipWhisperModel(
        device="cuda:0", modality=cfg["modality"], max_seq_len=cfg["max_seq_len"], fusion_scale=cfg["fs"],
        _provided_tokenizer=SimpleNamespace(pad_token_id=0), _provided_llm=StubLLM(table).to(dev),
        _provided_whisper=StubWhisper(Da).to(dev), _provided_clip=StubClip(Dv).to(dev), **kw)
    sd = {k[len("in."):]: v.to(dev) for k, v in d.items() if k.startswith("in.") and "connector.linear" in k}
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    return m


def load(name):
    z = np.load(GOLDEN / f"ref_{name}.npz")
    d = {k: torch.from_numpy(z[k]) for k in z.files if k != "cfg"}
    return d, ast.literal_eval(str(z["cfg"]))


def drive(m, d, cfg, dev, train):
    audio = video = None
    if "in.audio_feats" in d:
        m.whisper.features = d["in.audio_feats"].to(dev)
        audio = torch.zeros(2, 80, 4, device=dev)
    if "in.clip_hidden" in d:
        m.clip.hidden = d["in.clip_hidden"].to(dev)
        video = torch.zeros(2, cfg["Tv"], 3, 2, 2, device=dev)
    m.llm.upstream = d["in.upstream"]
    m.train(train)
    prompt = d["in.prompt"].to(dev) if "in.prompt" in d else None
    res = m(audio=audio, video=video, prompt=prompt, labels=d["in.labels"].to(dev))
    res["loss"].backward()
    torch.cuda.synchronize()
    return m.llm.calls[-1]


@pytest.mark.parametrize("llm_dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("name", CASES)
def test_model_forward_backward_matches_executing_reference(avc, cuda_dev, name, llm_dtype):
    d, cfg = load(name)
    m = build_model(avc, cuda_dev, d, cfg, llm_dtype)
    rec = drive(m, d, cfg, cuda_dev, cfg["train"])
    emb, ref = rec["inputs_embeds"], d["out.inputs_embeds"]
    assert emb.dtype == llm_dtype and emb.shape == ref.shape
    assert_close(emb, ref, "inputs_embeds")
    P = min(d["in.prompt"].shape[1], 32) if "in.prompt" in d else 0
    if P and not cfg["train"]:  # copied text rows are exact (in the LLM dtype)
        assert torch.equal(emb[:, :P].cpu(), ref[:, :P].to(llm_dtype))
    assert rec["attention_mask"].dtype == torch.int64 and torch.equal(rec["attention_mask"].cpu(), d["out.attention_mask"])
    assert rec["labels"].dtype == torch.int64 and torch.equal(rec["labels"].cpu(), d["out.labels"])
    for n in ("audio_connector", "video_connector"):
        for p in ("weight", "bias"):
            key = f"out.{n}.linear.{p}.grad"
            g = getattr(getattr(m, n).linear, p).grad
            if key in d:
                assert g is not None and g.dtype == torch.float32
                assert_close(g, d[key], key)
            else:
                assert g is None or float(g.abs().max()) == 0.0


def test_state_dict_keys_and_shapes_match_reference(avc, cuda_dev):
    d, cfg = load("both_prompt_eval")
    m = build_model(avc, cuda_dev, d, cfg, torch.float32)
    names = dict(m.named_parameters())
    for n, shape in [("audio_connector.linear.weight", (48, 32)), ("audio_connector.linear.bias", (48,)),
                     ("video_connector.linear.weight", (48, 16)), ("video_connector.linear.bias", (48,))]:
        assert tuple(names[n].shape) == shape and names[n].dtype == torch.float32
    assert set(m.audio_connector.state_dict()) == {"linear.weight", "linear.bias"}


def test_modality_connector_module_fwd_bwd(avc, cuda_dev):
    """ModalityConnector(input_dim, output_dim, device=)(x) as decode.py:211-221 / encode_audio use it."""
    g = torch.Generator().manual_seed(5)
    conn = avc.ModalityConnector(input_dim=64, output_dim=96, device="cuda:0")
    assert avc.create_modality_connector("simple", 64, 96, device="cuda:0", max_seq_len=7).linear.weight.shape == (96, 64)
    with torch.no_grad():
        conn.linear.bias.copy_(torch.randn(96, generator=g))
    x = torch.randn(3, 50, 64, generator=g)
    up = torch.randn(3, 50, 96, generator=g)
    w, b = conn.linear.weight.detach().cpu().clone().requires_grad_(True), conn.linear.bias.detach().cpu().clone().requires_grad_(True)
    ref = O.reference_connector(x, w, b)
    (ref * up).sum().backward()
    y = conn(x.to(cuda_dev))
    assert y.dtype == torch.float32
    (y * up.to(cuda_dev)).sum().backward()
    torch.cuda.synchronize()
    assert_close(y, ref, "y")
    assert_close(conn.linear.weight.grad, w.grad, "dW")
    assert_close(conn.linear.bias.grad, b.grad, "db")
    with pytest.raises(NotImplementedError):
        avc.create_modality_connector("qformer", 64, 96)


def _rand_params(g, H, Ka, Kv):
    return (torch.randn(H, Ka, generator=g) / Ka ** 0.5, torch.randn(H, generator=g) * 0.1,
            torch.randn(H, Kv, generator=g) / Kv ** 0.5, torch.randn(H, generator=g) * 0.1)


@pytest.mark.parametrize("fusion", ["concat", "sum"])
def test_stride_stack_rate_aligned_vs_oracle(avc, cuda_dev, fusion):
    """cfg2-style: k_a=4 audio + k_v=2 video frames per token (rate-aligned), ragged tails, prompt."""
    g = torch.Generator().manual_seed(31)
    B, Ta, Tv, Da, Dv, H, P, V = 3, 50, 23, 64, 32, 128, 6, 40
    a, v = torch.randn(B, Ta, Da, generator=g), torch.randn(B, Tv, Dv, generator=g)
    wa, ba, wv, bv = _rand_params(g, H, 4 * Da, 2 * Dv)
    table = torch.randn(V, H, generator=g)
    prompt = torch.randint(1, V, (B, P), generator=g)
    labels = torch.randint(0, V, (B, 9), generator=g)
    spec = O.ConnectorSpec(fusion=fusion, fusion_scale=0.3, audio_stride=4, video_stride=2, max_seq_len=64)
    emb_r, mask_r, lab_r, flags_r = O.connector_forward(a, v, wa, ba, wv, bv, spec, prompt_ids=prompt,
                                                        embed_table=table, labels=labels)
    N = emb_r.shape[1] - P
    assert N == 13
    up = torch.randn(B, P + N, H, generator=g)
    grads_r = O.connector_grads(a, v, wa, ba, wv, bv, spec, up[:, P:])
    dev = cuda_dev
    params = [t.to(dev).requires_grad_(True) for t in (wa, ba, wv, bv)]
    plan = avc.FusePlan(fusion=fusion, fusion_scale=0.3, audio_stride=4, video_stride=2, max_seq_len=64)
    emb, mask, lab = avc.fused_connector(a.to(dev), v.to(dev), *params, plan, prompt_ids=prompt.to(dev),
                                         embed_table=table.to(dev, torch.bfloat16), labels=labels.to(dev),
                                         out_dtype=torch.bfloat16, check=True)
    (emb.float() * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    assert_close(emb[:, P:], emb_r[:, P:], "AV rows")
    assert torch.equal(emb[:, :P].cpu(), emb_r[:, :P].to(torch.bfloat16))
    assert torch.equal(mask.cpu(), mask_r) and torch.equal(lab.cpu(), lab_r)
    for p, gr, n in zip(params, grads_r, ["dWa", "dba", "dWv", "dbv"]):
        assert_close(p.grad, gr, n)


def test_ragged_video_only_placeholders_vs_oracle(avc, cuda_dev):
    """cfg4-style: video-only, per-sample valid lengths, placeholders at ragged offsets, real masks."""
    g = torch.Generator().manual_seed(41)
    B, Tv, Dv, H, V, S, PH = 4, 40, 64, 128, 60, 58, 59
    lens = [40, 17, 1, 29]
    v = torch.randn(B, Tv, Dv, generator=g)
    wa, ba, wv, bv = _rand_params(g, H, 8, Dv)
    table = torch.randn(V, H, generator=g)
    ids = torch.zeros(B, S, dtype=torch.int64)
    for b, n in enumerate(lens):
        row = torch.randint(1, PH, (S,), generator=g)
        start = 2 + 3 * b
        row[start:start + n] = PH
        row[start + n + 4:] = 0
        ids[b] = row
    spec = O.ConnectorSpec(modality="video", mask_mode=1, label_mode=1)
    tok_r, _ = O.connector_tokens(None, v, wa, ba, wv, bv, spec, video_valid=torch.tensor(lens))
    emb_r, mask_r, lab_r = O.splice_tokens(tok_r, ids, PH, table, 0, spec, ntok=torch.tensor(lens))
    dev = cuda_dev
    wv_d, bv_d = wv.to(dev).requires_grad_(True), bv.to(dev).requires_grad_(True)
    plan = avc.FusePlan(modality="video", mask_mode=1, label_mode=1)
    emb, mask, lab = avc.fused_connector(None, v.to(dev), None, None, wv_d, bv_d, plan, input_ids=ids.to(dev),
                                         placeholder_id=PH, embed_table=table.to(dev, torch.bfloat16),
                                         out_dtype=torch.bfloat16, video_lengths=lens, check=True)
    up = torch.randn(B, S, H, generator=g)
    (emb.float() * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    is_ph = ids == PH
    assert_close(emb[is_ph.to(dev)], emb_r[is_ph], "AV rows")
    assert torch.equal(emb.cpu()[~is_ph], emb_r[~is_ph].to(torch.bfloat16))
    assert torch.equal(mask.cpu(), mask_r) and torch.equal(lab.cpu(), lab_r)
    # oracle gradients: upstream restricted to the placeholder rows, per sample
    wv_c, bv_c = wv.clone().requires_grad_(True), bv.clone().requires_grad_(True)
    tok, _ = O.connector_tokens(None, v, wa, ba, wv_c, bv_c, spec, video_valid=torch.tensor(lens))
    loss = sum((tok[b, :n] * up[b][is_ph[b]]).sum() for b, n in enumerate(lens))
    loss.backward()
    assert_close(wv_d.grad, wv_c.grad, "dWv")
    assert_close(bv_d.grad, bv_c.grad, "dbv")


def test_adaptive_projection_fwd_bwd(avc, cuda_dev):
    """Train-time length adaptation kernels vs the oracle restatement of _adaptive_projection."""
    g = torch.Generator().manual_seed(51)
    for S, Lt in [(28, 10), (8, 19), (1500, 256)]:
        x = torch.randn(2, S, 64, generator=g)
        up = torch.randn(2, Lt, 64, generator=g)
        xc = x.clone().requires_grad_(True)
        ref = O.reference_adaptive_projection(xc, Lt)
        (ref * up).sum().backward()
        xd = x.to(cuda_dev).requires_grad_(True)
        out = avc.adaptive_projection(xd, Lt)
        (out * up.to(cuda_dev)).sum().backward()
        torch.cuda.synchronize()
        # fp32 in, fp32 accumulate: only the order of the window sum differs
        assert torch.allclose(out.cpu(), ref.detach(), rtol=1e-5, atol=1e-5)
        assert torch.allclose(xd.grad.cpu(), xc.grad, rtol=1e-5, atol=1e-5)


def test_no_cpu_fallback(avc, cuda_dev):
    plan = avc.FusePlan(modality="audio")
    w, b = torch.zeros(16, 8), torch.zeros(16)
    with pytest.raises(avc._lib.ConnectorError):
        avc.fused_connector(torch.zeros(1, 4, 8), None, w, b, None, None, plan)
    with pytest.raises(avc._lib.ConnectorError):
        avc.ClipWhisperModel(device="cpu", _provided_llm=object(), _provided_tokenizer=object())


@pytest.mark.parametrize("modality,ka,kv", [("both", 4, 2), ("audio", 2, 1), ("video", 1, 1), ("both", 1, 1)])
def test_step_engine_fused_equals_unfused_and_oracle(avc, cuda_dev, modality, ka, kv):
    """engine.ConnectorStep (what bench.py times): the gather-free step (GEMM reads the tower outputs and
    d(inputs_embeds) in place) against the gather / splice-bwd step and against the CPU oracle."""
    from audio_visual_llm_b200.engine import ConnectorStep, StepShape

    shape = StepShape(batch=3, audio_frames=40 * ka, video_frames=40 * kv, audio_dim=64, video_dim=32, hidden=128,
                      prompt_len=5, vocab=50, label_len=30)
    plan = avc.FusePlan(modality=modality, fusion="concat", audio_stride=ka, video_stride=kv, max_seq_len=4096)
    fused = ConnectorStep(shape, plan, cuda_dev, seed=3, fuse_gather=True)
    plain = ConnectorStep(shape, plan, cuda_dev, seed=3, fuse_gather=False)
    assert fused.direct and not plain.direct
    outs = []
    for eng in (fused, plain):
        emb, mask, lab = eng.forward()
        g = eng.backward(allreduce=False)
        torch.cuda.synchronize()
        assert int(eng.status.item()) == 0
        outs.append((emb.clone(), mask.clone(), lab.clone(), {k: v.clone() for k, v in g.views.items()}))
    (e1, m1, l1, g1), (e2, m2, l2, g2) = outs
    assert torch.equal(e1, e2) and torch.equal(m1, m2) and torch.equal(l1, l2)  # same k-block order: same bits
    for k in g1:
        assert rel_err(g1[k], g2[k]) <= 1e-5, k  # reduction split per sample vs packed: fp32 reassociation only
    # oracle on the same (bf16-rounded) inputs
    eng = fused
    f32 = lambda t: None if t is None else t.detach().float().cpu()
    spec = O.ConnectorSpec(modality=modality, fusion="concat", audio_stride=ka, video_stride=kv, max_seq_len=4096)
    wa = f32(eng.wa) if eng.use_a else torch.zeros(128, 8)
    wv = f32(eng.wv) if eng.use_v else torch.zeros(128, 8)
    ba = f32(eng.ba) if eng.use_a else torch.zeros(128)
    bv = f32(eng.bv) if eng.use_v else torch.zeros(128)
    ids = eng.input_ids.cpu()
    tok, _ = O.connector_tokens(f32(eng.audio), f32(eng.video), wa, ba, wv, bv, spec)
    emb_r, mask_r, lab_r = O.splice_tokens(tok, ids, eng.placeholder_id, f32(eng.embed_table), 0, spec,
                                           labels=eng.labels_in.cpu())
    P = shape.prompt_len
    assert_close(e1[:, P:], emb_r[:, P:], "AV rows")
    assert torch.equal(e1[:, :P].cpu(), emb_r[:, :P].to(torch.bfloat16))
    assert torch.equal(m1.cpu(), mask_r) and torch.equal(l1.cpu(), lab_r)
    grads = O.connector_grads(f32(eng.audio), f32(eng.video), wa, ba, wv, bv, spec, f32(eng.d_emb)[:, P:])
    names = ["audio_connector.linear.weight", "audio_connector.linear.bias", "video_connector.linear.weight",
             "video_connector.linear.bias"]
    for n, gr in zip(names, grads):
        if n in g1:
            assert_close(g1[n], gr, n)


def _mlp_params(g, Hd, K, H):
    return [torch.randn(Hd, K, generator=g) / K ** 0.5, torch.randn(Hd, generator=g) * 0.1,
            torch.randn(H, Hd, generator=g) / Hd ** 0.5, torch.randn(H, generator=g) * 0.1]


@pytest.mark.parametrize("case", ["direct_concat", "padded_sum", "ragged_video"])
def test_mlp_projector_fwd_bwd_vs_oracle(avc, cuda_dev, case):
    """Linear -> GELU -> Linear projector through the fused connector: forward and all eight parameter gradients."""
    g = torch.Generator().manual_seed(61)
    B, Da, Dv, Hd, H, P, V = 2, 32, 16, 96, 64, 3, 30
    dev = cuda_dev
    kwargs = {}
    if case == "direct_concat":  # both streams dense, stride 4 / 2 -> free views, no gather
        Ta, Tv, ka, kv, fusion = 40, 20, 4, 2, "concat"
        a, v = torch.randn(B, Ta, Da, generator=g), torch.randn(B, Tv, Dv, generator=g)
        av = vv = None
    elif case == "padded_sum":   # reference-style: k=1, video shorter -> padded rows must get no video output at all
        Ta, Tv, ka, kv, fusion = 24, 9, 1, 1, "sum"
        a, v = torch.randn(B, Ta, Da, generator=g), torch.randn(B, Tv, Dv, generator=g)
        av = vv = None
    else:
        Ta, Tv, ka, kv, fusion = 8, 20, 1, 1, "sum"
        a, v = None, torch.randn(B, Tv, Dv, generator=g)
        av, vv = None, torch.tensor([20, 7])
    pa, pv = _mlp_params(g, Hd, ka * Da, H), _mlp_params(g, Hd, kv * Dv, H)
    modality = "video" if a is None else "both"
    spec = O.ConnectorSpec(modality=modality, fusion=fusion, fusion_scale=0.3, audio_stride=ka, video_stride=kv,
                           max_seq_len=64, mask_mode=1 if vv is not None else 0)
    pa_c = [t.clone().requires_grad_(True) for t in pa]
    pv_c = [t.clone().requires_grad_(True) for t in pv]
    tok = O.connector_tokens_mlp(a, v, pa_c, pv_c, spec, video_valid=vv)
    N = tok.shape[1]
    table = torch.randn(V, H, generator=g)
    if vv is None:
        prompt = torch.randint(1, V - 1, (B, P), generator=g)
        ids = torch.cat([prompt, torch.full((B, N), V - 1)], 1)
        ntok = None
    else:
        lens = vv.tolist()
        ids = torch.zeros(B, P + N, dtype=torch.int64)
        for b, n in enumerate(lens):
            row = torch.randint(1, V - 1, (P + N,), generator=g)
            row[1:1 + n] = V - 1
            row[1 + n + 2:] = 0
            ids[b] = row
        ntok = vv
    emb_r, mask_r, _ = O.splice_tokens(tok, ids, V - 1, table, 0, spec, ntok=ntok)
    up = torch.randn(emb_r.shape, generator=g)
    (emb_r * up).sum().backward()
    pa_d = [t.to(dev).requires_grad_(True) for t in pa]
    pv_d = [t.to(dev).requires_grad_(True) for t in pv]
    plan = avc.FusePlan(modality=modality, fusion=fusion, fusion_scale=0.3, audio_stride=ka, video_stride=kv,
                        max_seq_len=64, mask_mode=1 if vv is not None else 0)
    common = dict(embed_table=table.to(dev, torch.bfloat16), out_dtype=torch.bfloat16, check=True,
                  mlp_audio=pa_d if a is not None else None, mlp_video=pv_d)
    if vv is None:
        emb, mask, _ = avc.fused_connector(a.to(dev), v.to(dev), None, None, None, None, plan,
                                           prompt_ids=ids[:, :P].to(dev), placeholder_id=V - 1, **common)
    else:
        emb, mask, _ = avc.fused_connector(None, v.to(dev), None, None, None, None, plan, input_ids=ids.to(dev),
                                           placeholder_id=V - 1, video_lengths=vv.tolist(), **common)
    (emb.float() * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    is_ph = ids == V - 1
    assert_close(emb[is_ph.to(dev)], emb_r[is_ph], "AV rows")
    assert torch.equal(emb.cpu()[~is_ph], emb_r[~is_ph].to(torch.bfloat16))
    assert torch.equal(mask.cpu(), mask_r)
    names = ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]
    for tag, dev_p, cpu_p in (("audio", pa_d, pa_c), ("video", pv_d, pv_c)):
        if tag == "audio" and a is None:
            continue
        for n, pd, pc in zip(names, dev_p, cpu_p):
            assert_close(pd.grad, pc.grad, f"{tag}.{n}")
    # forward-only (no grad): GELU runs in the GEMM epilogue when every token is present
    with torch.no_grad():
        if vv is None:
            emb2, _, _ = avc.fused_connector(a.to(dev), v.to(dev), None, None, None, None, plan,
                                             prompt_ids=ids[:, :P].to(dev), placeholder_id=V - 1, **common)
            assert_close(emb2[is_ph.to(dev)], emb_r[is_ph], "AV rows (epilogue GELU)")


def test_mlp_modality_connector_module(avc, cuda_dev):
    g = torch.Generator().manual_seed(62)
    conn = avc.create_modality_connector("mlp", 64, 96, device="cuda:0", hidden_dim=128)
    assert set(conn.state_dict()) == {"fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"}
    with torch.no_grad():
        conn.fc1.bias.copy_(torch.randn(128, generator=g) * 0.1)
        conn.fc2.bias.copy_(torch.randn(96, generator=g) * 0.1)
    x = torch.randn(3, 40, 64, generator=g)
    p = [t.detach().cpu().clone().requires_grad_(True) for t in conn.mlp_params()]
    ref = torch.nn.functional.gelu(x @ p[0].t() + p[1]) @ p[2].t() + p[3]
    up = torch.randn(ref.shape, generator=g)
    (ref * up).sum().backward()
    y = conn(x.to(cuda_dev))
    (y * up.to(cuda_dev)).sum().backward()
    torch.cuda.synchronize()
    assert y.dtype == torch.float32
    assert_close(y, ref, "y")
    for t, r, n in zip(conn.mlp_params(), p, ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]):
        assert_close(t.grad, r.grad, n)


def test_forward_only_reuses_packed_weights_until_they_change(avc, cuda_dev):
    """Inference fast path: the bf16 weight pack is cached across no-grad calls and invalidated by an in-place
    parameter update (torch version counter)."""
    from audio_visual_llm_b200 import connector_ops as C

    g = torch.Generator().manual_seed(71)
    conn = avc.ModalityConnector(64, 96, device="cuda:0")
    x = torch.randn(1, 30, 64, generator=g).to(cuda_dev)
    C._PACK_CACHE.clear()
    with torch.no_grad():
        y1 = conn(x)
        n1 = len(C._PACK_CACHE)
        y2 = conn(x)
        assert len(C._PACK_CACHE) == n1 == 1 and torch.equal(y1, y2)
        conn.linear.weight.mul_(2.0)  # in-place update bumps the version counter -> re-pack
        y3 = conn(x)
        assert len(C._PACK_CACHE) == 1 and not torch.equal(y3, y1)
    ref = x.cpu() @ conn.linear.weight.detach().cpu().t() + conn.linear.bias.detach().cpu()
    assert_close(y3, ref, "y after weight update")
    y4 = conn(x)  # grad mode with trainable weights: never cached
    assert len(C._PACK_CACHE) == 1 and y4.requires_grad


def test_cfg1_on_gpu_matches_executing_reference_subsample(avc, cuda_dev):
    """BASELINE.json configs[0] (Whisper-small 768 + CLIP 512 -> 2048, batch 2, 10 s, parity mode) on the B200
    against the sub-sampled outputs the executing reference produced for the same seeded inputs."""
    z = np.load(GOLDEN / "ref_cfg1_subsampled.npz")
    Bc, Ta, Tv, Da, Dv, H = 2, 500, 250, 768, 512, 2048
    g = torch.Generator().manual_seed(1234 + 1)
    a = torch.randn(Bc, Ta, Da, generator=g)
    v = torch.randn(Bc, Tv, Dv, generator=g)
    up = torch.randn(Bc, Ta, H, generator=torch.Generator().manual_seed(77))
    chk = np.array([a.double().sum().item(), v.double().sum().item(), up.double().sum().item()])
    if not np.allclose(chk, z["in.checksum"], rtol=0, atol=1e-6):
        pytest.skip("torch RNG stream differs from the one the fixture was generated with")
    torch.manual_seed(0)
    la = torch.nn.Linear(Da, H)
    torch.nn.init.xavier_uniform_(la.weight)
    lv = torch.nn.Linear(Dv, H)
    torch.nn.init.xavier_uniform_(lv.weight)
    gb = torch.Generator().manual_seed(1)
    ba, bv = torch.randn(H, generator=gb) * 0.02, torch.randn(H, generator=gb) * 0.02
    if not np.isclose(la.weight.double().sum().item(), float(z["w.checksum.audio_connector"]), atol=1e-6):
        pytest.skip("weight init RNG stream differs from the fixture's")
    dev = cuda_dev
    params = [t.detach().to(dev).requires_grad_(True) for t in (la.weight, ba, lv.weight, bv)]
    plan = avc.FusePlan(max_seq_len=512, fusion_scale=0.5)
    for out_dtype in (torch.float32, torch.bfloat16):
        for p in params:
            p.grad = None
        emb, mask, _ = avc.fused_connector(a.to(dev), v.to(dev), *params, plan, out_dtype=out_dtype)
        emb.backward(up.to(dev, out_dtype))
        torch.cuda.synchronize()
        assert_close(emb[:, ::25, ::64], torch.from_numpy(z["out.inputs_embeds[::25, ::64]"]), "inputs_embeds sample")
        assert int(mask.sum()) == int(z["out.attention_mask.sum"])
        assert_close(params[0].grad[::32, ::32], torch.from_numpy(z["out.audio_connector.linear.weight.grad[::32, ::32]"]), "dWa")
        assert_close(params[2].grad[::32, ::32], torch.from_numpy(z["out.video_connector.linear.weight.grad[::32, ::32]"]), "dWv")
        assert_close(params[1].grad, torch.from_numpy(z["out.audio_connector.linear.bias.grad"]), "dba")
        assert_close(params[3].grad, torch.from_numpy(z["out.video_connector.linear.bias.grad"]), "dbv")


def test_gradients_are_deterministic(avc, cuda_dev):
    """No split-K atomics anywhere: two runs of the same step give bit-identical dW / db (fused and unfused)."""
    from audio_visual_llm_b200.engine import ConnectorStep, StepShape

    shape = StepShape(batch=4, audio_frames=400, video_frames=200, audio_dim=128, video_dim=64, hidden=256,
                      prompt_len=8, vocab=100)
    plan = avc.FusePlan(fusion="concat", audio_stride=4, video_stride=2, max_seq_len=4096)
    for fuse in (True, False):
        eng = ConnectorStep(shape, plan, cuda_dev, seed=1, fuse_gather=fuse)
        eng.step(allreduce=False)
        torch.cuda.synchronize()
        first = eng.bucket.flat.clone()
        emb1 = eng.emb.clone()
        eng.bucket.flat.zero_()
        eng.step(allreduce=False)
        torch.cuda.synchronize()
        assert torch.equal(eng.bucket.flat, first) and torch.equal(eng.emb, emb1)


@pytest.mark.parametrize("case", ["zero_length_sample", "single_token", "frames_fewer_than_stride", "odd_widths"])
def test_edge_cases_vs_oracle(avc, cuda_dev, case):
    """Empty / degenerate inputs: a sample with no valid frames, one token, fewer frames than the stride, widths that
    are not multiples of the GEMM tile (H = 72, D = 40)."""
    g = torch.Generator().manual_seed(81)
    dev = cuda_dev
    H, V, PH = 72, 40, 39
    if case == "zero_length_sample":
        B, Tv, Dv, k = 3, 12, 40, 1
        lens = [5, 0, 12]
    elif case == "single_token":
        B, Tv, Dv, k = 1, 1, 40, 1
        lens = [1]
    elif case == "frames_fewer_than_stride":
        B, Tv, Dv, k = 2, 3, 40, 4   # 3 frames, stride 4 -> one token holding 3 real frames and one zero frame
        lens = [3, 2]
    else:
        B, Tv, Dv, k = 2, 9, 40, 1
        lens = [9, 4]
    v = torch.randn(B, Tv, Dv, generator=g)
    wv = torch.randn(H, k * Dv, generator=g) / (k * Dv) ** 0.5
    bv = torch.randn(H, generator=g) * 0.1
    table = torch.randn(V, H, generator=g)
    spec = O.ConnectorSpec(modality="video", video_stride=k, mask_mode=1, label_mode=1)
    vv = torch.tensor(lens)
    wv_c, bv_c = wv.clone().requires_grad_(True), bv.clone().requires_grad_(True)
    tok, _ = O.connector_tokens(None, v, None, None, wv_c, bv_c, spec, video_valid=vv)
    ntok = [-(-n // k) for n in lens]
    S = 3 + max(max(ntok), 1)
    ids = torch.zeros(B, S, dtype=torch.int64)
    for b, n in enumerate(ntok):
        row = torch.randint(1, PH, (S,), generator=g)
        row[1:1 + n] = PH
        row[1 + n + 1:] = 0
        ids[b] = row
    emb_r, mask_r, lab_r = O.splice_tokens(tok, ids, PH, table, 0, spec, ntok=torch.tensor(ntok))
    up = torch.randn(emb_r.shape, generator=g)
    (emb_r * up).sum().backward()
    wv_d, bv_d = wv.to(dev).requires_grad_(True), bv.to(dev).requires_grad_(True)
    plan = avc.FusePlan(modality="video", video_stride=k, mask_mode=1, label_mode=1)
    emb, mask, lab = avc.fused_connector(None, v.to(dev), None, None, wv_d, bv_d, plan, input_ids=ids.to(dev),
                                         placeholder_id=PH, embed_table=table.to(dev, torch.bfloat16),
                                         out_dtype=torch.bfloat16, video_lengths=lens, check=True)
    (emb.float() * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    is_ph = ids == PH
    if is_ph.any():
        assert_close(emb[is_ph.to(dev)], emb_r[is_ph], "AV rows")
    assert torch.equal(emb.cpu()[~is_ph], emb_r[~is_ph].to(torch.bfloat16))
    assert torch.equal(mask.cpu(), mask_r) and torch.equal(lab.cpu(), lab_r)
    assert_close(wv_d.grad, wv_c.grad, "dWv")
    assert_close(bv_d.grad, bv_c.grad, "dbv")


def test_all_samples_empty_is_handled(avc, cuda_dev):
    """M = 0: every sample has zero valid frames -> only text rows, zero gradients, no kernel faults."""
    dev = cuda_dev
    H, Dv, V = 64, 32, 20
    v = torch.randn(2, 6, Dv).to(dev)
    wv = torch.randn(H, Dv, device=dev, requires_grad=True)
    bv = torch.randn(H, device=dev, requires_grad=True)
    ids = torch.randint(1, V - 1, (2, 5)).to(dev)
    table = torch.randn(V, H).to(dev, torch.bfloat16)
    plan = avc.FusePlan(modality="video", mask_mode=1)
    emb, mask, _ = avc.fused_connector(None, v, None, None, wv, bv, plan, input_ids=ids, placeholder_id=V - 1,
                                       embed_table=table, out_dtype=torch.bfloat16, video_lengths=[0, 0], check=True)
    emb.float().sum().backward()
    torch.cuda.synchronize()
    assert torch.equal(emb, table[ids]) and int(mask.sum()) == 10
    assert float(wv.grad.abs().max()) == 0.0 and float(bv.grad.abs().max()) == 0.0


def test_rate_alignment_at_stride_one_repeats_video_frames(avc, cuda_dev):
    """align="rate", stride 1: token j = [audio frame j ; video frame j // 2] (50 Hz audio against 25 fps video)."""
    g = torch.Generator().manual_seed(91)
    B, Ta, Tv, Da, Dv, H = 2, 21, 10, 32, 16, 64   # 21 audio frames vs 10 video frames -> 20 video tokens, 21 tokens
    a, v = torch.randn(B, Ta, Da, generator=g), torch.randn(B, Tv, Dv, generator=g)
    wa, ba, wv, bv = _rand_params(g, H, Da, Dv)
    spec = O.ConnectorSpec(fusion="sum", fusion_scale=0.6, video_repeat=2, max_seq_len=64)
    tok_r, flags_r = O.connector_tokens(a, v, wa, ba, wv, bv, spec)
    assert tok_r.shape[1] == 21 and flags_r[0, 20] == 1 and flags_r[0, 19] == 3
    # hand check of the pairing: token 5 uses video frame 2
    manual = 0.6 * (a[:, 5] @ wa.t() + ba) + 0.4 * (v[:, 2] @ wv.t() + bv)
    assert torch.allclose(tok_r[:, 5], manual, atol=1e-5)
    up = torch.randn(tok_r.shape, generator=g)
    grads_r = O.connector_grads(a, v, wa, ba, wv, bv, spec, up)
    dev = cuda_dev
    params = [t.to(dev).requires_grad_(True) for t in (wa, ba, wv, bv)]
    plan = avc.FusePlan(fusion="sum", fusion_scale=0.6, video_repeat=2, max_seq_len=64)
    emb, mask, _ = avc.fused_connector(a.to(dev), v.to(dev), *params, plan, out_dtype=torch.float32, check=True)
    (emb * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    assert_close(emb, tok_r, "tokens")
    for p, gr, n in zip(params, grads_r, ["dWa", "dba", "dWv", "dbv"]):
        assert_close(p.grad, gr, n)
    # the model maps align="rate", stride=1 onto video_repeat = 2
    m_kw = dict(device="cuda:0", modality="both", align="rate", stride=1, _provided_tokenizer=SimpleNamespace(pad_token_id=0),
                _provided_llm=StubLLM(torch.zeros(8, H)).to(dev), _provided_whisper=StubWhisper(Da).to(dev),
                _provided_clip=StubClip(Dv).to(dev))
    m = avc.ClipWhisperModel(**m_kw)
    assert (m.audio_stride, m.video_stride, m.video_repeat) == (1, 1, 2) and m._plan().video_repeat == 2


def _random_case(seed):
    """One seeded configuration of the fused connector: modality, fusion, strides, ragged or dense, widths that are
    multiples of 8 only, any LLM dtype, with or without a prompt / input gradients."""
    g = torch.Generator().manual_seed(1000 + seed)

    def pick(seq):
        return seq[int(torch.randint(0, len(seq), (1,), generator=g))]

    c = SimpleNamespace(seed=seed, g=g)
    c.modality = pick(["audio", "video", "both", "both"])
    c.fusion = pick(["sum", "concat"])
    c.ka, c.kv = pick([1, 2, 4]), pick([1, 2])
    c.B = pick([1, 2, 3, 5])
    c.Ta, c.Tv = pick([5, 16, 33, 70]), pick([3, 8, 20, 41])
    c.Da, c.Dv = 8 * pick([1, 3, 8, 12]), 8 * pick([1, 2, 5, 8])
    c.H = 8 * pick([1, 9, 16, 33, 40])
    c.fs = pick([0.5, 0.25, 0.8])
    c.out_dtype = pick([torch.bfloat16, torch.bfloat16, torch.float32, torch.float16])
    c.ragged = pick([False, True])
    c.P = pick([0, 3, 7])
    c.dx = pick([False, False, True])
    c.max_seq_len = pick([256, 12])
    return c


@pytest.mark.parametrize("seed", range(48))
def test_randomised_configurations_vs_oracle(avc, cuda_dev, seed):
    """Seeded sweep over the connector's configuration space (the product of what the other tests fix one at a time):
    outputs, masks, labels, parameter gradients and -- when the features require grad -- input gradients against the
    CPU oracle.  Index / mask / label / copied-row results bit-exact; projected values within REL_TOL / COS_TOL."""
    c = _random_case(seed)
    g, dev = c.g, cuda_dev
    use_a, use_v = c.modality in ("audio", "both"), c.modality in ("video", "both")
    V, PH = 50, 49
    # bf16-representable inputs / weights so that the comparison measures the kernels, not the input rounding
    a = torch.randn(c.B, c.Ta, c.Da, generator=g).bfloat16().float() if use_a else None
    v = torch.randn(c.B, c.Tv, c.Dv, generator=g).bfloat16().float() if use_v else None
    wa, ba, wv, bv = _rand_params(g, c.H, c.ka * c.Da, c.kv * c.Dv)
    table = torch.randn(V, c.H, generator=g).to(c.out_dtype)
    spec = O.ConnectorSpec(modality=c.modality, fusion=c.fusion, fusion_scale=c.fs, audio_stride=c.ka,
                           video_stride=c.kv, max_seq_len=c.max_seq_len, mask_mode=1, label_mode=1)
    plan = avc.FusePlan(modality=c.modality, fusion=c.fusion, fusion_scale=c.fs, audio_stride=c.ka, video_stride=c.kv,
                        max_seq_len=c.max_seq_len, mask_mode=1, label_mode=1)
    N = plan.tokens(c.Ta if use_a else None, c.Tv if use_v else None)
    la = lv = None
    if c.ragged:
        la = [int(torch.randint(0, c.Ta + 1, (1,), generator=g)) for _ in range(c.B)] if use_a else None
        lv = [int(torch.randint(0, c.Tv + 1, (1,), generator=g)) for _ in range(c.B)] if use_v else None
        ntok = [plan.tokens(la[b] if use_a else None, lv[b] if use_v else None) for b in range(c.B)]
    else:
        ntok = [N] * c.B
    S = c.P + max(max(ntok), 1) + 2
    ids = torch.zeros(c.B, S, dtype=torch.int64)
    for b, n in enumerate(ntok):
        row = torch.randint(1, PH, (S,), generator=g)
        row[c.P:c.P + n] = PH
        row[c.P + n + 1:] = 0   # one text id after the AV run, then padding
        ids[b] = row
    leaves = [t.clone().requires_grad_(True) if t is not None else None for t in (a, v, wa, ba, wv, bv)]
    a_c, v_c, wa_c, ba_c, wv_c, bv_c = leaves
    tok, _ = O.connector_tokens(a_c, v_c, wa_c, ba_c, wv_c, bv_c, spec,
                                audio_valid=torch.tensor(la) if la is not None else None,
                                video_valid=torch.tensor(lv) if lv is not None else None)
    emb_r, mask_r, lab_r = O.splice_tokens(tok, ids, PH, table.float(), 0, spec, ntok=torch.tensor(ntok))
    up = torch.randn(emb_r.shape, generator=g)
    if emb_r.requires_grad:
        (emb_r * up).sum().backward()

    def dev_leaf(t, grad):
        return None if t is None else t.to(dev).requires_grad_(grad)

    a_d, v_d = dev_leaf(a, c.dx), dev_leaf(v, c.dx)
    wa_d, ba_d, wv_d, bv_d = (dev_leaf(t, True) for t in (wa, ba, wv, bv))
    emb, mask, lab = avc.fused_connector(a_d, v_d, wa_d if use_a else None, ba_d if use_a else None,
                                         wv_d if use_v else None, bv_d if use_v else None, plan,
                                         input_ids=ids.to(dev), placeholder_id=PH, embed_table=table.to(dev),
                                         out_dtype=c.out_dtype, audio_lengths=la, video_lengths=lv, check=True)
    assert emb.dtype == c.out_dtype
    (emb.float() * up.to(dev)).sum().backward()
    torch.cuda.synchronize()
    is_ph = ids == PH
    what = f"seed {seed}: {vars(c)}"
    if is_ph.any():
        assert_close(emb[is_ph.to(dev)].float(), emb_r[is_ph], "AV rows, " + what)
    assert torch.equal(emb.cpu()[~is_ph], emb_r[~is_ph].to(c.out_dtype)), what
    assert torch.equal(mask.cpu(), mask_r) and torch.equal(lab.cpu(), lab_r), what
    pairs = []
    if use_a:
        pairs += [(wa_d, wa_c, "dWa"), (ba_d, ba_c, "dba")]
    if use_v:
        pairs += [(wv_d, wv_c, "dWv"), (bv_d, bv_c, "dbv")]
    if c.dx:
        pairs += [(t_d, t_c, n) for t_d, t_c, n in ((a_d, a_c, "dA"), (v_d, v_c, "dV")) if t_d is not None]
    for got, ref, name in pairs:
        ref_g = ref.grad if ref.grad is not None else torch.zeros_like(ref)
        got_g = got.grad if got.grad is not None else torch.zeros_like(got)
        if float(ref_g.abs().max()) == 0.0:
            assert float(got_g.abs().max()) == 0.0, f"{name} must be zero, {what}"
        else:
            assert_close(got_g, ref_g, f"{name}, {what}")
