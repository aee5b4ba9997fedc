"""Trainer-step kernels against the reference's own optimizer recipe run on the CPU:
torch.optim.AdamW(betas=(0.9, 0.95), eps=1e-8) with bias / non-bias weight-decay groups
(clip_whisper_trainer.py:171-207) after torch.nn.utils.clip_grad_norm_ over ALL parameters (:458)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sumsq_is_deterministic_and_exact(avc, cuda_dev):
    L = avc._lib
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1_000_003, generator=g)
    xd = x.to(cuda_dev)
    out = torch.zeros(1, device=cuda_dev)
    ws = L.sumsq_workspace(cuda_dev)
    L.sumsq(xd, out, ws)
    first = float(out.item())
    ref = float((x.double() ** 2).sum())
    assert abs(first - ref) <= 1e-5 * ref
    L.sumsq(xd, out, ws)
    assert float(out.item()) == first  # same bits on a second run
    L.sumsq(xd, out, ws, accumulate=True)
    assert abs(float(out.item()) - 2 * ref) <= 2e-5 * ref


@pytest.mark.parametrize("other_norm", [0.0, 3.0])
def test_adamw_with_global_clip_matches_torch(avc, cuda_dev, other_norm):
    from audio_visual_llm_b200.parallel import GradBucket
    from audio_visual_llm_b200.trainer_step import ConnectorAdamW

    g = torch.Generator().manual_seed(1)
    H, Ka, Kv = 64, 96, 32
    shapes = {"audio_connector.linear.weight": (H, Ka), "video_connector.linear.weight": (H, Kv),
              "audio_connector.linear.bias": (H,), "video_connector.linear.bias": (H,)}
    cpu = {n: torch.nn.Parameter(torch.randn(*s, generator=g) * 0.1) for n, s in shapes.items()}
    extra = torch.nn.Parameter(torch.zeros(4))  # stands for the LoRA parameters sharing the global-norm clip
    dev = {n: p.detach().clone().to(cuda_dev) for n, p in cpu.items()}
    decay = [p for n, p in cpu.items() if "bias" not in n]
    no_decay = [p for n, p in cpu.items() if "bias" in n]
    ref_opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.01}, {"params": no_decay, "weight_decay": 0.0}],
                                lr=2e-3, betas=(0.9, 0.95), eps=1e-8)
    bucket = GradBucket(shapes, cuda_dev)
    opt = ConnectorAdamW(dev.items(), bucket, lr=2e-3, weight_decay=0.01, betas=(0.9, 0.95), eps=1e-8,
                         max_grad_norm=0.5)
    sa = 0.3
    packed = torch.zeros(H, Ka + Kv, dtype=torch.bfloat16, device=cuda_dev)
    opt.attach_packed("audio_connector.linear.weight", packed[:, :Ka], sa)
    opt.attach_packed("video_connector.linear.weight", packed[:, Ka:], 1 - sa)
    for step in range(4):
        for n, p in cpu.items():
            p.grad = torch.randn(p.shape, generator=g) * (0.05 if step % 2 else 2.0)  # clipped and unclipped steps
            bucket[n].copy_(p.grad)
        extra.grad = torch.full((4,), other_norm / 2.0)
        torch.nn.utils.clip_grad_norm_(list(cpu.values()) + [extra], 0.5)
        ref_opt.step()
        other = torch.tensor([other_norm ** 2], device=cuda_dev)
        opt.step(other_sumsq=other)
        torch.cuda.synchronize()
        for n in cpu:
            # fp32 elementwise update: identical formula, different contraction of the scalar prefactors
            assert torch.allclose(dev[n].cpu(), cpu[n].detach(), rtol=2e-5, atol=2e-7), (step, n)
    exp = torch.cat([(dev["audio_connector.linear.weight"] * sa).bfloat16(),
                     (dev["video_connector.linear.weight"] * torch.tensor(1 - sa, dtype=torch.float32)).bfloat16()], 1)
    assert torch.equal(packed.view(torch.int16), exp.view(torch.int16))


@pytest.mark.parametrize("fuse_gather", [True, False])
def test_engine_train_step_equals_step_then_optimizer(avc, cuda_dev, fuse_gather):
    """ConnectorStep.train_step() (optimizer attached: AdamW emits the bf16 weight pack, the forward skips its pack
    launches) must leave the same weights, bit for bit, as step() followed by a stand-alone ConnectorAdamW.step()."""
    from audio_visual_llm_b200.engine import ConnectorStep, StepShape
    from audio_visual_llm_b200.trainer_step import ConnectorAdamW

    L = avc._lib
    shape = StepShape(batch=3, audio_frames=40, video_frames=20, audio_dim=32, video_dim=16, hidden=128, prompt_len=5,
                      vocab=100)
    plan = avc.FusePlan(fusion="concat", audio_stride=4, video_stride=2, max_seq_len=64)
    a = ConnectorStep(shape, plan, cuda_dev, seed=5, fuse_gather=fuse_gather)
    b = ConnectorStep(shape, plan, cuda_dev, seed=5, fuse_gather=fuse_gather)
    kw = dict(lr=1e-2, weight_decay=0.01, max_grad_norm=0.5)
    a.attach_optimizer(**kw)
    opt_b = ConnectorAdamW(b._named_params(), bucket=b.bucket, **kw)
    w0 = a.wa.clone()
    for _ in range(3):
        a.train_step()
        b.step()
        opt_b.step()
    torch.cuda.synchronize()
    assert not torch.equal(a.wa, w0), "the optimizer must have moved the weights"
    for (n, pa), (_, pb) in zip(a._named_params(), b._named_params()):
        assert torch.equal(pa, pb), n
    assert torch.equal(a.emb, b.emb) and torch.equal(a.bucket.flat, b.bucket.flat)
    # the pack the optimizer maintains is exactly what the forward's own pack kernel would produce
    ref = torch.empty_like(a.wp)
    L.pack_weight(a.wa, ref[:, :a.Ka], a.sa)
    L.pack_weight(a.wv, ref[:, a.Ka:], a.sv)
    assert torch.equal(ref, a.wp)
    a.detach_optimizer()
    a.step()
    torch.cuda.synchronize()
    with pytest.raises(L.ConnectorError):
        a.train_step()
