"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by EXECUTING the reference (oracle/ref_loader.py).

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Each fixture holds the seeded inputs and what the unmodified reference produced for them: the tensors its
forward()/encode() hand to the LLM (inputs_embeds, attention_mask, labels) and autograd's connector gradients
for the upstream gradient G stored alongside.  cfg1 (BASELINE.json configs[0]) is stored sub-sampled.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_loader as R  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

# name -> case description.  Small shapes: B=2, Da=32, Dv=16, H=48, vocab 64, pad id 0.
CASES = {
    # both streams, audio longer, prompt, eval branch with short labels (right-pad -100)
    "both_prompt_eval": dict(modality="both", Ta=24, Tv=12, fs=0.5, max_seq_len=512, P=4, L=10, train=False),
    # max_seq_len truncation (24 -> 16), asymmetric fusion scale, labels longer than the sequence (truncate)
    "both_truncate_eval": dict(modality="both", Ta=24, Tv=12, fs=0.3, max_seq_len=16, P=0, L=40, train=False),
    # video longer than audio: audio rows past Ta carry no audio bias
    "both_video_longer": dict(modality="both", Ta=9, Tv=20, fs=0.7, max_seq_len=256, P=3, L=23, train=False),
    # single-modality branches ignore max_seq_len (clip_whisper_model.py:436-443)
    "audio_only_nocap": dict(modality="audio", Ta=24, Tv=None, fs=0.5, max_seq_len=16, P=2, L=26, train=False),
    "video_only_cls": dict(modality="video", Ta=None, Tv=12, fs=0.5, max_seq_len=256, P=0, L=5, train=False),
    # prompt longer than 32 tokens is cut to 32 (clip_whisper_model.py:469, 481-482)
    "prompt_cap32": dict(modality="both", Ta=8, Tv=8, fs=0.5, max_seq_len=256, P=40, L=40, train=False),
    # training branch: adaptive avg-pool to the label length / linear interpolation up to it
    "train_pool": dict(modality="both", Ta=24, Tv=12, fs=0.5, max_seq_len=512, P=4, L=10, train=True),
    "train_interp": dict(modality="both", Ta=8, Tv=6, fs=0.5, max_seq_len=512, P=0, L=19, train=True),
    # round 2: more of the reference's branches
    # video only with a prompt, training branch: 5 + 10 rows interpolated up to 30 label positions
    "train_video_only_prompt": dict(modality="video", Ta=None, Tv=10, fs=0.5, max_seq_len=256, P=5, L=30, train=True),
    # audio only, training branch: 40 rows pooled down to 12
    "train_audio_only_pool": dict(modality="audio", Ta=40, Tv=None, fs=0.5, max_seq_len=256, P=0, L=12, train=True),
    # fusion_scale 1.0 (the video projection is multiplied by 0), labels exactly as long as the sequence
    "both_equal_len_fs1": dict(modality="both", Ta=16, Tv=16, fs=1.0, max_seq_len=256, P=2, L=18, train=False),
    # fusion_scale 0.0, fused length exactly at the cap
    "both_fs0_cap_exact": dict(modality="both", Ta=16, Tv=16, fs=0.0, max_seq_len=16, P=0, L=16, train=False),
    # training branch with labels as long as the sequence: no length adaptation
    "train_same_len": dict(modality="both", Ta=12, Tv=6, fs=0.4, max_seq_len=512, P=4, L=16, train=True),
}
B, DA, DV, H, VOCAB, PAD, NP = 2, 32, 16, 48, 64, 0, 3


def make_inputs(name: str, c: dict):
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    d = {}
    if c["Ta"] is not None:
        d["audio_feats"] = torch.randn(B, c["Ta"], DA, generator=g)
    if c["Tv"] is not None:
        d["clip_hidden"] = torch.randn(B * c["Tv"], 1 + NP, DV, generator=g)
    if c["P"]:
        d["prompt"] = torch.randint(1, VOCAB, (B, c["P"]), generator=g)
    lab = torch.randint(1, VOCAB, (B, c["L"]), generator=g)
    lab[0, c["L"] // 2:] = PAD  # pad tail on one sample
    d["labels"] = lab
    d["bias_a"] = torch.randn(H, generator=g) * 0.02
    d["bias_v"] = torch.randn(H, generator=g) * 0.02
    return d


def run_case(name: str, c: dict):
    inp = make_inputs(name, c)
    m = R.build_reference_model(DA, DV, H, modality=c["modality"], max_seq_len=c["max_seq_len"],
                                fusion_scale=c["fs"], vocab=VOCAB, pad_token_id=PAD, seed=0)
    with torch.no_grad():  # non-zero biases so the pad-after-projection bias mask is exercised
        m.audio_connector.linear.bias.copy_(inp["bias_a"])
        m.video_connector.linear.bias.copy_(inp["bias_v"])
    out = R.run_reference(m, inp.get("audio_feats"), inp.get("clip_hidden"), c["Tv"], prompt=inp.get("prompt"),
                          labels=inp["labels"], train=c["train"])
    S = out["inputs_embeds"].shape[1]
    g = torch.Generator().manual_seed(1234)
    upstream = torch.randn(B, S, H, generator=g)  # what StubLLM used (same seed, same shape)
    rec = {f"in.{k}": v.numpy() for k, v in inp.items()}
    rec["in.upstream"] = upstream.numpy()
    rec["in.embed_table"] = m.llm.embed.weight.detach().numpy()
    for n in ("audio_connector", "video_connector"):
        rec[f"in.{n}.linear.weight"] = getattr(m, n).linear.weight.detach().numpy()
        rec[f"in.{n}.linear.bias"] = getattr(m, n).linear.bias.detach().numpy()
    for k, v in out.items():
        rec[f"out.{k}"] = v.numpy()
    rec["cfg"] = np.array(repr(c))
    return rec


def cfg1():
    """BASELINE.json configs[0]: Whisper-small(768)+CLIP(512)->2048, batch 2, 10 s clips, fp32, parity mode."""
    Bc, Ta, Tv, Da, Dv, Hc = 2, 500, 250, 768, 512, 2048
    g = torch.Generator().manual_seed(1234 + 1)
    a = torch.randn(Bc, Ta, Da, generator=g)
    v = torch.randn(Bc, Tv, Dv, generator=g)
    m = R.build_reference_model(Da, Dv, Hc, modality="both", max_seq_len=512, fusion_scale=0.5, seed=0)
    gb = torch.Generator().manual_seed(1)
    with torch.no_grad():
        m.audio_connector.linear.bias.copy_(torch.randn(Hc, generator=gb) * 0.02)
        m.video_connector.linear.bias.copy_(torch.randn(Hc, generator=gb) * 0.02)
    gu = torch.Generator().manual_seed(77)
    up = torch.randn(Bc, Ta, Hc, generator=gu)
    out = R.run_reference(m, a, v.unsqueeze(2).reshape(Bc * Tv, 1, Dv), Tv, upstream=up, call="encode")
    rec = {
        "in.checksum": np.array([a.double().sum().item(), v.double().sum().item(), up.double().sum().item()]),
        "out.inputs_embeds[::25, ::64]": out["inputs_embeds"][:, ::25, ::64].numpy(),
        "out.inputs_embeds.sum": np.array(out["inputs_embeds"].double().sum().item()),
        "out.attention_mask.sum": np.array(out["attention_mask"].sum().item()),
    }
    for n in ("audio_connector", "video_connector"):
        rec[f"out.{n}.linear.weight.grad[::32, ::32]"] = out[f"{n}.linear.weight.grad"][::32, ::32].numpy()
        rec[f"out.{n}.linear.bias.grad"] = out[f"{n}.linear.bias.grad"].numpy()
        rec[f"w.checksum.{n}"] = np.array(getattr(m, n).linear.weight.double().sum().item())
    return rec


def adaptive_connector():
    """The reference's `adaptive` connector type (modality_connector.py:239-382) built by its own factory, eval mode:
    weights, one short input (attention block only) and one longer than 512 frames (two stride-2 convolutions first),
    outputs and the gradients of both dense projections for a fixed upstream gradient."""
    mc, _ = R.load_reference_modules()
    torch.manual_seed(11)
    conn = mc.create_modality_connector("adaptive", 32, 48, device="cpu", dtype=torch.float32, max_seq_len=640).eval()
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():  # the factory zero-initialises every bias: give them values so that they are exercised
        for n, p in conn.named_parameters():
            if n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    x = torch.randn(2, 40, 32, generator=g)
    x_long = torch.randn(1, 520, 32, generator=g)
    up = torch.randn(2, 40, 48, generator=g)
    y = conn(x)
    (y * up).sum().backward()
    rec = {f"sd.{k}": v.detach().numpy() for k, v in conn.state_dict().items()}
    rec.update({"in.x": x.numpy(), "in.x_long": x_long.numpy(), "in.upstream": up.numpy(), "out.y": y.detach().numpy(),
                "out.input_proj.weight.grad": conn.input_proj.weight.grad.numpy(),
                "out.input_proj.bias.grad": conn.input_proj.bias.grad.numpy(),
                "out.output_proj.weight.grad": conn.output_proj.weight.grad.numpy(),
                "out.output_proj.bias.grad": conn.output_proj.bias.grad.numpy()})
    with torch.no_grad():
        rec["out.y_long"] = conn(x_long).numpy()
    return rec


def main():
    if not R.available():
        raise SystemExit("/root/reference is not mounted: goldens can only be generated in the build container")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    for name, c in CASES.items():
        np.savez_compressed(GOLDEN / f"ref_{name}.npz", **run_case(name, c))
        print("wrote", name)
    np.savez_compressed(GOLDEN / "ref_cfg1_subsampled.npz", **cfg1())
    print("wrote cfg1")
    np.savez_compressed(GOLDEN / "adaptive_connector.npz", **adaptive_connector())
    print("wrote adaptive connector")


if __name__ == "__main__":
    main()
