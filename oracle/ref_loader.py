"""TEST INFRASTRUCTURE ONLY -- loads the *executing* reference (read-only, /root/reference) on CPU.

Recipe: stub `peft`, load modality_connector.py and clip_whisper_model.py by file
path under a synthetic package (so the package __init__, which needs librosa/soundfile/matplotlib, never
runs), bypass ClipWhisperModel.__init__ (it downloads towers) and set only the attributes that
encode()/forward() read.  Towers are stubs returning preset features; the LLM is a stub with an embedding
table that records what it is handed.  The reference's encode()/forward()/autograd then run unmodified.

Only available in the build container (the GPU box has no /root/reference); used by make_golden.py and by
CPU tests that skip when the path is absent.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path
from types import SimpleNamespace

import torch
import torch.nn as nn

REF_ROOT = Path("/root/reference")
_MODELS = REF_ROOT / "src" / "clip_whisper" / "models"
_PKG = "_avsr_ref_models"


def available() -> bool:
    return (_MODELS / "clip_whisper_model.py").exists()


def load_reference_modules():
    """Returns (modality_connector module, clip_whisper_model module) of the reference."""
    if _PKG + ".clip_whisper_model" in sys.modules:
        return sys.modules[_PKG + ".modality_connector"], sys.modules[_PKG + ".clip_whisper_model"]
    if "peft" not in sys.modules:
        stub = types.ModuleType("peft")
        stub.LoraConfig = object
        stub.get_peft_model = lambda model, cfg: model
        sys.modules["peft"] = stub
    pkg = types.ModuleType(_PKG)
    pkg.__path__ = [str(_MODELS)]
    sys.modules[_PKG] = pkg
    mods = []
    for name in ("modality_connector", "clip_whisper_model"):
        spec = importlib.util.spec_from_file_location(f"{_PKG}.{name}", _MODELS / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"{_PKG}.{name}"] = mod
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


class _StubWhisperEncoder(nn.Module):
    def __init__(self, owner):
        super().__init__()
        self._owner = [owner]

    def forward(self, audio, attention_mask=None, output_hidden_states=True, return_dict=True):
        return SimpleNamespace(last_hidden_state=self._owner[0].features)


class StubWhisper(nn.Module):
    """whisper.encoder(...) -> .last_hidden_state = preset [B, Ta, Da] features."""

    def __init__(self):
        super().__init__()
        self.dummy = nn.Parameter(torch.zeros(1))
        self.features = None
        self.encoder = _StubWhisperEncoder(self)


class StubClip(nn.Module):
    """clip(flat_video, return_dict=True) -> .last_hidden_state = preset [B*F, 1+Np, Dv] hidden states."""

    def __init__(self):
        super().__init__()
        self.dummy = nn.Parameter(torch.zeros(1))
        self.hidden = None

    def forward(self, flat_video, return_dict=True):
        return SimpleNamespace(last_hidden_state=self.hidden)


class StubLLM(nn.Module):
    """Embedding table + a recorder: loss = sum(inputs_embeds * G) for a fixed G so gradients are defined."""

    def __init__(self, vocab: int, hidden: int, seed: int = 99):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.embed = nn.Embedding(vocab, hidden)
        with torch.no_grad():
            self.embed.weight.copy_(torch.randn(vocab, hidden, generator=g))
        self.calls = []
        self.upstream = None  # G, set by the caller ([B, S, H]) or generated lazily

    def get_input_embeddings(self):
        return self.embed

    def forward(self, inputs_embeds=None, attention_mask=None, labels=None, return_dict=True):
        self.calls.append({"inputs_embeds": inputs_embeds, "attention_mask": attention_mask, "labels": labels})
        G = self.upstream
        if G is None or G.shape != inputs_embeds.shape:
            g = torch.Generator().manual_seed(1234)
            G = torch.randn(inputs_embeds.shape, generator=g)
        return SimpleNamespace(loss=(inputs_embeds * G).sum(), logits=inputs_embeds)


def build_reference_model(audio_dim, video_dim, llm_dim, *, modality="both", max_seq_len=256, fusion_scale=0.5,
                          vocab=64, pad_token_id=0, connector_type="simple", seed=0):
    """ClipWhisperModel with __init__ bypassed; connectors built by the reference's own factory."""
    mc, cwm = load_reference_modules()
    m = cwm.ClipWhisperModel.__new__(cwm.ClipWhisperModel)
    nn.Module.__init__(m)
    m.device = "cpu"
    m.use_fp16 = False
    m.freeze_encoders = False  # with True the reference gives the connectors no gradient (clip_whisper_model.py:1096)
    m.modality = modality
    m.max_seq_len = max_seq_len
    m.fusion_scale = fusion_scale
    m.dtype = torch.float32
    m.connector_type = connector_type
    m.audio_dim, m.video_dim, m.llm_dim = audio_dim, video_dim, llm_dim
    m.tokenizer = SimpleNamespace(pad_token_id=pad_token_id)
    m.llm = StubLLM(vocab, llm_dim)
    m.whisper = StubWhisper()
    m.clip = StubClip()
    torch.manual_seed(seed)
    m._setup_projections()  # reference factory: xavier-uniform W, zero b (modality_connector.py:35-36)
    return m


def run_reference(m, audio_feats=None, clip_hidden=None, frames=None, prompt=None, labels=None, train=False,
                  upstream=None, call="forward"):
    """Drive the reference's encode()/forward() with preset tower outputs.

    audio_feats [B, Ta, Da]; clip_hidden [B*F, 1+Np, Dv] with frames=F.  Returns a dict of CPU tensors."""
    m.train(train)
    for p in m.parameters():
        p.grad = None
    audio = video = None
    if audio_feats is not None:
        m.whisper.features = audio_feats
        audio = torch.zeros(audio_feats.shape[0], 80, 4)
    if clip_hidden is not None:
        m.clip.hidden = clip_hidden
        video = torch.zeros(clip_hidden.shape[0] // frames, frames, 3, 2, 2)
    m.llm.calls.clear()
    m.llm.upstream = upstream
    out = {}
    if call == "encode":
        emb, mask = m.encode(audio, video, prompt)
        out["inputs_embeds"], out["attention_mask"] = emb.detach(), mask
        if upstream is not None:
            (emb * upstream).sum().backward()
    else:
        res = m.forward(audio=audio, video=video, prompt=prompt, labels=labels, return_loss=True)
        rec = m.llm.calls[-1]
        out["inputs_embeds"] = rec["inputs_embeds"].detach()
        out["attention_mask"] = rec["attention_mask"]
        if rec["labels"] is not None:
            out["labels"] = rec["labels"]
        out["loss"] = res["loss"].detach()
        res["loss"].backward()
    for name in ("audio_connector", "video_connector"):
        lin = getattr(m, name).linear
        if lin.weight.grad is not None:
            out[f"{name}.linear.weight.grad"] = lin.weight.grad.clone()
            out[f"{name}.linear.bias.grad"] = lin.bias.grad.clone()
    return out
