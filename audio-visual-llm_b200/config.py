"""Config keys of the connector path, loadable from the reference's `configs/clip_whisper.yaml`.

Keys that touch the hot path in the reference: `model.modality` (:21), `model.use_fp16` (:22),
`model.freeze_encoders` (:28), `model.fusion_scale` (:30), `data.max_seq_len` (:13), plus the CLI-only
`connector_type` (scripts/clip_whisper/train.py:77).  The reference's train.py merges the yaml FLAT
(train.py:134-139) so nested keys are silently ignored unless repeated as CLI flags; here both the nested layout
of the shipped yaml and flat keys are honoured, flat keys winning.
New keys (defaults = reference behaviour): `fusion` (sum|concat), `stride`, `align` (index|rate), `audio_stride`,
`video_stride`, `mask_mode`, `label_mode`, `connector_type: mlp`.
"""
from __future__ import annotations

from typing import Any, Dict, Mapping

REFERENCE_KEYS = {
    "llm_path": "meta-llama/Llama-2-7b-chat-hf", "whisper_model": "openai/whisper-medium",
    "clip_model": "openai/clip-vit-base-patch32", "use_fp16": False, "use_4bit": False, "use_lora": True,
    "lora_r": 16, "lora_alpha": 32, "lora_dropout": 0.05, "freeze_encoders": True, "freeze_llm": False,
    "modality": "both", "max_seq_len": 256, "fusion_scale": 0.5, "connector_type": "simple",
}
NEW_KEYS = {"fusion": "sum", "stride": 1, "align": "index", "audio_stride": None, "video_stride": None,
            "mask_mode": 0, "label_mode": 0}


def model_kwargs(cfg: Mapping[str, Any]) -> Dict[str, Any]:
    """Constructor keyword arguments for ClipWhisperModel from a (nested or flat) config mapping."""
    out = dict(REFERENCE_KEYS)
    out.update(NEW_KEYS)
    model = cfg.get("model", {}) or {}
    data = cfg.get("data", {}) or {}
    for k in list(out):
        if k in model:
            out[k] = model[k]
    if "max_seq_len" in data:
        out["max_seq_len"] = data["max_seq_len"]  # configs/clip_whisper.yaml:13 keeps it under data:
    for k in list(out):
        if k in cfg and not isinstance(cfg[k], Mapping):
            out[k] = cfg[k]  # flat keys (what train.py's CLI merge produces) win
    if out["modality"] not in ("audio", "video", "both"):
        raise ValueError(f"modality must be audio|video|both, got {out['modality']!r}")
    if out["fusion"] not in ("sum", "concat"):
        raise ValueError(f"fusion must be sum|concat, got {out['fusion']!r}")
    if out["align"] not in ("index", "rate"):
        raise ValueError(f"align must be index|rate, got {out['align']!r}")
    return out


def load_config(path: str) -> Dict[str, Any]:
    """Read a yaml file (e.g. the reference's configs/clip_whisper.yaml) and return ClipWhisperModel kwargs."""
    import yaml

    with open(path) as f:
        return model_kwargs(yaml.safe_load(f) or {})
