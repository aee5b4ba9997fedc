"""Drop-in `ClipWhisperModel` for the connector path of the reference's
src/clip_whisper/models/clip_whisper_model.py: same `encode / encode_audio / encode_video / forward / generate`
interface, constructor keywords and attributes, with the connector (align -> stack -> fuse -> project -> splice
+ masks) running as sm_100a CUDA.  Whisper, CLIP and the LLM stay untouched PyTorch modules and are called
exactly where the reference calls them (clip_whisper_model.py:1098-1103, 1138, 601-613, 1337-1340).

New keyword arguments default to the reference's behaviour:
  fusion="sum" | "concat"      stride=1 (frames stacked per token)      align="index" | "rate"
  audio_stride / video_stride  (override stride/align)                   mask_mode / label_mode (0 = reference)
`align="rate"` pairs 2 Whisper frames (50 Hz) with 1 CLIP frame (25 fps): k_a = stride, k_v = stride / 2; at
stride 1 every video frame is used for two consecutive tokens instead (video_repeat = 2).

`use_fp16=True` gives an fp16 LLM / fp16 `inputs_embeds` as in the reference (:164); `freeze_encoders=False` lets the
connector's input gradients (dX GEMM + the gather's transpose) flow into the towers (:1096, :1136); `use_lora`,
`use_4bit` and `freeze_llm` act on the LLM tower exactly as in the reference's `_load_llm` (:896-1016).

There is no CPU fallback (the reference's `"cuda" if torch.cuda.is_available() else "cpu"` default,
clip_whisper_model.py:91, is removed): the device must be an sm_100 GPU.
"""
from __future__ import annotations

import logging
from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from .connector_ops import MAX_PROMPT_LEN, FusePlan, fused_connector
from .modality_connector import create_modality_connector
from .seq_adapt import adapt_mask, adaptive_projection


class ClipWhisperModel(nn.Module):
    def __init__(
        self,
        llm_path: str = "meta-llama/Llama-2-7b-chat-hf",
        whisper_model: str = "openai/whisper-medium",
        clip_model: str = "openai/clip-vit-base-patch32",
        device: str = "cuda",
        use_fp16: bool = False,
        use_4bit: bool = False,
        use_lora: bool = True,
        lora_r: int = 16,
        lora_alpha: int = 32,
        lora_dropout: float = 0.05,
        freeze_encoders: bool = True,
        freeze_llm: bool = False,
        modality: str = "both",
        max_seq_len: int = 256,
        fusion_scale: float = 0.5,
        connector_type: str = "simple",
        _provided_tokenizer=None,
        _provided_llm=None,
        _provided_whisper=None,
        _provided_clip=None,
        # ---- new keys (defaults = reference behaviour)
        fusion: str = "sum",
        stride: int = 1,
        align: str = "index",
        audio_stride: Optional[int] = None,
        video_stride: Optional[int] = None,
        mask_mode: int = 0,
        label_mode: int = 0,
        audio_dim: Optional[int] = None,
        video_dim: Optional[int] = None,
    ):
        super().__init__()
        if not str(device).startswith("cuda"):
            raise L.ConnectorError(f"device={device!r}: the B200 connector path has no CPU fallback")
        idx = torch.device(device).index
        L.require_device(0 if idx is None else idx)
        if align not in ("index", "rate"):
            raise ValueError("align must be 'index' or 'rate'")
        self.device = device
        self.use_fp16 = use_fp16
        self.freeze_encoders = freeze_encoders
        self.freeze_llm = freeze_llm
        self.modality = modality
        self.max_seq_len = max_seq_len
        self.fusion_scale = fusion_scale
        self.connector_type = connector_type
        self.fusion = fusion
        ka = audio_stride if audio_stride is not None else stride
        if video_stride is not None:
            kv = video_stride
        elif align == "rate":
            if stride == 1:
                kv = 1  # every 25 fps video frame is paired with two 50 Hz audio frames (video_repeat = 2)
            elif stride % 2:
                raise ValueError("align='rate' needs stride 1 or an even stride (2 audio frames per video frame)")
            else:
                kv = stride // 2
        else:
            kv = stride
        self.audio_stride, self.video_stride = ka, kv
        self.video_repeat = 2 if (align == "rate" and stride == 1 and video_stride is None) else 1
        self.mask_mode, self.label_mode = mask_mode, label_mode
        self.dtype = torch.float16 if use_fp16 else torch.float32  # clip_whisper_model.py:164
        self.use_4bit, self.use_lora = use_4bit, use_lora
        self.lora_r, self.lora_alpha, self.lora_dropout = lora_r, lora_alpha, lora_dropout
        self.grad_sync = None  # parallel.FusedGradSync once enable_data_parallel() has been called

        self.tokenizer = _provided_tokenizer
        self.llm = _provided_llm
        self.whisper = _provided_whisper
        self.clip = _provided_clip
        if self.llm is None or self.tokenizer is None:
            self.tokenizer, self.llm = self._load_llm(llm_path)
        elif freeze_llm:
            self._freeze_llm(self.llm, keep_lora=use_lora)  # a provided LLM is frozen the same way (:1005-1014)
        if self.whisper is None and modality in ("audio", "both"):
            self.whisper = self._load_tower("WhisperModel", whisper_model)
        if self.clip is None and modality in ("video", "both"):
            self.clip = self._load_tower("CLIPVisionModel", clip_model)
        if freeze_encoders:
            for tower in (self.whisper, self.clip):
                if tower is not None:
                    for p in tower.parameters():
                        p.requires_grad = False  # clip_whisper_model.py:873, 893

        # dims as the reference derives them (:216, :247, :1148-1157); defaults for an absent modality (:223, :254)
        self.audio_dim = audio_dim or (self.whisper.config.d_model if self.whisper is not None else 1024)
        self.video_dim = video_dim or (self.clip.config.hidden_size if self.clip is not None else 768)
        self.llm_dim = self._get_llm_dim()
        self._setup_projections()

    # ------------------------------------------------------------------ construction helpers
    def _load_llm(self, llm_path):
        """clip_whisper_model.py:896-1016: optional 4-bit load, optional LoRA wrap (scaled-down init), optional
        freeze that leaves the LoRA weights trainable.  The LLM is a tower (untouched PyTorch / peft / bitsandbytes);
        a flag whose library is missing raises instead of being ignored."""
        from transformers import AutoModelForCausalLM, AutoTokenizer

        tok = self.tokenizer if self.tokenizer is not None else AutoTokenizer.from_pretrained(llm_path)
        if tok.pad_token is None:
            tok.pad_token = tok.eos_token  # clip_whisper_model.py:955-959
        if self.use_4bit:
            try:
                import bitsandbytes  # noqa: F401
                from transformers import BitsAndBytesConfig
            except ImportError as e:
                raise NotImplementedError("use_4bit=True needs bitsandbytes (clip_whisper_model.py:905-928), which is "
                                          "not installed; pass use_4bit=False or a pre-quantised _provided_llm") from e
            q = BitsAndBytesConfig(load_in_4bit=True, bnb_4bit_compute_dtype=self.dtype, bnb_4bit_use_double_quant=True,
                                   bnb_4bit_quant_type="nf4")
            llm = AutoModelForCausalLM.from_pretrained(llm_path, device_map="auto", quantization_config=q,
                                                       torch_dtype=self.dtype)
        else:
            llm = AutoModelForCausalLM.from_pretrained(llm_path, torch_dtype=self.dtype).to(self.device)
        if self.use_lora:
            try:
                from peft import LoraConfig, get_peft_model
            except ImportError as e:
                raise NotImplementedError("use_lora=True (the reference default) needs peft (clip_whisper_model.py:"
                                          "962-1002), which is not installed; pass use_lora=False or a _provided_llm "
                                          "that is already wrapped") from e
            targets = ["q_proj", "k_proj", "v_proj", "o_proj"] if "llama" in llm_path.lower() else \
                ["query", "key", "value", "dense"]
            llm = get_peft_model(llm, LoraConfig(r=self.lora_r, lora_alpha=self.lora_alpha,
                                                 lora_dropout=self.lora_dropout, bias="none", task_type="CAUSAL_LM",
                                                 target_modules=targets, init_lora_weights="gaussian",
                                                 fan_in_fan_out=False))
            with torch.no_grad():  # the reference scales the fresh LoRA weights down by 100 (:984-995)
                for n, prm in llm.named_parameters():
                    if "lora_" in n and prm.requires_grad:
                        prm.mul_(0.01)
        if self.freeze_llm:
            self._freeze_llm(llm, keep_lora=self.use_lora)
        return tok, llm

    @staticmethod
    def _freeze_llm(llm, keep_lora: bool):
        """freeze_llm (clip_whisper_model.py:1005-1014): everything frozen except, with LoRA, the adapter weights."""
        for n, prm in llm.named_parameters():
            prm.requires_grad = bool(keep_lora and "lora" in n)

    def _load_tower(self, cls_name, path):
        import transformers

        return getattr(transformers, cls_name).from_pretrained(path).to(self.device).eval()

    def _get_llm_dim(self):
        return self.llm.get_input_embeddings().weight.shape[1]  # clip_whisper_model.py:1148-1157

    def _setup_projections(self):
        """Both connectors are always created (clip_whisper_model.py:1159-1190); with stride-k stacking the input
        width is k * tower width."""
        self.audio_connector = create_modality_connector(
            connector_type=self.connector_type, input_dim=self.audio_dim * self.audio_stride,
            output_dim=self.llm_dim, device=self.device, dtype=self.dtype, max_seq_len=self.max_seq_len)
        self.video_connector = create_modality_connector(
            connector_type=self.connector_type, input_dim=self.video_dim * self.video_stride,
            output_dim=self.llm_dim, device=self.device, dtype=self.dtype, max_seq_len=self.max_seq_len)

    def _plan(self) -> FusePlan:
        return FusePlan(modality=self.modality, fusion=self.fusion, fusion_scale=self.fusion_scale,
                        max_seq_len=self.max_seq_len, audio_stride=self.audio_stride,
                        video_stride=self.video_stride, video_repeat=self.video_repeat, mask_mode=self.mask_mode,
                        label_mode=self.label_mode)

    def train(self, mode: bool = True):
        """Mode switches drop the forward-only bf16 weight-pack cache: an optimizer that updated the parameters through
        raw pointers between two evaluations did not bump torch's version counter (connector_ops.pack_projector)."""
        from .connector_ops import invalidate_pack_cache

        invalidate_pack_cache()
        return super().train(mode)

    def enable_data_parallel(self, process_group=None, multimem=None):
        """One process per GPU: route the projector gradients of `forward()`'s backward into a peer-mapped bucket that
        the dW GEMM launch all-reduces itself (parallel.FusedGradSync; insertion point between loss.backward() and
        clip_grad_norm_, clip_whisper_trainer.py:454-458).  The parameters' .grad become views of that bucket."""
        from .parallel import FusedGradSync

        if self.connector_type != "simple":
            raise NotImplementedError("the fused gradient all-reduce covers the linear projector ('simple')")
        use_a, use_v = self.modality in ("audio", "both"), self.modality in ("video", "both")
        ac, vc = self.audio_connector.linear, self.video_connector.linear
        self.grad_sync = FusedGradSync(ac.weight if use_a else None, ac.bias if use_a else None,
                                       vc.weight if use_v else None, vc.bias if use_v else None,
                                       process_group=process_group, multimem=multimem)
        return self.grad_sync

    # ------------------------------------------------------------------ towers (untouched PyTorch)
    def _whisper_features(self, audio, attention_mask=None):
        """Input checks and tower call of encode_audio (clip_whisper_model.py:1067-1103)."""
        if audio is None:
            raise ValueError("Audio input cannot be None")
        if len(audio.shape) != 2 and (len(audio.shape) != 3 or audio.shape[1] != 80):
            raise ValueError("Audio input should have shape [batch_size, sequence_length] or "
                             f"[batch_size, 80, time_steps], but got {audio.shape}")
        if len(audio.shape) == 3 and attention_mask is None:
            attention_mask = torch.ones((audio.shape[0], audio.shape[2]), dtype=torch.long, device=audio.device)
        whisper_dtype = next(self.whisper.parameters()).dtype
        if audio.dtype != whisper_dtype:
            audio = audio.to(whisper_dtype)
        with torch.set_grad_enabled(not self.freeze_encoders and torch.is_grad_enabled()):  # :1096
            out = self.whisper.encoder(audio, attention_mask=attention_mask, output_hidden_states=True,
                                       return_dict=True)
        return out.last_hidden_state

    def _clip_cls_features(self, video):
        """Input checks, tower call and CLS select of encode_video (clip_whisper_model.py:1108-1142).  The CLS rows
        are returned as a strided VIEW of last_hidden_state; the gather kernel reads them in place."""
        if len(video.shape) != 5 or video.shape[2] != 3:
            raise ValueError("Video input should have shape [batch_size, frames, 3, height, width], "
                             f"but got {video.shape}")
        B, F = video.shape[0], video.shape[1]
        flat = video.view(B * F, 3, video.shape[3], video.shape[4])
        clip_dtype = next(self.clip.parameters()).dtype
        if flat.dtype != clip_dtype:
            flat = flat.to(clip_dtype)
        with torch.set_grad_enabled(not self.freeze_encoders and torch.is_grad_enabled()):  # :1136
            hidden = self.clip(flat, return_dict=True).last_hidden_state  # [B*F, 1+Np, Dv]
        if not hidden.is_contiguous():
            hidden = hidden.contiguous()
        n1, Dv = hidden.shape[1], hidden.shape[2]
        return torch.as_strided(hidden, (B, F, Dv), (F * n1 * Dv, n1 * Dv, 1))

    def encode_audio(self, audio, attention_mask=None):
        """Whisper encoder -> audio connector: [B, Ta, H] (clip_whisper_model.py:1067-1106).  Stand-alone use
        requires audio_stride == 1 (the connector input is then the tower width, as in the reference)."""
        feats = self._whisper_features(audio, attention_mask)
        if self.audio_stride != 1:
            feats = _stack_view(feats, self.audio_stride)
        return self.audio_connector(feats)

    def encode_video(self, video, attention_mask=None):
        """CLIP -> CLS rows -> video connector: [B, F, H] (clip_whisper_model.py:1108-1146)."""
        feats = self._clip_cls_features(video)
        if self.video_stride != 1:
            feats = _stack_view(feats.contiguous(), self.video_stride)
        return self.video_connector(feats)

    # ------------------------------------------------------------------ the hot path
    def _prompt_ids(self, prompt):
        """Token ids of the prompt, capped at 32 (clip_whisper_model.py:464-482)."""
        if prompt is None:
            return None
        if isinstance(prompt, str) or (isinstance(prompt, (list, tuple)) and prompt and isinstance(prompt[0], str)):
            ids = self.tokenizer(prompt, return_tensors="pt", padding=True, truncation=True,
                                 max_length=MAX_PROMPT_LEN).input_ids
        else:
            ids = prompt
        ids = ids.to(self.device)
        return ids[:, :MAX_PROMPT_LEN]

    def _embed_prompt(self, prompt):
        ids = self._prompt_ids(prompt)
        return None if ids is None else self.llm.get_input_embeddings()(ids)

    def _fused(self, audio, video, prompt=None, labels=None, input_ids=None, placeholder_id=-1,
               audio_lengths=None, video_lengths=None, total_tokens=None):
        a = v = None
        if self.modality in ("audio", "both") and audio is not None:
            a = self._whisper_features(audio)
        if self.modality in ("video", "both") and video is not None:
            v = self._clip_cls_features(video)
        if a is None and v is None:
            raise ValueError("No valid inputs provided - both audio and video are None")
        llm_dtype = next(self.llm.parameters()).dtype  # clip_whisper_model.py:454
        table = self.llm.get_input_embeddings().weight
        pad_id = self.tokenizer.pad_token_id if self.tokenizer is not None else 0
        prompt_ids = self._prompt_ids(prompt)
        if self.connector_type == "adaptive":
            raise NotImplementedError("the 'adaptive' connector is a stand-alone module (encode_audio / encode_video); "
                                      "the fused encode() / forward() path needs connector_type 'simple' or 'mlp'")
        if self.connector_type == "mlp":
            out = fused_connector(
                a, v, None, None, None, None, self._plan(), input_ids=input_ids,
                prompt_ids=prompt_ids, placeholder_id=placeholder_id, embed_table=table.detach(),
                labels=labels, pad_id=pad_id, out_dtype=llm_dtype, audio_lengths=audio_lengths,
                video_lengths=video_lengths, total_tokens=total_tokens, mlp_audio=self.audio_connector.mlp_params(),
                mlp_video=self.video_connector.mlp_params())
        else:
            out = fused_connector(
                a, v, self.audio_connector.linear.weight, self.audio_connector.linear.bias,
                self.video_connector.linear.weight, self.video_connector.linear.bias, self._plan(),
                input_ids=input_ids, prompt_ids=prompt_ids, placeholder_id=placeholder_id,
                embed_table=table.detach(), labels=labels, pad_id=pad_id, out_dtype=llm_dtype,
                audio_lengths=audio_lengths, video_lengths=video_lengths, total_tokens=total_tokens,
                grad_sync=self.grad_sync if (self.training and torch.is_grad_enabled()) else None)
        if table.requires_grad and torch.is_grad_enabled():
            out = (self._differentiable_text_rows(out[0], table, input_ids, prompt_ids, placeholder_id),) + out[1:]
        return out

    def _differentiable_text_rows(self, emb, table, input_ids, prompt_ids, placeholder_id):
        """The splice kernel copies text rows from a detached table.  When the LLM's input embeddings are trainable
        (no LoRA / no freeze) the reference's `embedding_layer(prompt_ids)` (clip_whisper_model.py:484-485) is
        differentiable, so the text rows are re-written from the live embedding layer (same values, with a graph)."""
        layer = self.llm.get_input_embeddings()
        if input_ids is None:
            if prompt_ids is None:
                return emb
            P = prompt_ids.shape[1]
            return torch.cat([layer(prompt_ids).to(emb.dtype), emb[:, P:]], dim=1)
        ids = input_ids.to(emb.device)
        text = (ids != placeholder_id) & (ids >= 0) & (ids < table.shape[0])
        rows = layer(ids.clamp(0, table.shape[0] - 1)).to(emb.dtype)
        return torch.where(text.unsqueeze(-1), rows, emb)

    def encode(self, audio=None, video=None, prompt=None, input_ids=None, placeholder_id=-1,
               audio_lengths=None, video_lengths=None, total_tokens=None):
        """(inputs_embeds [B, P+T, H] in the LLM dtype, attention_mask int64 [B, P+T]) -- clip_whisper_model.py:407-462.

        Extension: `input_ids` containing `placeholder_id` runs selects the splice positions instead of the
        reference's fixed `[prompt | AV]` layout."""
        emb, mask, _ = self._fused(audio, video, prompt, None, input_ids, placeholder_id, audio_lengths,
                                   video_lengths, total_tokens)
        return emb, mask

    def forward(self, audio=None, video=None, prompt=None, labels=None, return_loss=True, input_ids=None,
                placeholder_id=-1, audio_lengths=None, video_lengths=None, total_tokens=None):
        """clip_whisper_model.py:489-619."""
        use_labels = labels is not None and return_loss
        if use_labels:
            labels = self._coerce_labels(labels)
        # eval rule (pad -> -100, truncate / right-pad) is produced by the splice kernel
        emb, mask, lab = self._fused(audio, video, prompt, labels if use_labels else None, input_ids,
                                     placeholder_id, audio_lengths, video_lengths, total_tokens)
        if use_labels and self.training and labels.shape[1] != emb.shape[1]:
            # training branch (:577-585): sequence is resampled to the label length, labels only get pad -> -100
            emb = adaptive_projection(emb, labels.shape[1])
            mask = adapt_mask(mask, labels.shape[1])
            lab = labels.clone()
            lab[lab == self.tokenizer.pad_token_id] = -100
        if use_labels:
            outputs = self.llm(inputs_embeds=emb, attention_mask=mask, labels=lab, return_dict=True)
        else:
            outputs = self.llm(inputs_embeds=emb, attention_mask=mask, return_dict=True)
        if return_loss:
            return {"loss": outputs.loss, "logits": outputs.logits}
        return {"logits": outputs.logits}

    def _coerce_labels(self, labels):
        """List -> tensor coercions of clip_whisper_model.py:505-567 (strings go through the tokenizer)."""
        if isinstance(labels, list):
            if all(isinstance(x, torch.Tensor) for x in labels):
                labels = torch.stack(labels)
            elif all(isinstance(x, str) for x in labels):
                labels = self.tokenizer(labels, return_tensors="pt", padding=True, truncation=True,
                                        max_length=self.max_seq_len).input_ids
            else:
                labels = torch.tensor(labels)
        if not isinstance(labels, torch.Tensor):
            raise TypeError(f"labels must be a tensor or a list, got {type(labels).__name__}")
        return labels.to(self.device)

    @torch.no_grad()
    def generate(self, audio=None, video=None, prompt=None, pixel_values=None, max_new_tokens=100, do_sample=False,
                 temperature=1.0, top_p=0.9, max_length=None):
        """clip_whisper_model.py:1240-1348: modality follows the inputs that are present, then llm.generate."""
        if video is None and pixel_values is not None:
            video = pixel_values
        if max_new_tokens is None:
            max_new_tokens = max_length if max_length is not None else 100
        original = self.modality
        if audio is not None and video is not None:
            self.modality = "both"
        elif audio is not None:
            self.modality = "audio"
        elif video is not None:
            self.modality = "video"
        try:
            emb, mask = self.encode(audio, video, prompt)
        finally:
            self.modality = original
        return self.llm.generate(inputs_embeds=emb, attention_mask=mask, max_new_tokens=max_new_tokens,
                                 do_sample=do_sample, temperature=temperature, top_p=top_p)

    # ------------------------------------------------------------------ checkpoint compatibility
    def save_connectors(self, out_dir):
        """audio_connector.pt / video_connector.pt with the reference's keys (clip_whisper_model.py:745-746)."""
        import os

        os.makedirs(out_dir, exist_ok=True)
        torch.save(self.audio_connector.state_dict(), os.path.join(out_dir, "audio_connector.pt"))
        torch.save(self.video_connector.state_dict(), os.path.join(out_dir, "video_connector.pt"))

    def load_connectors(self, src_dir):
        import os

        for name in ("audio_connector", "video_connector"):
            sd = torch.load(os.path.join(src_dir, f"{name}.pt"), map_location=self.device)
            getattr(self, name).load_state_dict({k: v.float() for k, v in sd.items()})
        logging.info("loaded connector weights from %s", src_dir)


def _stack_view(feats: torch.Tensor, k: int) -> torch.Tensor:
    """[B, T, D] -> [B, T // k, k*D] (frames k*j .. k*j+k-1 side by side); T must divide by k for the free view."""
    B, T, D = feats.shape
    if T % k:
        raise ValueError(f"{T} frames do not divide by stride {k}; use encode() (the gather pads the ragged tail)")
    return feats.reshape(B, T // k, k * D)
