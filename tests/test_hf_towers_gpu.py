"""Drop-in check with REAL Hugging Face tower classes (tiny random-init configs, no downloads): the B200
ClipWhisperModel calls WhisperModel.encoder / CLIPVisionModel / LlamaForCausalLM exactly where the reference does
(clip_whisper_model.py:1098-1103, 1138, 601-613, 1337-1340) and its outputs match the oracle applied to the same
tower outputs."""
from types import SimpleNamespace

import pytest
import torch

from oracle import connector_oracle as O

pytestmark = pytest.mark.gpu


def tiny_towers(dev):
    from transformers import (CLIPVisionConfig, CLIPVisionModel, LlamaConfig, LlamaForCausalLM, WhisperConfig,
                              WhisperModel)

    torch.manual_seed(0)
    whisper = WhisperModel(WhisperConfig(d_model=64, encoder_layers=1, decoder_layers=1, encoder_attention_heads=2,
                                         decoder_attention_heads=2, encoder_ffn_dim=128, decoder_ffn_dim=128,
                                         num_mel_bins=80, max_source_positions=1500, vocab_size=128, pad_token_id=0,
                                         bos_token_id=1, eos_token_id=2, decoder_start_token_id=1)).to(dev).eval()
    clip = CLIPVisionModel(CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=1,
                                            num_attention_heads=2, image_size=32, patch_size=16)).to(dev).eval()
    llm = LlamaForCausalLM(LlamaConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=1,
                                       num_attention_heads=2, num_key_value_heads=2, vocab_size=96,
                                       max_position_embeddings=4096, pad_token_id=0, bos_token_id=1,
                                       eos_token_id=2)).to(dev)
    tok = SimpleNamespace(pad_token_id=0)
    return whisper, clip, llm, tok


@pytest.mark.parametrize("mode", ["both_parity", "both_stride"])
def test_model_with_real_hf_towers(avc, cuda_dev, mode):
    dev = cuda_dev
    whisper, clip, llm, tok = tiny_towers(dev)
    kw = dict(max_seq_len=1536) if mode == "both_parity" else dict(max_seq_len=1536, fusion="concat", stride=4,
                                                                  align="rate")
    m = avc.ClipWhisperModel(device="cuda:0", modality="both", fusion_scale=0.4, _provided_tokenizer=tok,
                             _provided_llm=llm, _provided_whisper=whisper, _provided_clip=clip, **kw)
    assert (m.audio_dim, m.video_dim, m.llm_dim) == (64, 32, 128)
    g = torch.Generator().manual_seed(3)
    B, F = 2, 8
    audio = torch.randn(B, 80, 3000, generator=g).to(dev)          # log-mel, as the dataset produces
    video = torch.randn(B, F, 3, 32, 32, generator=g).to(dev)
    prompt = torch.randint(1, 96, (B, 5), generator=g).to(dev)
    labels = torch.randint(0, 96, (B, 40), generator=g).to(dev)
    with torch.no_grad():
        for c in (m.audio_connector, m.video_connector):
            c.linear.bias.normal_(0, 0.05)
    m.eval()
    emb, mask = m.encode(audio, video, prompt)
    # oracle on the tower outputs the model saw
    with torch.no_grad():
        a_feats = whisper.encoder(audio, return_dict=True).last_hidden_state.float().cpu()
        v_hidden = clip(video.view(B * F, 3, 32, 32), return_dict=True).last_hidden_state.float().cpu()
    v_feats = O.reference_cls_select(v_hidden, B, F)
    spec = O.ConnectorSpec(fusion=m.fusion, fusion_scale=0.4, max_seq_len=1536, audio_stride=m.audio_stride,
                           video_stride=m.video_stride)
    p = [t.detach().float().cpu() for t in (m.audio_connector.linear.weight, m.audio_connector.linear.bias,
                                            m.video_connector.linear.weight, m.video_connector.linear.bias)]
    emb_r, mask_r, lab_r, _ = O.connector_forward(a_feats, v_feats, *p, spec, prompt_ids=prompt.cpu(),
                                                  embed_table=llm.get_input_embeddings().weight.detach().float().cpu(),
                                                  labels=labels.cpu())
    assert emb.shape == emb_r.shape and emb.dtype == torch.float32
    got, ref = emb.double().cpu(), emb_r.double()
    assert float((got - ref).abs().max() / ref.abs().max()) <= 1e-2
    assert float(torch.dot(got.flatten(), ref.flatten()) / (got.norm() * ref.norm())) >= 0.9999
    assert torch.equal(mask.cpu(), mask_r)
    # forward with loss through the real LLM, backward into the connectors, generate
    out = m(audio=audio, video=video, prompt=prompt, labels=labels)
    assert torch.isfinite(out["loss"])
    out["loss"].backward()
    for c in (m.audio_connector, m.video_connector):
        assert c.linear.weight.grad is not None and torch.isfinite(c.linear.weight.grad).all()
        assert float(c.linear.weight.grad.abs().max()) > 0
    ids = m.generate(audio=audio, video=video, prompt=prompt, max_new_tokens=3)
    assert ids.shape[0] == B
    # single-modality generate switches the modality like the reference (clip_whisper_model.py:1280-1294)
    ids = m.generate(audio=audio, max_new_tokens=2)
    assert ids.shape[0] == B and m.modality == "both"
