"""Forward-only (decode / generate) latency of the connector through the public API, batch 1..8
(inference needs fwd only, latency-bound).  One JSON line per case."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import audio_visual_llm_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
pkg._lib.require_device(0)
g = torch.Generator(device="cuda").manual_seed(0)
H, Da, Dv, V, P = 4096, 1024, 1024, 32000, 16
table = (torch.randn(V, H, generator=g, device=dev) * 0.02).bfloat16()
for name, ka, kv, fusion in (("stride4_concat", 4, 2, "concat"), ("parity_k1_sum", 1, 1, "sum")):
    wa = torch.randn(H, ka * Da, generator=g, device=dev) * 0.02
    wv = torch.randn(H, kv * Dv, generator=g, device=dev) * 0.02
    ba, bv = torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    plan = pkg.FusePlan(fusion=fusion, audio_stride=ka, video_stride=kv, max_seq_len=1536)
    for B in (1, 2, 4, 8):
        a = torch.randn(B, 1500, Da, generator=g, device=dev).bfloat16()
        v = torch.randn(B, 750, Dv, generator=g, device=dev).bfloat16()
        prompt = torch.randint(1, V, (B, P), generator=g, device=dev)

        def run():
            with torch.no_grad():
                return pkg.fused_connector(a, v, wa, ba, wv, bv, plan, prompt_ids=prompt, embed_table=table,
                                           out_dtype=torch.bfloat16)

        for _ in range(10):
            run()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(50):
            emb, mask, _ = run()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 50
        print(json.dumps({"case": name, "batch": B, "fused_tokens": emb.shape[1] - P, "ms_per_encode": round(ms, 4),
                          "tokens_per_s": round(B * (emb.shape[1] - P) / ms * 1e3)}), flush=True)
