"""Forward-only (decode / generate) latency of the connector through the public API, batch 1..8
(inference needs fwd only, latency-bound).  One JSON line per case."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import audio_visual_llm_b200 as pkg  # noqa: E402
from audio_visual_llm_b200.engine import GraphedEncoder  # noqa: E402

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

        graphed = GraphedEncoder(lambda a_, v_, p_: pkg.fused_connector(a_, v_, wa, ba, wv, bv, plan, prompt_ids=p_,
                                                                        embed_table=table, out_dtype=torch.bfloat16))

        def timed(fn, n=200):
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            s.record()
            for _ in range(n):
                out = fn()
            e.record()
            host_ms = (time.perf_counter() - t0) / n * 1e3   # host time to ENQUEUE one encode (the launch-bound part)
            torch.cuda.synchronize()
            return s.elapsed_time(e) / n, host_ms, out

        ms, host_ms, (emb, mask, _) = timed(run)
        gms, ghost_ms, (gemb, gmask, _) = timed(lambda: graphed(a, v, prompt))
        assert torch.equal(gemb, emb) and torch.equal(gmask, mask), "graph replay must give the eager call's bits"
        print(json.dumps({"case": name, "batch": B, "fused_tokens": emb.shape[1] - P, "ms_per_encode": round(ms, 4),
                          "host_ms_per_encode": round(host_ms, 4), "graphed_ms_per_encode": round(gms, 4),
                          "graphed_host_ms": round(ghost_ms, 4),
                          "tokens_per_s": round(B * (emb.shape[1] - P) / ms * 1e3),
                          "graphed_tokens_per_s": round(B * (emb.shape[1] - P) / gms * 1e3)}), flush=True)
