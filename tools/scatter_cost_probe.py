"""Forward GEMM at cfg2 with the scatter epilogue (projected rows written straight into the AV region of inputs_embeds,
32-row boxes that straddle a sample boundary stored row by row) against the packed output, and the whole engine step
with / without the side-stream kernels -- where does the in-step forward lose time against the isolated kernel?"""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import audio_visual_llm_b200 as pkg  # noqa: E402
from audio_visual_llm_b200.engine import ConnectorStep, StepShape  # noqa: E402

L = pkg._lib
dev = torch.device("cuda:0")
B, N, P, H, Ka, Kv = 32, 375, 16, 4096, 4096, 2048
xa = torch.randn(B * N, Ka, device=dev).to(torch.bfloat16)
xv = torch.randn(B * N, Kv, device=dev).to(torch.bfloat16)
W = (torch.randn(H, Ka + Kv, device=dev) / 78.0).to(torch.bfloat16)
b0, b1 = torch.randn(H, device=dev), torch.randn(H, device=dev)
emb = torch.empty(B, P + N, H, dtype=torch.bfloat16, device=dev)
Y = torch.empty(B * N, H, dtype=torch.bfloat16, device=dev)


def bench(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return round(s.elapsed_time(e) / n, 4)


print(json.dumps({"fwd_packed_ms": bench(lambda: L.proj_fwd([xa, xv], [W[:, :Ka], W[:, Ka:]], Y, bias0=b0, bias1=b1)),
                  "fwd_scatter_ms": bench(lambda: L.proj_fwd([xa, xv], [W[:, :Ka], W[:, Ka:]], emb[:, P:, :], bias0=b0, bias1=b1))}))
shape = StepShape(batch=B, audio_frames=1500, video_frames=750, audio_dim=1024, video_dim=1024, hidden=H)
plan = pkg.FusePlan(fusion="concat", audio_stride=4, video_stride=2, max_seq_len=4096)
for side in ("1", "0"):
    os.environ["AVC_SIDE_STREAMS"] = side
    eng = ConnectorStep(shape, plan, dev, seed=1)
    ms = bench(eng.step, 100)
    eng.enable_kernel_timing()
    for _ in range(20):
        eng.step()
    torch.cuda.synchronize()
    k = {n: round(sum(s.elapsed_time(e) for s, e in ev) / len(ev), 4) for n, ev in eng.events.items() if ev}
    eng.events = None
    g = eng.capture_graph()
    print(json.dumps({"side_streams": side, "step_ms": ms, "graphed_step_ms": bench(g.replay, 100), "kernel_ms": k}))
    del eng
