"""dW GEMM with and without the bias work items at a single-round shape (cfg3: 40 tiles + 8 bias items on 74 workers),
with the per-CTA counters of the pairs that ran a regular tile and of those that ran a bias item."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import audio_visual_llm_b200 as pkg  # noqa: E402

L = pkg._lib
dev = torch.device("cuda:0")
B, N, P, H, K = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "64,1500,16,4096,1280").split(","))
dy = torch.randn(B, P + N, H, device=dev).to(torch.bfloat16)
x = torch.randn(B, N, K, device=dev).to(torch.bfloat16)
dw = torch.empty(H, K, device=dev)
db = torch.empty(H, device=dev)
present = L.present_operand(B, N, dev)


def plain():
    L.proj_bwd_dw(dy, [x], [dw], [1.0], dy_row_base=P)


def with_db():
    L.proj_bwd_dw(dy, [x], [dw], [1.0], dy_row_base=P, bias=(present, db, None, 1.0, 1.0))


for name, fn in (("plain", plain), ("with_db", with_db)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        fn()
    e.record()
    torch.cuda.synchronize()
    prof = torch.zeros(3 * 148, 8, dtype=torch.int64, device=dev)
    L.debug_gemm_profile(prof)
    fn()
    torch.cuda.synchronize()
    L.debug_gemm_profile(None)
    p = prof.cpu().double()[:148]
    lead = p[0::2]
    rows = []
    for w in range(74):
        rows.append((w, int(lead[w, 5]), round(float(lead[w, 6]) / 1e3), round(float(lead[w, 0]) / 1e3),
                     round(float(lead[w, 1]) / 1e3)))
    ts = prof.cpu()[296:].double()
    t0 = ts[:, 0][ts[:, 0] > 0].min()
    stamp = lambda c: round(float((ts[:, c][ts[:, c] > 0].max() - t0) / 1e3), 1)
    first = lambda c: round(float((ts[:, c][ts[:, c] > 0].min() - t0) / 1e3), 1)
    print(json.dumps({"case": name, "ms": round(s.elapsed_time(e) / 20, 4), "plan": L.dw_plan(dy, [x], name == "with_db"),
                      "us_since_launch": {"first_cta_start": 0.0, "last_cta_start": stamp(0), "last_mma_commit": stamp(5),
                                          "first_epilogue_end": first(1), "last_epilogue_end": stamp(1), "kernel_end": stamp(7)}}))
    print("  worker: items, producer loop kc, producer wait-empty kc, MMA wait-full kc")
    for r in rows[:3] + rows[38:50] + rows[72:]:
        print("   ", r)
