"""Is the forward GEMM's operand wait a DRAM-latency effect?  The same N = 4096, K = 2048 problem with an activation
matrix that fits in L2 (M = 9472: every iteration after the first hits L2) and one that streams from HBM (M = 37888),
both exact multiples of the 74 x 16 tile wave; ours next to cuBLAS, TFLOP/s in short back-to-back loops."""
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
H = 4096
for M, K in [(9472, 2048), (37888, 2048), (9472, 6144), (37888, 6144)]:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(H, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    Y = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
    for name, fn in (("ours", lambda: L.proj_fwd([A], [W], Y)), ("cublas", lambda: torch.matmul(A, W.t(), out=Y))):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 200
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / n
        print(json.dumps({"M": M, "K": K, "A_MB": round(M * K * 2 / 1e6, 1), "impl": name, "ms": round(ms, 4),
                          "TFLOPs": round(2.0 * M * K * H / ms / 1e9, 1)}), flush=True)
        torch.cuda.synchronize()
    del A, W, Y
