"""Back-to-back fwd + dW GEMM pairs for ~0.3 s (the power-capped regime of the training step): ours vs cuBLAS."""
import json, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry; entry.build()
import audio_visual_llm_b200 as pkg
L = pkg._lib
dev = torch.device("cuda:0")
M, K, H = 12000, 6144, 4096
A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(H, K, device=dev) / 78).bfloat16()
Y = torch.empty(M, H, dtype=torch.bfloat16, device=dev); dY = torch.randn(M, H, device=dev).bfloat16()
dW = torch.empty(H, K, dtype=torch.float32, device=dev); dWb = torch.empty(H, K, dtype=torch.bfloat16, device=dev)
bias = torch.zeros(H, device=dev)
def ours():
    L.proj_fwd([A], [W], Y, bias0=bias); L.proj_bwd_dw(dY, [A], [dW], [1.0])
def cublas():
    torch.matmul(A, W.t(), out=Y); torch.matmul(dY.t(), A, out=dWb)
for name, fn in (("ours", ours), ("cublas", cublas), ("ours", ours), ("cublas", cublas)):
    for _ in range(50): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 300
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    print(json.dumps({"impl": name, "ms_per_pair": round(ms, 4), "TFLOPs": round(4 * M * K * H / ms / 1e9, 1)}), flush=True)
    time.sleep(0.5)
