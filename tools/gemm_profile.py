"""Where a projector GEMM launch spends its cycles, from the kernel's own per-CTA counters (avc_debug_gemm_profile):
producer / MMA issuer / epilogue wait and work time at the BASELINE cfg2 shapes.  Usage: python tools/gemm_profile.py"""
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
B, N, P, H, Ka, Kv = 32, 375, 16, 4096, 4096, 2048
M, K = B * N, Ka + Kv
if len(sys.argv) > 1:   # M,Ka,Kv,H of another shape, e.g. cfg3: 96000,1280,0,4096
    M, Ka, Kv, H = (int(x) for x in sys.argv[1].split(","))
    K = Ka + Kv
A = torch.randn(M, K, device=dev).to(torch.bfloat16)
W = (torch.randn(H, K, device=dev) / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(H, device=dev)
Y = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
dY = torch.randn(M, H, device=dev).to(torch.bfloat16)
dW = torch.empty(H, K, dtype=torch.float32, device=dev)
prof = torch.zeros(3 * 148, 8, dtype=torch.int64, device=dev)


def run(name, fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    prof.zero_()
    L.debug_gemm_profile(prof)
    fn()
    torch.cuda.synchronize()
    L.debug_gemm_profile(None)
    pall = prof.cpu().double()
    p, ph = pall[:148], pall[148:296]
    lead = p[0::2]   # even CTAs lead their pair (MMA issuer counters live there)
    tiles = p[:, 5].clamp_min(1)
    rec = {"kernel": name, "ms": round(ms, 4), "kcycles_kernel(producer loop)": round(float(p[:, 6].mean()) / 1e3, 1),
           "producer_wait_empty_kc": round(float(p[:, 0].mean()) / 1e3, 1),
           "mma_wait_full_kc": round(float(lead[:, 1].mean()) / 1e3, 1),
           "mma_wait_tempty_kc": round(float(lead[:, 2].mean()) / 1e3, 1),
           "epi_wait_tfull_kc": round(float(p[:, 3].mean()) / 1e3, 1),
           "epi_body_kc": round(float(p[:, 4].mean()) / 1e3, 1), "items_per_cta": round(float(p[:, 5].mean()), 2),
           "epi_body_kc_per_item": round(float((p[:, 4] / tiles).mean()) / 1e3, 2),
           "epi_phase_kc_per_item[wait_buf,ldtm,math_sts,fence,tma_store]":
               [round(float((ph[:, i] / tiles).mean()) / 1e3, 2) for i in range(5)]}
    print(json.dumps(rec), flush=True)


if Kv:
    run("proj_fwd", lambda: L.proj_fwd([A[:, :Ka], A[:, Ka:]], [W[:, :Ka], W[:, Ka:]], Y, bias0=bias, bias1=bias))
    run("proj_bwd_dw", lambda: L.proj_bwd_dw(dY, [A[:, :Ka], A[:, Ka:]], [dW[:, :Ka], dW[:, Ka:]], [1.0, 1.0]))
else:
    run("proj_fwd", lambda: L.proj_fwd([A], [W], Y, bias0=bias))
    run("proj_bwd_dw", lambda: L.proj_bwd_dw(dY, [A], [dW], [1.0]))
for name, fn in (("cublas_fwd", lambda: torch.matmul(A, W.t(), out=Y)),
                 ("cublas_dw", lambda: torch.matmul(dY.t(), A))):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        fn()
    e.record()
    torch.cuda.synchronize()
    print(json.dumps({"kernel": name, "ms": round(s.elapsed_time(e) / 20, 4)}), flush=True)
