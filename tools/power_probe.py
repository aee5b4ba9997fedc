"""Sustained (>= 1.5 s) back-to-back loops of the projector GEMMs and of cuBLAS at the BASELINE cfg2 shapes, with
nvidia-smi power / SM clock sampled during each loop: TFLOP/s, W and MHz side by side.  Kernel variants are selected
through the AVC_GEMM_* environment knobs, which the library reads at every launch.
Usage: python tools/power_probe.py [seconds]"""
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import audio_visual_llm_b200 as pkg  # noqa: E402

L = pkg._lib
dev = torch.device("cuda:0")
SECS = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5
B, N, H, Ka, Kv = 32, 375, 4096, 4096, 2048
M, K = B * N, Ka + Kv
A = torch.randn(M, K, device=dev).to(torch.bfloat16)
W = (torch.randn(H, K, device=dev) / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(H, device=dev)
Y = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
dY = torch.randn(M, H, device=dev).to(torch.bfloat16)
dW = torch.empty(H, K, dtype=torch.float32, device=dev)
dWb = torch.empty(H, K, dtype=torch.bfloat16, device=dev)
present = L.present_operand(1, M, dev)
db0, db1 = torch.empty(H, device=dev), torch.empty(H, device=dev)
FLOPS = 2.0 * M * K * H


class Smi:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader,nounits",
                                   "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.p.stdout:
            try:
                w, mhz = (float(x) for x in line.split(","))
                self.rows.append((time.time(), w, mhz))
            except ValueError:
                pass

    def window(self, t0, t1):
        r = [x for x in self.rows if t0 + 0.3 <= x[0] <= t1]
        if not r:
            return None, None
        return sum(x[1] for x in r) / len(r), sum(x[2] for x in r) / len(r)


smi = Smi()
time.sleep(0.5)


def probe(name, fn, env=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        n = 0
        t0 = time.time()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        while time.time() - t0 < SECS:
            for _ in range(100):
                fn()
            n += 100
            torch.cuda.synchronize()
        e.record()
        torch.cuda.synchronize()
        t1 = time.time()
        ms = s.elapsed_time(e) / n
        w, mhz = smi.window(t0, t1)
        print(json.dumps({"kernel": name, "env": env or {}, "ms": round(ms, 4), "TFLOPs": round(FLOPS / ms / 1e9, 1),
                          "watts": None if w is None else round(w, 1), "sm_mhz": None if mhz is None else round(mhz),
                          "GFLOP_per_J": None if not w else round(FLOPS / (ms * 1e-3) / w / 1e9, 1)}), flush=True)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    time.sleep(1.0)  # let the board cool between probes


def fwd():
    L.proj_fwd([A[:, :Ka], A[:, Ka:]], [W[:, :Ka], W[:, Ka:]], Y, bias0=bias, bias1=bias)


def dw():
    L.proj_bwd_dw(dY, [A], [dW], [1.0], bias=(present, db0, db1, 1.0, 1.0))


variants = [v for v in os.environ.get("PROBE_VARIANTS", "").split(";") if v]
probe("cublas_fwd", lambda: torch.matmul(A, W.t(), out=Y))
probe("cublas_dw_bf16out", lambda: torch.matmul(dY.t(), A, out=dWb))
probe("proj_fwd", fwd)
probe("proj_bwd_dw_db", dw)
for v in variants:
    env = dict(kv.split("=") for kv in v.split())
    probe("proj_fwd", fwd, env)
    if any(k.endswith("_NT") or k in ("AVC_GEMM_PREFETCH", "AVC_GEMM_GROUP_M") for k in env):
        probe("proj_bwd_dw_db", dw, env)
smi.p.terminate()
