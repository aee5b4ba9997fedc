"""TEST INFRASTRUCTURE ONLY -- times the oracle port of the reference connector on the host CPU.

This is the `cpu_baseline` leg of bench.py and its `--impl reference` arm (kind "port": the reference is
Python/PyTorch and cannot travel to the GPU box, so its restatement in oracle/connector_oracle.py -- pinned
against the executing reference by tests/test_oracle_golden.py -- is what runs).  fp32, all host threads, the
same torch CPU kernels (`F.linear` -> oneDNN/MKL sgemm, autograd) the reference's connector dispatches to.
One step = connector fwd (stack -> project -> fuse -> splice + masks) + bwd (dW, db) for `batch` samples.
"""
from __future__ import annotations

import os
import time

import torch

from . import connector_oracle as O


def make_case(workload: dict, batch: int, seed: int = 1234):
    g = torch.Generator().manual_seed(seed)
    H = workload["hidden"]
    use_a, use_v = workload["modality"] in ("audio", "both"), workload["modality"] in ("video", "both")
    ka, kv = workload["audio_stride"], workload["video_stride"]
    a = torch.randn(batch, workload["audio_frames"], workload["audio_dim"], generator=g) if use_a else None
    v = torch.randn(batch, workload["video_frames"], workload["video_dim"], generator=g) if use_v else None
    Ka, Kv = ka * workload["audio_dim"], kv * workload["video_dim"]
    wa = torch.randn(H, Ka, generator=g) * (6.0 / (H + Ka)) ** 0.5
    wv = torch.randn(H, Kv, generator=g) * (6.0 / (H + Kv)) ** 0.5
    ba, bv = torch.randn(H, generator=g) * 0.02, torch.randn(H, generator=g) * 0.02
    spec = O.ConnectorSpec(modality=workload["modality"], fusion=workload["fusion"],
                           fusion_scale=workload["fusion_scale"], max_seq_len=workload["max_seq_len"],
                           audio_stride=ka, video_stride=kv)
    P = workload["prompt_len"]
    prompt = torch.randint(1, 1000, (batch, P), generator=g)
    table = torch.randn(1001, H, generator=g) * 0.02
    labels = torch.randint(1, 1000, (batch, 256), generator=g)
    N = O.token_counts(spec, workload["audio_frames"] if use_a else None, workload["video_frames"] if use_v else None)
    up = torch.randn(batch, P + N, H, generator=g)
    return dict(a=a, v=v, params=[wa, ba, wv, bv], spec=spec, prompt=prompt, table=table, labels=labels, up=up,
                tokens=batch * N)


def one_step(case):
    params = [p.detach().requires_grad_(True) for p in case["params"]]
    emb, mask, lab, _ = O.connector_forward(case["a"], case["v"], *params, case["spec"], prompt_ids=case["prompt"],
                                            embed_table=case["table"], labels=case["labels"])
    (emb * case["up"]).sum().backward()
    return emb, mask, lab, [p.grad for p in params]


def time_cpu(workload: dict, batch: int, steps: int, warmup: int):
    """Returns (fused tokens/s, seconds per step, threads)."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    case = make_case(workload, batch)
    for _ in range(warmup):
        one_step(case)
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step(case)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return case["tokens"] / dt, dt, threads
