#!/usr/bin/env python
"""Connector fwd+bwd benchmark (BASELINE.json metric: fused tokens/sec).

    python bench.py [--gpus N --steps K --warmup W]            our arm (sm_100a kernels)
    python bench.py --impl reference [...]                     reference arm: the oracle port on host CPU cores
    python bench.py --config cfg1|cfg2|cfg2p|cfg3|cfg3k4|cfg4  the other BASELINE.json configs (default cfg2)

Workload at every N: BASELINE.json configs[1] per GPU (weak scaling) --
Whisper-medium(1024) + CLIP ViT-L/14(1024) -> 4096, concat fusion, stride 4 (k_a=4 audio + k_v=2 video frames
per token, rate-aligned), batch 32, 30 s clips, bf16.  One step = gather -> projector GEMM -> splice(+masks) ->
splice-bwd -> dW GEMM (+ db) [-> projector-grad all-reduce when N > 1]; synthetic N(0,1) features and
random-init weights (no datasets / checkpoints offline).

Prints ONE JSON line (rank 0).  `value` has inputs resident in HBM (the driver-sized region: K steps); `sustained` is
the same loop over >= 1 s (power-capped clocks); `e2e` runs the public API (`fused_connector` + backward) from pinned
HOST tensors with the H2D copy of the step's inputs and a D2H read of its results inside the timed region;
`gpu_eager` is the reference connector's semantics in eager PyTorch (cuBLAS) on the same GPU; `self_check` compares
sampled outputs / gradients of the benchmarked engine with an fp64 recomputation.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

_COMMON = dict(fusion_scale=0.5, prompt_len=16)
CONFIGS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case (parity knobs: k = 1, index-aligned, sum fusion)
    "cfg1": dict(workload="cfg1: Whisper-small(768)+CLIP ViT-B/32(512)->2048, sum fusion, k=1, batch 2/GPU, 10 s, bf16",
                 modality="both", fusion="sum", max_seq_len=512, audio_stride=1, video_stride=1, audio_frames=500,
                 video_frames=250, audio_dim=768, video_dim=512, hidden=2048, batch_per_gpu=2, **_COMMON),
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "cfg2": dict(workload="cfg2: Whisper-medium(1024)+CLIP ViT-L/14(1024)->4096, concat, stride 4 (k_a=4,k_v=2), batch 32/GPU, 30 s, bf16",
                 modality="both", fusion="concat", max_seq_len=1536, audio_stride=4, video_stride=2, audio_frames=1500,
                 video_frames=750, audio_dim=1024, video_dim=1024, hidden=4096, batch_per_gpu=32, **_COMMON),
    # cfg2': the parity variant of cfg2 the reference code itself can run (k = 1, sum fusion, max_seq_len 1536)
    "cfg2p": dict(workload="cfg2p: cfg2 shapes at the reference's knobs (k=1, index-aligned, sum fusion, max_seq_len 1536), batch 32/GPU, bf16",
                  modality="both", fusion="sum", max_seq_len=1536, audio_stride=1, video_stride=1, audio_frames=1500,
                  video_frames=750, audio_dim=1024, video_dim=1024, hidden=4096, batch_per_gpu=32, **_COMMON),
    # BASELINE.json configs[2]
    "cfg3": dict(workload="cfg3: audio-only, Whisper-large-v3(1280)->4096, k=1, 30 s, batch 64/GPU, bf16",
                 modality="audio", fusion="sum", max_seq_len=1536, audio_stride=1, video_stride=1, audio_frames=1500,
                 video_frames=0, audio_dim=1280, video_dim=1024, hidden=4096, batch_per_gpu=64, **_COMMON),
    "cfg3k4": dict(workload="cfg3k4: audio-only, Whisper-large-v3(1280)->4096, stride 4, 30 s, batch 64/GPU, bf16",
                   modality="audio", fusion="sum", max_seq_len=1536, audio_stride=4, video_stride=1, audio_frames=1500,
                   video_frames=0, audio_dim=1280, video_dim=1024, hidden=4096, batch_per_gpu=64, **_COMMON),
    # BASELINE.json configs[3]: ragged -- the stand-alone gather and the ragged splice are on the product path here
    "cfg4": dict(workload="cfg4: video-only, CLIP ViT-L/14(1024) 25 fps 16 s ->4096, batch 64/GPU, variable-length placeholders N_i~U{100..400}, bf16",
                 modality="video", fusion="sum", max_seq_len=1536, audio_stride=1, video_stride=1, audio_frames=0,
                 video_frames=400, audio_dim=1024, video_dim=1024, hidden=4096, batch_per_gpu=64,
                 ragged_range=(100, 400), **_COMMON),
}
WORKLOAD = CONFIGS["cfg2"]
METRIC = "connector fused tokens/sec fwd+bwd"
UNIT = "fused tokens/s"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms.  Started BEFORE the warm-up (nvidia-smi takes a few
    hundred ms to produce its first line); `window(t0, t1)` summarises the samples of one timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.fast, self._fast_stop = [], threading.Event()   # NVML samples every ~2 ms: (time, sm MHz, reason bits)

    def _nvml_loop(self):
        """The same counters straight from NVML every ~2 ms, so that the driver-sized 17 ms region holds more than one
        sample; nvidia-smi (20 ms) stays the record when NVML is unavailable."""
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._fast_stop.is_set():
                self.fast.append((time.time(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(reasons(h))))
                time.sleep(0.002)
        except Exception:
            return

    def start(self):
        threading.Thread(target=self._nvml_loop, daemon=True).start()
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 6:
                self.rows.append((time.time(), parts))

    def wait_first_sample(self, timeout_s: float = 5.0):
        t = time.time()
        while self.proc is not None and not self.rows and time.time() - t < timeout_s:
            time.sleep(0.01)

    def stop(self):
        self._fast_stop.set()
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = list(self.rows)
        inside = [r for r in rows if t0 <= r[0] <= t1 + 0.03]
        window = "timed region"
        if not inside and rows:  # region shorter than the sampling period: take the samples closest to it
            mid = 0.5 * (t0 + t1)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            window = "nearest samples to the timed region"
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, r in inside:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
               "samples": len(sm), "window": window}
        fast = [f for f in list(self.fast) if t0 <= f[0] <= t1]
        if len(fast) >= 3:
            # NVML reason bits: 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal_slowdown, 0x40 hw_thermal_slowdown
            bits = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}
            clk = sorted(f[1] for f in fast)
            seen = set()
            for f in fast:
                seen |= {n for b, n in bits.items() if f[2] & b}
            out["nvml"] = {"sm_mhz": clk[len(clk) // 2], "sm_mhz_min": clk[0], "samples": len(clk),
                           "reasons": sorted(seen)}
            if window != "timed region" or len(sm) < 3:
                # too few nvidia-smi lines inside a short region: the NVML samples ARE inside it
                out.update(sm_mhz=clk[len(clk) // 2], reasons=sorted(set(out["reasons"]) | seen if window == "timed region"
                                                                      else seen),
                           samples=len(clk), window="timed region (NVML, 2 ms period)")
        return out


def reference_arm(args, rank: int):
    """The reference's CPU connector (oracle port) on the host cores; each step is a bounded sample of the config."""
    if rank != 0:
        return
    from oracle import cpu_baseline

    w = CONFIGS[args.config]
    sample_batch = min(8, w["batch_per_gpu"])
    tok_s, dt, threads = cpu_baseline.time_cpu(_dense_workload(w), sample_batch, args.steps, args.warmup)
    sample = (f"batch {sample_batch} of {w['batch_per_gpu']} (same shapes) per step, fp32 torch CPU, "
              f"{args.warmup} warm-up + {args.steps} timed fwd+bwd steps; oracle port of the reference connector "
              "(the reference's own Python needs /root/reference, which does not exist on the GPU box: "
              "oracle/ref_loader.py pins the port against it in the build container)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": tok_s, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {k: v for k, v in w.items()},
        "cpu_baseline": {"value": tok_s, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": tok_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def _dense_workload(w):
    """The CPU port runs dense streams; the ragged config is timed at its mean valid length."""
    d = dict(w)
    if "ragged_range" in d:
        lo, hi = d.pop("ragged_range")
        for k in ("audio_frames", "video_frames"):
            if d[k]:
                d[k] = (lo + min(hi, d[k])) // 2
    return d


def make_engine(pkg, w, dev, seed, **kw):
    from audio_visual_llm_b200.engine import ConnectorStep, StepShape

    plan = pkg.FusePlan(modality=w["modality"], fusion=w["fusion"], fusion_scale=w["fusion_scale"],
                        max_seq_len=w["max_seq_len"], audio_stride=w["audio_stride"], video_stride=w["video_stride"])
    shape = StepShape(batch=w["batch_per_gpu"], audio_frames=w["audio_frames"], video_frames=w["video_frames"],
                      audio_dim=w["audio_dim"], video_dim=w["video_dim"], hidden=w["hidden"],
                      prompt_len=w["prompt_len"])
    if w.get("ragged_range"):
        kw["ragged_range"] = tuple(w["ragged_range"])
    return ConnectorStep(shape, plan, dev, seed=seed, **kw), plan, shape


def timed_region(torch, dist, eng, steps, world, dev, sampler=None):
    """K steps between barrier + synchronize on both sides, CUDA events on the launching stream, MAX over ranks."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.enable_kernel_timing()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    t_start.record()
    for _ in range(steps):
        eng.step()
    t_end.record()
    barrier()
    w1 = time.time()
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = {n: sum(s.elapsed_time(e) for s, e in ev) / len(ev) for n, ev in eng.events.items() if ev}
    if "proj_bwd_dw_v" in kernel_ms:  # overlapped NCCL schedule: the dW GEMM runs as two launches
        kernel_ms["proj_bwd_dw"] += kernel_ms.pop("proj_bwd_dw_v")
    eng.events = None
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    clocks = sampler.window(w0, w1) if sampler is not None else None
    return elapsed_ms / steps, kernel_ms, clocks


def gemm_roofline(eng, kernel_ms, clocks, peaks, traffic):
    gemm_ms = 0.5 * (kernel_ms["proj_fwd"] + kernel_ms["proj_bwd_dw"])
    achieved_tf = eng.gemm_flops() / (gemm_ms * 1e-3) / 1e12
    # denominator: the sustained cuBLAS figure when the timed region ran under the power cap (back-to-back steps),
    # the burst figure otherwise -- as MEASURED_PEAKS.json defines them
    capped = bool(clocks and "sw_power_cap" in (clocks.get("reasons") or []))
    peak_tf = peaks["tf_sustained"] if (capped and peaks.get("tf_sustained")) else peaks["tf_burst"]
    return {"bound": "tensor",
            "kernel": "proj_gemm (tcgen05 cta_group::2 projector GEMM: fwd TN + dW NT launches, averaged)",
            "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
            "traffic": traffic,
            "peak_source": peaks["source"] + (", sustained figure (sw_power_cap active during the timed region)"
                                              if peak_tf != peaks["tf_burst"] else ", burst figure"),
            "frac_of_burst_peak": achieved_tf / peaks["tf_burst"],
            "frac_of_sustained_peak": achieved_tf / peaks["tf_sustained"] if peaks.get("tf_sustained") else None,
            "flops_per_launch": eng.gemm_flops(), "avg_launch_ms": gemm_ms}


# ---------------------------------------------------------------------------------------------- self check (fp64)
def _stream_columns(torch, feat, k, rep, N, valid, cols):
    """Columns `cols` of the stacked operand of one stream as fp64 [B, N, len(cols)] from the RAW tower output
    (token j reads the stack j // rep: frames k * (j // rep) .. + k - 1; frames at or past the valid length are zero)."""
    B, T, D = feat.shape
    j = torch.arange(N, device=feat.device)
    fr = (j // rep)[:, None] * k + (cols // D)[None, :]          # [N, C] frame index
    ok = fr < T
    x = feat[:, fr.clamp(max=T - 1), (cols % D)[None, :].expand(N, -1)].double()   # [B, N, C]
    x = x * ok[None].double()
    if valid is not None:
        x = x * (fr[None] < valid.view(B, 1, 1)).double()
    return x


def self_check(torch, dist, eng, world, dev, n_rows=64, n_side=32):
    """fp64 recomputation, from the raw inputs, of `n_rows` sampled output rows, n_side x n_side sampled dW entries per
    stream and the whole bias gradients of ONE step of the benchmarked engine; at N > 1 also: every rank's bucket is
    bit-identical and equals the fp64 mean over ranks on the samples."""
    g = torch.Generator(device="cpu").manual_seed(99)
    s, p = eng.shape, eng.plan
    B, N, P, H = s.batch, eng.N, s.prompt_len, s.hidden
    eng.bucket.flat.fill_(float("nan"))
    eng.emb.fill_(float("nan"))
    eng.step()
    torch.cuda.synchronize()
    if eng.bucket.peer is not None:
        eng.bucket.peer.check()
    counts = torch.tensor(eng.counts, device=dev)
    streams = []
    if eng.use_a:
        streams.append(("audio", eng.audio, p.audio_stride, p.audio_repeat, eng.audio_valid, eng.wa, eng.ba, eng.sa,
                        "audio_connector.linear"))
    if eng.use_v:
        streams.append(("video", eng.video, p.video_stride, p.video_repeat, eng.video_valid, eng.wv, eng.bv, eng.sv,
                        "video_connector.linear"))
    # ---- sampled output rows
    bs = torch.randint(0, B, (n_rows,), generator=g).to(dev)
    js = (torch.rand(n_rows, generator=g).to(dev) * counts[bs]).long().clamp(max=N - 1)
    ref = torch.zeros(n_rows, H, dtype=torch.float64, device=dev)
    for _, feat, k, rep, valid, w, b, sc, _n in streams:
        Kst = w.shape[1]
        x = _stream_columns(torch, feat, k, rep, N, valid, torch.arange(Kst, device=dev))[bs, js]   # [R, K_s]
        wp = (w * sc).to(torch.bfloat16).double()                 # what pack_weight feeds the tensor cores
        ref += x @ wp.t()
        T = feat.shape[1]
        ln = valid[bs] if valid is not None else torch.full_like(bs, T)
        present = ((js // rep) * k < ln).double()
        ref += present[:, None] * (sc * b.double())[None, :]
    got = eng.emb[bs, P + js].double()
    out_rel = float((got - ref).abs().max() / ref.abs().max())
    out_cos = float((got * ref).sum() / (got.norm() * ref.norm()))
    # ---- gradients: local fp64 references for sampled dW entries and the full db, then the mean over ranks
    valid_tok = (torch.arange(N, device=dev)[None, :] < counts[:, None])                      # [B, N]
    hs = torch.randperm(H, generator=g)[:n_side].to(dev)
    dY = eng.d_emb[:, P:P + N][:, :, :].double() * valid_tok[:, :, None].double()              # [B, N, H] fp64
    dYs = dY[:, :, hs].reshape(B * N, n_side)
    worst_dw, worst_db, nsamp = 0.0, 0.0, 0
    for _, feat, k, rep, valid, w, b, sc, name in streams:
        Kst = w.shape[1]
        ks = torch.randperm(Kst, generator=g)[:n_side].to(dev)
        # every tile's four corners of this stream's dW as well: (h, k) at multiples of the 512 x 256 tile -/+ 1
        x = _stream_columns(torch, feat, k, rep, N, valid, ks).reshape(B * N, n_side)
        ref_dw = sc * (dYs.t() @ x)                                                             # [n_side, n_side]
        T = feat.shape[1]
        ln = valid if valid is not None else torch.full((B,), T, device=dev)
        present = (((torch.arange(N, device=dev) // rep) * k)[None, :] < ln[:, None]).double()  # [B, N]
        ref_db = sc * (dY * present[:, :, None]).sum((0, 1))
        if world > 1:
            dist.all_reduce(ref_dw)
            dist.all_reduce(ref_db)
            ref_dw /= world
            ref_db /= world
        got_dw = eng.bucket[name + ".weight"][hs][:, ks].double()
        got_db = eng.bucket[name + ".bias"].double()
        rms = float(ref_dw.pow(2).mean().sqrt()) + 1e-30
        worst_dw = max(worst_dw, float(((got_dw - ref_dw).abs() / (ref_dw.abs() + rms)).max()))
        rms_b = float(ref_db.pow(2).mean().sqrt()) + 1e-30
        worst_db = max(worst_db, float(((got_db - ref_db).abs() / (ref_db.abs() + rms_b)).max()))
        nsamp += n_side * n_side
    res = {"output_rows": n_rows, "output_max_rel": out_rel, "output_cosine": out_cos, "dw_entries": nsamp,
           "dw_max_rel": worst_dw, "db_max_rel": worst_db,
           "tolerances": {"output_max_rel": 1e-2, "output_cosine": 0.9999, "dw_max_rel": 1e-3, "db_max_rel": 1e-3},
           "reference": "fp64 recomputation from the raw inputs (bf16 operands as the tensor cores see them)"}
    ok = out_rel <= 1e-2 and out_cos >= 0.9999 and worst_dw <= 1e-3 and worst_db <= 1e-3
    if world > 1:
        flat = eng.bucket.flat
        sig = torch.stack([flat.view(torch.int32).long().sum(), flat[:1 << 20].view(torch.int32).long().sum()])
        sigs = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        same = all(bool(torch.equal(sigs[0], x)) for x in sigs)
        res["ranks_bit_identical"] = same
        ok = ok and same
    res["ok"] = bool(ok and math.isfinite(out_rel) and math.isfinite(worst_dw) and math.isfinite(worst_db))
    return res


# ---------------------------------------------------------------------------------------------- eager torch / cuBLAS
def trainer_leg(torch, dist, eng, steps, world):
    """The reference's whole trainer step for the connector parameters (clip_whisper_trainer.py:453-464): forward,
    backward, gradient all-reduce, global-norm clip, AdamW -- `ConnectorStep.train_step()`.  The AdamW kernel writes the
    bf16 pack of the updated weights, so the forward's pack launches disappear.  All ranks run it; max over ranks."""
    params = eng.attach_optimizer(lr=1e-6).params   # a tiny rate: the weights stay where the other legs expect them
    for _ in range(3):
        eng.train_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.train_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if eng.bucket.peer is not None:
        eng.bucket.peer.check()
    eng.detach_optimizer()
    npack = int(eng.use_a) + int(eng.use_v)
    return {"ms_per_step": ms, "value": eng.fused_tokens * world / (ms * 1e-3), "unit": UNIT, "steps": steps,
            "launches_per_step": eng.launches_per_step - npack + 2 + len(params),
            "what": "fwd + bwd (+ fused gradient all-reduce) + global-norm clip + AdamW on the flat bucket "
                    "(ConnectorStep.train_step); the optimizer kernel emits the bf16 weight pack of the next forward"}


def graphed_leg(torch, eng, steps):
    """The same step replayed as one CUDA graph (engine.ConnectorStep.capture_graph): what is left when the host-side
    launch cost is taken out.  Single process only."""
    graph = eng.capture_graph()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        graph.replay()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    return {"ms_per_step": ms, "value": eng.fused_tokens / (ms * 1e-3), "unit": UNIT, "steps": steps,
            "what": "the whole step (all launches, side-stream work included) captured once and replayed as one CUDA graph"}


def gpu_eager_leg(torch, eng, iters=20):
    """The reference connector's semantics as eager PyTorch (bf16 autocast, cuBLAS GEMMs) on the same GPU, same shapes,
    same inputs: per-modality nn.Linear, pad, weighted sum, prompt-embedding cat, ones mask, label rule, autograd
    backward (dW, db into fp32 master grads).  Two formulations are timed -- `two_linear` is the reference's structure
    (modality_connector.py:43-44, clip_whisper_model.py:424-462), `cat_linear` concatenates [a ; v] first and runs ONE
    linear -- plus the bare cuBLAS GEMMs at the two projector shapes."""
    import torch.nn.functional as F

    if eng.ragged:
        return None
    s, p = eng.shape, eng.plan
    B, N, P, H = s.batch, eng.N, s.prompt_len, s.hidden
    dev = eng.device
    prm = {}
    if eng.use_a:
        prm["wa"], prm["ba"] = eng.wa.clone().requires_grad_(True), eng.ba.clone().requires_grad_(True)
    if eng.use_v:
        prm["wv"], prm["bv"] = eng.wv.clone().requires_grad_(True), eng.bv.clone().requires_grad_(True)
    prompt_ids = eng.input_ids[:, :P]

    def stack(x, k, rep):
        Bx, T, D = x.shape
        n = -(-T // k)
        if n * k != T:
            x = F.pad(x, (0, 0, 0, n * k - T))
        x = x.reshape(Bx, n, k * D)
        if rep > 1:
            x = x.repeat_interleave(rep, 1)
        return x

    def fit(y):  # _pad_or_truncate (clip_whisper_model.py:320-374)
        if y.shape[1] > N:
            return y[:, :N]
        if y.shape[1] < N:
            return torch.cat([y, torch.zeros(B, N - y.shape[1], y.shape[2], dtype=y.dtype, device=dev)], 1)
        return y

    def tail(av):
        emb = torch.cat([F.embedding(prompt_ids, eng.embed_table), av], 1)
        mask = torch.ones(B, P + N, dtype=torch.int64, device=dev)
        lab = eng.labels_in.clone()
        lab[lab == 0] = -100
        S = P + N
        if lab.shape[1] > S:
            lab = lab[:, :S]
        elif lab.shape[1] < S:
            lab = torch.cat([lab, torch.full((B, S - lab.shape[1]), -100, dtype=lab.dtype, device=dev)], 1)
        return emb, mask, lab

    def two_linear():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ys = []
            if eng.use_a:
                ys.append(eng.sa * fit(F.linear(stack(eng.audio, p.audio_stride, p.audio_repeat), prm["wa"], prm["ba"])))
            if eng.use_v:
                ys.append(eng.sv * fit(F.linear(stack(eng.video, p.video_stride, p.video_repeat), prm["wv"], prm["bv"])))
            av = ys[0] if len(ys) == 1 else ys[0] + ys[1]
            emb, mask, lab = tail(av.to(torch.bfloat16))
        emb.backward(eng.d_emb)

    def cat_linear():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            xs, ws, bs = [], [], []
            if eng.use_a:
                xs.append(fit(stack(eng.audio, p.audio_stride, p.audio_repeat)))
                ws.append(eng.sa * prm["wa"])
                bs.append(eng.sa * prm["ba"])
            if eng.use_v:
                xs.append(fit(stack(eng.video, p.video_stride, p.video_repeat)))
                ws.append(eng.sv * prm["wv"])
                bs.append(eng.sv * prm["bv"])
            x = xs[0] if len(xs) == 1 else torch.cat(xs, -1)
            w = ws[0] if len(ws) == 1 else torch.cat(ws, 1)
            b = bs[0] if len(bs) == 1 else bs[0] + bs[1]
            emb, mask, lab = tail(F.linear(x, w, b))
        emb.backward(eng.d_emb)

    def timeit(fn, n=iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / n

    def step(fn):
        def run():
            for q in prm.values():
                q.grad = None
            fn()
        return run

    out = {"what": "reference connector semantics in eager PyTorch bf16 autocast (cuBLAS GEMMs + elementwise / cat "
                   "kernels + autograd) on the same GPU, inputs resident in HBM, same shapes as `value`",
           "iters": iters}
    # the zero-padded video stream of index-aligned `both` (cfg1 / cfg2p) makes cat_linear's bias mask wrong; it is only
    # a valid formulation when every token carries both streams
    variants = {"two_linear": two_linear}
    if not (eng.use_a and eng.use_v) or -(-s.audio_frames // p.audio_stride) == -(-s.video_frames // p.video_stride):
        variants["cat_linear"] = cat_linear
    for name, fn in variants.items():
        ms = timeit(step(fn))
        out[name] = {"ms_per_step": ms, "value": eng.fused_tokens / (ms * 1e-3)}
    best = min(variants, key=lambda n: out[n]["ms_per_step"])
    out["best"] = best
    out["ms_per_step"] = out[best]["ms_per_step"]
    out["value"] = out[best]["value"]
    out["unit"] = UNIT
    del prm
    # bare cuBLAS at the two projector GEMM shapes (bf16 operands; fp32 output for dW when torch exposes it)
    M, K = eng.M, eng.K
    A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    W = torch.randn(H, K, device=dev, dtype=torch.bfloat16)
    dYm = torch.randn(M, H, device=dev, dtype=torch.bfloat16)
    fwd_ms = timeit(lambda: torch.matmul(A, W.t()))
    try:
        torch.mm(dYm.t(), A, out_dtype=torch.float32)
        dw_ms = timeit(lambda: torch.mm(dYm.t(), A, out_dtype=torch.float32))
        dw_out = "fp32"
    except (TypeError, RuntimeError):
        dw_ms = timeit(lambda: torch.matmul(dYm.t(), A))
        dw_out = "bf16"
    fl = 2.0 * M * K * H
    out["cublas"] = {"fwd_ms": fwd_ms, "fwd_TFLOPs": fl / fwd_ms / 1e9, "dw_ms": dw_ms, "dw_TFLOPs": fl / dw_ms / 1e9,
                     "dw_out": dw_out, "shape_MKH": [M, K, H],
                     "note": "torch.matmul back to back, no epilogue work (no bias, no scatter, no db, no fp32 scale)"}
    del A, W, dYm
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sustained region, the eager-torch leg, the unfused step and the strong-scaling block")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling (BASELINE configs[4]: batch 256 over 2/4/8 GPUs): split this many samples over "
                         "the ranks instead of 32 per GPU; the default (0) is the weak-scaling contract")
    ap.add_argument("--overlap", action="store_true",
                    help="N > 1, NCCL schedule only: all-reduce the audio-weight span under the video-weight dW launch")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry

    entry.build()
    import audio_visual_llm_b200 as pkg

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG is left as the caller set it (the driver reads NCCL's own log); whatever NCCL / c10d print on
        # stdout while the communicator is created goes to stderr, so that rank 0's stdout carries ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    peaks = load_peaks()
    w = dict(CONFIGS[args.config])
    scaling = "weak"
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} does not divide over {world} ranks")
        per = w["batch_per_gpu"]
        w["batch_per_gpu"] = args.global_batch // world
        w["workload"] = w["workload"].replace(f"batch {per}/GPU", f"global batch {args.global_batch} ({w['batch_per_gpu']}/GPU)")
        scaling = "strong"
    eng, plan, shape = make_engine(pkg, w, dev, 1234 + rank)
    eng.overlap_comm = bool(args.overlap) or eng.overlap_comm
    multicast = eng.bucket.peer is not None and eng.bucket.peer.mc is not None  # (the bucket is closed further down)

    # ------------------------------------------------------------------ device-resident throughput (`value`)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        eng.step()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.wait_first_sample()
    ms_per_step, kernel_ms, clocks = timed_region(torch, dist, eng, args.steps, world, dev,
                                                  sampler if rank == 0 else None)
    per_rank = None
    if world > 1:
        # every rank's GEMM times: ranks run at different power-capped clocks, and the fused dW + all-reduce launch
        # (like any all-reduce) ends with the slowest rank
        mine = torch.tensor([kernel_ms.get("proj_fwd", 0.0), kernel_ms.get("proj_bwd_dw", 0.0)], device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"proj_fwd_ms": [round(float(x[0]), 4) for x in allr],
                    "proj_bwd_dw_ms": [round(float(x[1]), 4) for x in allr]}
    value = eng.fused_tokens * world / (ms_per_step * 1e-3)
    if eng.ragged and world > 1:  # ragged: ranks hold different token counts
        t = torch.tensor([float(eng.fused_tokens)], device=dev)
        dist.all_reduce(t)
        value = float(t.item()) / (ms_per_step * 1e-3)
    assert int(eng.status.item()) == 0, "placeholder / token count mismatch"
    if eng.bucket.peer is not None:
        eng.bucket.peer.check()  # a fused all-reduce launch that gave up on a peer invalidates the run

    # ------------------------------------------------------------------ the same loop over >= 1 s (power-capped clocks)
    sustained = None
    if not args.no_extras:
        n_sus = int(min(5000, max(50, math.ceil(1100.0 / ms_per_step))))
        ms_sus, k_sus, clk_sus = timed_region(torch, dist, eng, n_sus, world, dev, sampler if rank == 0 else None)
        sustained = {"steps": n_sus, "ms_per_step": ms_sus, "value": value * ms_per_step / ms_sus, "unit": UNIT,
                     "clocks": clk_sus, "kernel_ms": k_sus}
    if rank == 0:
        time.sleep(0.05)
        sampler.stop()

    # ------------------------------------------------------------------ fp64 self check of the benchmarked engine
    check = self_check(torch, dist, eng, world, dev)

    # the same step with the stand-alone gather and splice-bwd kernels (the general path: ragged lengths, CLS
    # views, explicit placeholder layouts); gives the per-kernel HBM numbers of the kernels the fused step skips
    unfused = None
    if rank == 0 and eng.direct and not args.no_extras:
        eng2, _, _ = make_engine(pkg, w, dev, 1234 + rank, fuse_gather=False, fused_allreduce=False)
        for _ in range(args.warmup):
            eng2.step(allreduce=False)
        torch.cuda.synchronize()
        eng2.enable_kernel_timing()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n2 = max(3, min(args.steps, 50))
        u0.record()
        for _ in range(n2):
            eng2.step(allreduce=False)
        u1.record()
        torch.cuda.synchronize()
        k2 = {n: sum(s.elapsed_time(e) for s, e in ev) / len(ev) for n, ev in eng2.events.items() if ev}
        unfused = {"ms_per_step": u0.elapsed_time(u1) / n2, "steps": n2, "kernel_ms": k2,
                   "launches_per_step": eng2.launches_per_step}
        for n in ("gather", "splice_bwd"):
            kernel_ms.setdefault(n, k2[n])
        kernel_ms["splice_fwd_unfused"] = k2["splice_fwd"]
        del eng2
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()

    # ------------------------------------------------------------------ the trainer step (clip + AdamW) beside it (needs the engine's peer
    # bucket, which the strong-scaling block and the e2e leg close: one peer / multicast bucket at a time)
    trainer = None
    if not args.no_extras:
        trainer = trainer_leg(torch, dist, eng, max(20, min(args.steps, 200)), world)

    # ------------------------------------------------------------------ strong scaling (BASELINE configs[4]) beside it
    strong = None
    if world > 1 and not args.global_batch and args.config == "cfg2" and not args.no_extras and 256 % world == 0:
        strong = strong_scaling_block(torch, dist, pkg, w, dev, rank, world, eng, ms_per_step, args)

    # ------------------------------------------------------------------ end to end from pinned host tensors
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, dist, pkg, eng, plan, dev, args, world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel
    traffic = None
    tpath = ROOT / "profiles" / "gemm_traffic.json"
    if tpath.exists() and args.config == "cfg2":
        traffic = json.loads(tpath.read_text()).get("dram_bytes_per_launch")
    roofline = gemm_roofline(eng, kernel_ms, clocks, peaks, traffic)
    if sustained is not None:
        r2 = gemm_roofline(eng, sustained["kernel_ms"], sustained["clocks"], peaks, traffic)
        sustained["roofline"] = {k: r2[k] for k in ("achieved", "peak", "frac", "peak_source", "avg_launch_ms")}

    def hbm(name, nbytes):
        ms = kernel_ms.get(name)
        return None if ms is None else {"ms": ms, "GBps": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / peaks["hbm"],
                                        "frac_hbm_nominal_8TBps": nbytes / ms / 1e6 / 8000.0}

    def tens(name):
        ms = kernel_ms[name]
        return {"ms": ms, "TFLOPs": eng.gemm_flops() / ms / 1e9, "frac_bf16_burst": eng.gemm_flops() / ms / 1e9 / peaks["tf_burst"]}

    standalone = " (stand-alone, unfused step)" if eng.direct else ""
    kernels = {
        "gather" + standalone: hbm("gather", eng.gather_bytes()),
        "proj_fwd": tens("proj_fwd"),
        ("splice_fwd (text rows + masks; AV rows are written by the GEMM epilogue)" if eng.direct else "splice_fwd"):
            hbm("splice_fwd", (4 * shape.batch * shape.prompt_len * shape.hidden + 16 * shape.batch * eng.S)
                if eng.direct else eng.splice_bytes()),
        "splice_fwd (stand-alone, unfused step)": hbm("splice_fwd_unfused", eng.splice_bytes()),
        "splice_bwd" + standalone: hbm("splice_bwd", 4 * eng.M * shape.hidden),
        "proj_bwd_dw (+ db work items)" if eng.bias_in_gemm else "proj_bwd_dw": tens("proj_bwd_dw"),
        "colsum": hbm("colsum", 2 * eng.M * shape.hidden),
    }
    kernels = {k: v for k, v in kernels.items() if v is not None}

    graphed = None
    if world == 1 and not args.no_extras:
        graphed = graphed_leg(torch, eng, max(20, min(args.steps, 200)))
    gpu_eager = None
    if not args.no_extras:
        # as many iterations as the timed region had steps (bounded), so that the eager / cuBLAS numbers are taken in the
        # same clock regime (burst for the driver's 20 steps, power-capped for long runs) as the ones they sit beside
        gpu_eager = gpu_eager_leg(torch, eng, iters=max(20, min(args.steps, 300)))
        if gpu_eager is not None:
            gpu_eager["ours_over_eager_step"] = gpu_eager["ms_per_step"] / ms_per_step
            gpu_eager["ours_over_cublas"] = {"fwd": gpu_eager["cublas"]["fwd_ms"] / kernel_ms["proj_fwd"],
                                             "dw": gpu_eager["cublas"]["dw_ms"] / kernel_ms["proj_bwd_dw"]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline

        sb, ss, sw = min(8, w["batch_per_gpu"]), 3, 1
        tok_s, dt, threads = cpu_baseline.time_cpu(_dense_workload(w), sb, ss, sw)
        cpu = {"value": tok_s, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"batch {sb} of {w['batch_per_gpu']} (same shapes), {sw} warm-up + {ss} timed fwd+bwd steps, "
                         f"fp32 torch CPU oracle port ({dt:.2f} s/step); the reference's own Python cannot travel to the "
                         "GPU box (no /root/reference there): oracle/ref_loader.py pins the port against it in the build "
                         "container (tests/test_oracle_golden.py)"}
        if args.config == "cfg2" and not args.no_extras:
            # cfg2': the knobs the reference code itself implements (k = 1, index-aligned, sum fusion), BASELINE.md section 3
            tp, dtp, _ = cpu_baseline.time_cpu(CONFIGS["cfg2p"], 4, 2, 1)
            cpu["cfg2p_parity_variant"] = {"value": tp, "unit": UNIT, "sample": f"batch 4 of 32, 1 warm-up + 2 timed steps ({dtp:.2f} s/step)"}

    if world == 1:
        collective = "none"
    elif eng.fused_allreduce:
        mc = multicast
        collective = ("projector-grad all-reduce fused into the dW GEMM launch (100.7 MB fp32 flat bucket; comm warps of "
                      "the GEMM CTAs reduce finished tiles over NVLink while later tiles are computed; transport: " +
                      ("NVSwitch multicast mapping, multimem.ld_reduce + multimem.st)" if mc
                       else "peer-mapped memory (CUDA IPC), peer loads + peer stores)"))
    else:
        collective = ("projector-grad all-reduce (NCCL sum of pre-scaled grads, 100.7 MB fp32 flat bucket; " +
                      ("audio-weight span overlapped with the video-weight dW launch)" if eng.overlap_comm
                       else "one call after the backward)"))
    if eng.direct:
        step_desc = ("fused: tower outputs -> 2-segment GEMM whose epilogue writes the AV rows of inputs_embeds -> text rows + "
                     "masks; the dW GEMM reads d(inputs_embeds) in place and produces db from extra work items of its tile schedule")
    else:
        step_desc = "gather -> GEMM -> splice (+ masks); splice-bwd -> dW GEMM (+ db work items)"
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {**{k: v for k, v in w.items()},
                   "global_batch": w["batch_per_gpu"] * world, "fused_tokens_per_step": int(round(value * ms_per_step * 1e-3)),
                   "parallelism": f"dp{world}", "step": step_desc, "collective": collective,
                   "l2": "no flush: one step streams several times the 126 MB L2 (features, W, embeds, grads)"},
        "roofline": roofline, "kernels": kernels, "per_rank_kernel_ms": per_rank, "sustained": sustained,
        "self_check": check, "gpu_eager": gpu_eager, "graphed": graphed, "strong_scaling": strong,
        "trainer_step": trainer,
        "unfused_step": unfused,
        "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
        "gpu_launches": eng.launches_per_step * args.steps,
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def strong_scaling_block(torch, dist, pkg, w, dev, rank, world, eng, ms_weak, args):
    """BASELINE configs[4]: global batch 256 split over the ranks (128 / 64 / 32 per GPU), same step, same collective."""
    per = 256 // world
    if per == w["batch_per_gpu"]:  # 8 ranks: the weak-scaling run IS the batch-256 run
        return {"global_batch": 256, "batch_per_gpu": per, "ms_per_step": ms_weak,
                "value": eng.fused_tokens * world / (ms_weak * 1e-3), "unit": UNIT, "same_as": "value (32 samples / GPU)"}
    w2 = dict(w, batch_per_gpu=per)
    if eng.bucket.peer is not None:
        eng.bucket.peer.close()   # one multicast / peer bucket at a time
    torch.cuda.empty_cache()
    eng2, _, _ = make_engine(pkg, w2, dev, 4321 + rank)
    for _ in range(3):
        eng2.step()
    torch.cuda.synchronize()
    n = max(5, min(args.steps, 30))
    ms, _, _ = timed_region(torch, dist, eng2, n, world, dev)
    if eng2.bucket.peer is not None:
        eng2.bucket.peer.check()
        eng2.bucket.peer.close()
    tokens = eng2.fused_tokens * world
    del eng2
    torch.cuda.empty_cache()
    return {"global_batch": 256, "batch_per_gpu": per, "steps": n, "ms_per_step": ms, "value": tokens / (ms * 1e-3),
            "unit": UNIT, "scaling": "strong"}


def run_e2e(torch, dist, pkg, eng, plan, dev, args, world):
    """Public-API step from pinned host inputs: H2D(features, ids, labels) -> fused_connector fwd -> backward
    [-> gradient all-reduce inside the dW launch, parallel.FusedGradSync] -> D2H(masks, labels, bias grads).
    Same metric, max over ranks."""
    from audio_visual_llm_b200.engine import HostFeeder
    from audio_visual_llm_b200.parallel import FusedGradSync

    from audio_visual_llm_b200 import numa

    s = eng.shape
    # input buffers on the GPU's own NUMA node when the host exposes one (pool boxes are single-node KVM guests:
    # "unbound:no NUMA choice", profiles/r02_h2d_numa_probe_n4.log)
    node = numa.gpu_numa_node(dev.index if dev.index is not None else torch.cuda.current_device())
    placement = []

    def pin(t):
        buf, how = numa.pin_like(t.cpu(), node)
        placement.append(how)
        return buf

    audio_h = pin(eng.audio) if eng.use_a else None
    video_h = pin(eng.video) if eng.use_v else None
    ids_h = pin(eng.input_ids)
    labels_h = pin(eng.labels_in)
    wa, ba, wv, bv = ((t.clone().requires_grad_(True) if t is not None else None)
                      for t in (eng.wa, eng.ba, eng.wv, eng.bv))
    mask_h = torch.empty(s.batch, eng.S, dtype=torch.int64).pin_memory()
    lab_h = torch.empty(s.batch, eng.S, dtype=torch.int64).pin_memory()
    nb = int(eng.use_a) + int(eng.use_v)
    db_h = torch.empty(nb, s.hidden, dtype=torch.float32).pin_memory()
    h2d = sum(t.nbytes for t in (audio_h, video_h, ids_h, labels_h) if t is not None)
    d2h = mask_h.nbytes + lab_h.nbytes + db_h.nbytes
    sync = None
    if world > 1:
        if eng.bucket.peer is not None:
            eng.bucket.peer.close()  # the engine's bucket is done; one multicast / peer bucket at a time
        sync = FusedGradSync(wa, ba, wv, bv)
    feeder = HostFeeder(dev)
    batch = (audio_h, video_h, ids_h, labels_h)
    lens = {}
    if eng.ragged:
        la, lv = eng.lengths_host
        lens = dict(audio_lengths=la, video_lengths=lv, total_tokens=eng.M)

    def step(last):
        if sync is None:
            for p in (wa, ba, wv, bv):
                if p is not None:
                    p.grad = None
        a, v, ids, lab_in = feeder.take()           # this step's inputs (H2D issued one step earlier, inside the region)
        if not last:
            feeder.prefetch(batch)                  # next step's H2D overlaps this step's kernels
        if eng.ragged:
            emb, mask, lab = pkg.fused_connector(a, v, wa, ba, wv, bv, plan, input_ids=ids, embed_table=eng.embed_table,
                                                 labels=lab_in, placeholder_id=eng.placeholder_id,
                                                 out_dtype=torch.bfloat16, grad_sync=sync, **lens)
        else:
            emb, mask, lab = pkg.fused_connector(a, v, wa, ba, wv, bv, plan, prompt_ids=ids[:, :s.prompt_len],
                                                 embed_table=eng.embed_table, labels=lab_in,
                                                 placeholder_id=eng.placeholder_id, out_dtype=torch.bfloat16,
                                                 grad_sync=sync)
        emb.backward(eng.d_emb)
        feeder.release()
        if sync is not None and not sync.fused:     # no peer mapping on this box: one NCCL all-reduce of the bucket
            sync.bucket.allreduce()
        mask_h.copy_(mask, non_blocking=True)
        lab_h.copy_(lab, non_blocking=True)
        for i, b in enumerate(t for t in (ba, bv) if t is not None):
            db_h[i].copy_(b.grad, non_blocking=True)

    def run(n):
        feeder.prefetch(batch)                      # first batch: its copy is inside the timed region too
        for i in range(n):
            step(i == n - 1)

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        fn()
        t1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    run(max(3, args.warmup))
    ms_step = timed(lambda: run(args.steps)) / args.steps
    if sync is not None:
        sync.check()
    tokens = eng.fused_tokens * world
    if eng.ragged and world > 1:
        t = torch.tensor([float(eng.fused_tokens)], device=dev)
        dist.all_reduce(t)
        tokens = float(t.item())
    # the ceiling of this box's host side: the same H2D copies alone, every rank at once, no kernels
    n_probe = max(3, min(args.steps, 10))
    dsts = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in batch if t is not None]
    srcs = [t for t in batch if t is not None]

    def copies():
        for _ in range(n_probe):
            for d_, s_ in zip(dsts, srcs):
                d_.copy_(s_, non_blocking=True)

    copies()
    ms_copy = timed(copies) / n_probe
    if sync is not None:
        sync.close()
    for t in (audio_h, video_h, ids_h, labels_h):
        if t is not None:
            numa.release(t)
    collective = "none" if world == 1 else (
        "fused into the dW GEMM launch through parallel.FusedGradSync (the parameters' .grad are views of the peer-mapped "
        "bucket)" if sync.fused else "NCCL all-reduce of the flat bucket")
    return {"value": tokens / (ms_step * 1e-3), "unit": UNIT, "ms_per_step": ms_step,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "collective": collective,
            "h2d_only_ms_per_step": ms_copy, "h2d_only_GBps_per_gpu": h2d / ms_copy / 1e6,
            "h2d_GBps_per_gpu_in_e2e": h2d / ms_step / 1e6,
            "host_numa": {"gpu_node": node, "nodes": numa.nodes(), "pinned_placement": sorted(set(placement))},
            "api": "HostFeeder (pinned host -> device, double-buffered on a copy stream) -> fused_connector(...) -> "
                   "emb.backward(dLLM) -> D2H of masks / labels / bias grads; every step's H2D is inside the timed region; "
                   "h2d_only_* = the same copies alone on every rank at once (the host-side ceiling of this box)"}


if __name__ == "__main__":
    main()
