"""Static-buffer training-step engine for the connector: one `step()` = the whole hot path once,
fwd + bwd, with every buffer pre-allocated (no allocator traffic, CUDA-graph friendly).

    pack weights -> gather -> projector GEMM -> splice (+masks)          forward
    splice-bwd -> dW GEMM (+ db work items) [-> grad all-reduce]         backward

This is the same kernel sequence `connector_ops._FusedConnectorFn` runs under autograd; the engine exists so
the data-parallel trainer (and bench.py) can drive it without per-step Python allocation.  The projector
gradients land in ONE flat fp32 bucket ([dWa | dWv | dba | dbv]); the only collective of the path -- the
all-reduce the trainer needs between backward() and clip_grad_norm_ (clip_whisper_trainer.py:454-458) -- runs inside
the dW GEMM launch over peer-mapped / multicast memory (`_backward_fused_allreduce`), or as one NCCL call after it.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib as L
from .connector_ops import FusePlan
from .parallel import GradBucket


@dataclass(frozen=True)
class StepShape:
    batch: int
    audio_frames: int
    video_frames: int
    audio_dim: int
    video_dim: int
    hidden: int
    prompt_len: int = 16
    vocab: int = 32000
    label_len: int = 256


class ConnectorStep:
    def __init__(self, shape: StepShape, plan: FusePlan, device, out_dtype=torch.bfloat16, seed: int = 0,
                 process_group=None, fuse_gather: bool = True, fused_allreduce: Optional[bool] = None,
                 ragged_range: Optional[tuple] = None, ragged_seed: int = 4):
        """ragged_range=(lo, hi): variable-length placeholders (BASELINE configs[3]) -- every sample gets a valid frame
        count drawn from U{lo..hi} (seed `ragged_seed`) for each stream, contributes tokens(len) fused tokens, and its
        id row is `[prompt | that many placeholders | pad]`; the step then runs gather -> GEMM -> ragged splice."""
        if out_dtype != torch.bfloat16:
            raise L.ConnectorError("the step engine runs the bf16 training configuration")
        L.require_device(torch.device(device).index or 0)
        self.shape, self.plan, self.device = shape, plan, torch.device(device)
        s, p, dev = shape, plan, self.device
        self.use_a = p.modality in ("audio", "both")
        self.use_v = p.modality in ("video", "both")
        self.sa, self.sv = p.scales(self.use_a, self.use_v)
        self.N = p.tokens(s.audio_frames if self.use_a else None, s.video_frames if self.use_v else None)
        self.M = s.batch * self.N
        self.ragged = ragged_range is not None
        self.audio_valid = self.video_valid = self.tok_offset = None
        self.counts = [self.N] * s.batch
        if self.ragged:
            from .connector_ops import ragged_token_offsets

            gr = torch.Generator(device="cpu").manual_seed(ragged_seed)
            lo, hi = ragged_range

            def draw(T):
                return torch.randint(lo, min(hi, T) + 1, (s.batch,), generator=gr).tolist()

            la = draw(s.audio_frames) if self.use_a else None
            lv = draw(s.video_frames) if self.use_v else None
            self.lengths_host = (la, lv)
            self.tok_offset, self.audio_valid, self.video_valid, self.M = ragged_token_offsets(
                p, s.batch, s.audio_frames if self.use_a else None, s.video_frames if self.use_v else None, la, lv, dev)
            offs = self.tok_offset.tolist()
            self.counts = [offs[i + 1] - offs[i] for i in range(s.batch)]
        self.Ka = p.audio_stride * s.audio_dim if self.use_a else 0
        self.Kv = p.video_stride * s.video_dim if self.use_v else 0
        self.K = self.Ka + self.Kv
        self.S = s.prompt_len + self.N
        self.placeholder_id = s.vocab  # one past the text vocabulary
        H = s.hidden
        g = torch.Generator(device="cpu").manual_seed(seed)
        bf = torch.bfloat16

        def randn(*shape_, scale=1.0):
            return (torch.randn(*shape_, generator=g) * scale)

        # ---- parameters (fp32 masters, reference init scale) and the flat gradient bucket
        self.wa = randn(H, self.Ka, scale=(6.0 / (H + self.Ka)) ** 0.5).to(dev) if self.use_a else None
        self.wv = randn(H, self.Kv, scale=(6.0 / (H + self.Kv)) ** 0.5).to(dev) if self.use_v else None
        self.ba = randn(H, scale=0.02).to(dev) if self.use_a else None
        self.bv = randn(H, scale=0.02).to(dev) if self.use_v else None
        sizes: Dict[str, tuple] = {}
        if self.use_a:
            sizes["audio_connector.linear.weight"] = (H, self.Ka)
        if self.use_v:
            sizes["video_connector.linear.weight"] = (H, self.Kv)
        if self.use_a:
            sizes["audio_connector.linear.bias"] = (H,)
        if self.use_v:
            sizes["video_connector.linear.bias"] = (H,)
        # Data parallel (N > 1): the gradient all-reduce runs INSIDE the dW GEMM launch over peer-mapped memory
        # (`avc_proj_bwd_dw_allreduce`); AVC_FUSED_ALLREDUCE=0 falls back to one NCCL all-reduce after the backward.
        import torch.distributed as dist
        ddp = dist.is_available() and dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        if fused_allreduce is None:
            fused_allreduce = ddp and dist.get_world_size(process_group) > 1 and \
                os.environ.get("AVC_FUSED_ALLREDUCE", "1") != "0"
        if fused_allreduce and ddp and dist.get_world_size(process_group) > L.COMM_MAX_WORLD:
            fused_allreduce = False  # the flag layout holds COMM_MAX_WORLD ranks; larger jobs all-reduce through NCCL
        self.fused_allreduce = bool(fused_allreduce)
        # Transport of the fused all-reduce: an NVSwitch multicast mapping of the buckets (multimem.ld_reduce adds in the
        # switch, multimem.st writes every rank) instead of peer loads / stores: (1 + 1 / world) x the bucket per GPU and
        # direction instead of 2 (world - 1) / world x.  AVC_COMM_MULTIMEM=0 selects the peer transport.
        # Falls back to the peer mapping, on all ranks alike, when multicast objects are not available.
        multimem = self.fused_allreduce and os.environ.get("AVC_COMM_MULTIMEM", "1") == "1"
        try:
            self.bucket = GradBucket(sizes, dev, process_group=process_group, peer=self.fused_allreduce,
                                     multimem=multimem)
        except L.ConnectorError as e:
            if not (self.fused_allreduce and ddp):
                raise
            # no peer mapping between the GPUs (PeerMemory agreed on the failure across ranks): NCCL all-reduce
            print(f"[avc] fused gradient all-reduce unavailable ({e}); using NCCL", file=sys.stderr)
            self.fused_allreduce = False
            self.bucket = GradBucket(sizes, dev, process_group=process_group)
        # ---- inputs resident in HBM
        self.audio = randn(s.batch, s.audio_frames, s.audio_dim).to(bf).to(dev) if self.use_a else None
        self.video = randn(s.batch, s.video_frames, s.video_dim).to(bf).to(dev) if self.use_v else None
        prompt = torch.randint(1, s.vocab, (s.batch, s.prompt_len), generator=g)
        ph = torch.full((s.batch, self.N), self.placeholder_id, dtype=torch.int64)
        if self.ragged:  # right-padded: the placeholders of sample b end after counts[b] positions, pad id 0 follows
            for b_, c_ in enumerate(self.counts):
                ph[b_, c_:] = 0
        self.input_ids = torch.cat([prompt, ph], 1).to(dev)
        labels = torch.randint(1, s.vocab, (s.batch, s.label_len), generator=g)
        labels[:, s.label_len * 3 // 4:] = 0  # pad tail (pad id 0)
        self.labels_in = labels.to(dev)
        self.embed_table = randn(s.vocab + 1, H, scale=0.02).to(bf).to(dev)
        self.d_emb = randn(s.batch, self.S, H).to(bf).to(dev)  # upstream gradient from the LLM
        # ---- workspaces
        self.A = torch.empty(self.M, self.K, dtype=bf, device=dev)
        self.flags = torch.empty(self.M, dtype=torch.uint8, device=dev)
        self.wp = torch.empty(H, self.K, dtype=bf, device=dev)
        self.Y = torch.empty(self.M, H, dtype=bf, device=dev)
        self.emb = torch.empty(s.batch, self.S, H, dtype=bf, device=dev)
        self.mask = torch.empty(s.batch, self.S, dtype=torch.int64, device=dev)
        self.labels_out = torch.empty(s.batch, self.S, dtype=torch.int64, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.dY = torch.empty(self.M, H, dtype=bf, device=dev)
        self.colsum_ws = L.colsum_workspace(H, dev)
        # db comes out of the dW GEMM launch (64-wide work items that contract the dY panels with the token-present
        # operand); AVC_BIAS_IN_GEMM=0 runs the stand-alone column-sum kernel next to the GEMM instead
        self.bias_in_gemm = os.environ.get("AVC_BIAS_IN_GEMM", "1") != "0"
        self._present = None
        self.sp = L.make_splice(self.input_ids, self.placeholder_id, 0, H, tokens_per_sample=self.N,
                                tok_offset=self.tok_offset,
                                embed_table=self.embed_table, attention_mask=self.mask, mask_mode=p.mask_mode,
                                label_mode=p.label_mode, labels_in=self.labels_in, labels_out=self.labels_out,
                                status=self.status)
        # same descriptor for the fused step: the GEMM epilogue has already written the AV rows of inputs_embeds
        self.sp_text = L.make_splice(self.input_ids, self.placeholder_id, 0, H, tokens_per_sample=self.N,
                                     embed_table=self.embed_table, attention_mask=self.mask, mask_mode=p.mask_mode,
                                     label_mode=p.label_mode, labels_in=self.labels_in, labels_out=self.labels_out,
                                     status=self.status, av_rows_in_place=True)
        self.emb_av = self.emb[:, s.prompt_len:, :]  # [B, N, H] view: where the projected rows live
        # Gather-free ("direct") mode: when every stream is dense and its frame count divides by the stride, the
        # stacked operand is a free reshape of the tower output, so the GEMMs read it in place (two K segments) and
        # the backward reads d(inputs_embeds) in place (the `[prompt | AV]` layout puts the AV rows of sample b at
        # rows P .. P+N-1).  Otherwise: gather -> GEMM, splice-bwd -> GEMM.
        def free(frames, k):
            return frames % k == 0 and frames // k == self.N
        self.direct = bool(fuse_gather and not self.ragged and p.audio_repeat == 1 and p.video_repeat == 1
                           and (not self.use_a or free(s.audio_frames, p.audio_stride))
                           and (not self.use_v or free(s.video_frames, p.video_stride)))
        npack = int(self.use_a) + int(self.use_v)
        # packs + {[gather] gemm splice [splice_bwd] gemm} [+ colsum]
        self.launches_per_step = npack + (3 if self.direct else 5) + (0 if self.bias_in_gemm else 1)
        self.events = None  # optional per-kernel CUDA events, see enable_kernel_timing()
        self.optimizer = None  # set by attach_optimizer(): train_step() then runs clip + AdamW after the backward
        self.nvtx = os.environ.get("AVC_NVTX", "0") == "1"
        # N > 1: optionally all-reduce the audio-weight span while the video-weight dW launch still runs.  Measured on
        # B200 x8 (profiles/README.md): NCCL needs ~48+ SMs to run at speed, which the persistent GEMM must give up, so
        # the overlapped schedule is no faster than one all-reduce after the backward (1.32 ms either way at N = 8,
        # 1.23 vs 1.24 ms at N = 2).  Off by default.
        self.overlap_comm = os.environ.get("AVC_OVERLAP_COMM", "0") == "1"
        # SMs the second dW launch leaves to the concurrent NCCL kernel
        self.comm_reserve_sms = int(os.environ.get("AVC_COMM_RESERVE_SMS", "48"))
        self._comm_stream = None
        self._side_stream = None
        self.prescale_grads = os.environ.get("AVC_PRESCALE_GRADS", "1") != "0"
        # run the small HBM-bound kernels of the fused step (text rows + masks, bias sums) on a side stream,
        # concurrently with the GEMMs
        self.side_streams = os.environ.get("AVC_SIDE_STREAMS", "1") != "0"
        self._num_sms = torch.cuda.get_device_properties(self.device).multi_processor_count

    # ------------------------------------------------------------------ algorithmic work per step
    @property
    def fused_tokens(self) -> int:
        return self.M

    def gemm_flops(self) -> int:
        return 2 * self.M * self.K * self.shape.hidden  # per GEMM launch (fwd, and again for dW)

    def gather_bytes(self) -> int:
        return 2 * self.M * self.K * 2

    def splice_bytes(self) -> int:
        return 2 * self.M * self.shape.hidden * 2 + 16 * self.shape.batch * self.S

    # ------------------------------------------------------------------ the step
    def enable_kernel_timing(self, names=("gather", "proj_fwd", "splice_fwd", "splice_bwd", "proj_bwd_dw",
                                          "proj_bwd_dw_v", "colsum")):
        self.events = {n: [] for n in names}

    def _timed(self, name, fn):
        if self.nvtx:  # AVC_NVTX=1: one NVTX range per launch, for `ncu --nvtx --nvtx-include "proj_fwd/"` and timelines
            torch.cuda.nvtx.range_push(name)
            try:
                self._timed_inner(name, fn)
            finally:
                torch.cuda.nvtx.range_pop()
        else:
            self._timed_inner(name, fn)

    def _timed_inner(self, name, fn):
        if self.events is None or name not in self.events:
            fn()
            return
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        self.events[name].append((s, e))

    def _fork(self):
        """Point on the main stream after which side-stream work may start."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return ev

    def _on_side(self, fork, name, fn):
        """Run a small HBM-bound kernel on the side stream, concurrently with the tensor-bound GEMM that was just
        enqueued on the main stream after `fork` (the GEMM's CTAs leave threads, registers and HBM bandwidth free on
        every SM; enqueueing the GEMM FIRST lets its CTAs take their SMs before the small kernel's blocks fill them).
        Returns the event the main stream must wait for before anything consumes the kernel's output."""
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        side = self._side_stream
        with torch.cuda.stream(side):
            side.wait_event(fork)
            self._timed(name, fn)
            done = torch.cuda.Event()
            done.record(side)
        return done

    def _named_params(self):
        named = []
        if self.use_a:
            named += [("audio_connector.linear.weight", self.wa), ("audio_connector.linear.bias", self.ba)]
        if self.use_v:
            named += [("video_connector.linear.weight", self.wv), ("video_connector.linear.bias", self.bv)]
        return named

    def _pack_weights(self):
        col = 0
        if self.use_a:
            L.pack_weight(self.wa, self.wp[:, :self.Ka], self.sa)
            col = self.Ka
        if self.use_v:
            L.pack_weight(self.wv, self.wp[:, col:], self.sv)

    def attach_optimizer(self, **adamw_kwargs):
        """Make `train_step()` a full trainer step (clip_whisper_trainer.py:453-464): forward, backward (+ all-reduce),
        global-norm clip and AdamW on the flat gradient bucket.  The AdamW kernel also writes the bf16, fusion-scaled
        copy of each updated weight straight into the packed GEMM operand, so the forward's two pack launches go away."""
        from .trainer_step import ConnectorAdamW

        opt = ConnectorAdamW(self._named_params(), bucket=self.bucket, **adamw_kwargs)
        if self.use_a:
            opt.attach_packed("audio_connector.linear.weight", self.wp[:, :self.Ka], self.sa)
        if self.use_v:
            opt.attach_packed("video_connector.linear.weight", self.wp[:, (self.Ka if self.use_a else 0):], self.sv)
        self._pack_weights()   # the pack of the current weights; from here on the optimizer keeps it current
        self.optimizer = opt
        return opt

    def detach_optimizer(self):
        self.optimizer = None

    def train_step(self, other_sumsq=None):
        """forward + backward (+ gradient all-reduce) + clip + AdamW; `other_sumsq` = the squared gradient norm of the
        non-connector parameters (LoRA), a device scalar, for the reference's GLOBAL clip."""
        if self.optimizer is None:
            raise L.ConnectorError("train_step() needs attach_optimizer() first")
        self.forward()
        g = self.backward(True)
        self.optimizer.step(other_sumsq=other_sumsq)
        return g

    def forward(self):
        p = self.plan
        if self.optimizer is None:
            self._pack_weights()
        if not self.direct:
            self._timed("gather", lambda: L.gather_fwd(self.audio, self.video, p.audio_stride, p.video_stride,
                                                       self.shape.batch, self.N, self.A, self.flags,
                                                       tok_offset=self.tok_offset, audio_valid=self.audio_valid,
                                                       video_valid=self.video_valid,
                                                       audio_repeat=p.audio_repeat, video_repeat=p.video_repeat))
        if self.use_a and self.use_v:
            b0, b1, s0, s1 = self.ba, self.bv, self.sa, self.sv
        elif self.use_a:
            b0, b1, s0, s1 = self.ba, None, self.sa, 0.0
        else:
            b0, b1, s0, s1 = None, self.bv, 0.0, self.sv
        if self.direct:
            xs = ([self.audio.view(self.M, self.Ka)] if self.use_a else []) + \
                 ([self.video.view(self.M, self.Kv)] if self.use_v else [])
            wsegs = ([self.wp[:, :self.Ka]] if self.use_a else []) + ([self.wp[:, self.Ka:]] if self.use_v else [])
            fork = self._fork() if self.side_streams else None
            self._timed("proj_fwd", lambda: L.proj_fwd(xs, wsegs, self.emb_av, bias0=b0, bias1=b1, bias_scale0=s0,
                                                       bias_scale1=s1))
            if self.side_streams:
                done = self._on_side(fork, "splice_fwd", lambda: L.splice_fwd(self.sp_text, None, self.emb))
                torch.cuda.current_stream().wait_event(done)
            else:
                self._timed("splice_fwd", lambda: L.splice_fwd(self.sp_text, None, self.emb))
            return self.emb, self.mask, self.labels_out
        else:
            self._timed("proj_fwd", lambda: L.proj_fwd([self.A], [self.wp], self.Y, bias0=b0, bias1=b1,
                                                       bias_scale0=s0, bias_scale1=s1, row_flags=self.flags))
        self._timed("splice_fwd", lambda: L.splice_fwd(self.sp, self.Y, self.emb))
        return self.emb, self.mask, self.labels_out

    def backward(self, allreduce: bool = True):
        g = self.bucket
        B, N, P = self.shape.batch, self.N, self.shape.prompt_len
        # data parallel: the 1 / world factor of the gradient mean is folded into the dW / bias-sum epilogues, so the
        # collective is a plain SUM (NVLS-capable) instead of AVG.  The fused launch only sums, so it always pre-scales
        # (AVC_PRESCALE_GRADS=0 switches the NCCL schedules to ReduceOp.AVG): either way bucket.flat ends as the MEAN.
        fused = allreduce and self.fused_allreduce
        inv = 1.0 / g.world_size() if (allreduce and (self.prescale_grads or fused)) else 1.0
        ga, gv = self.sa * inv, self.sv * inv
        pre = inv != 1.0
        dba = g["audio_connector.linear.bias"] if self.use_a else None
        dbv = g["video_connector.linear.bias"] if self.use_v else None
        if self.direct:
            dy, base = self.d_emb, P
            xa = self.audio.view(B, N, self.Ka) if self.use_a else None
            xv = self.video.view(B, N, self.Kv) if self.use_v else None
            cs = dict(dy_row_base=P, sum_rows=N)
        else:
            self._timed("splice_bwd", lambda: L.splice_bwd(self.sp, self.d_emb, self.dY))
            dy, base = self.dY, 0
            xa = self.A[:, :self.Ka] if self.use_a else None
            xv = self.A[:, self.Ka:] if self.use_v else None
            cs = dict(row_flags=self.flags)
        bias = None
        if self.bias_in_gemm:
            if self.direct:
                present = L.present_operand(B, N, self.device)
            else:
                if self._present is None:  # static engine: the gather writes the same flags every step
                    self._present = L.present_operand(1, self.M, self.device, row_flags=self.flags)
                present = self._present
            bias = (present, dba, dbv, ga, gv)
        if fused:
            return self._backward_fused_allreduce(dy, base, xa, xv, dba, dbv, ga, gv, cs, bias)
        overlap = allreduce and self.overlap_comm and g.world_size() > 1 and self.use_a and self.use_v
        # the bias column sums only read d(inputs_embeds): run them under the dW GEMM
        side_cs = self.side_streams and self.direct and bias is None
        fork = self._fork() if side_cs else None
        cs_done = None

        def bias_sums():
            L.colsum(dy, dba, dbv, self.colsum_ws, alpha0=ga, alpha1=gv, **cs)

        if not overlap:
            xs = ([xa] if self.use_a else []) + ([xv] if self.use_v else [])
            dws = ([g["audio_connector.linear.weight"]] if self.use_a else []) + \
                  ([g["video_connector.linear.weight"]] if self.use_v else [])
            al = ([ga] if self.use_a else []) + ([gv] if self.use_v else [])
            self._timed("proj_bwd_dw", lambda: L.proj_bwd_dw(dy, xs, dws, al, dy_row_base=base, bias=bias))
            if bias is not None:
                pass  # db came out of the dW launch
            elif side_cs:
                torch.cuda.current_stream().wait_event(self._on_side(fork, "colsum", bias_sums))
            else:
                self._timed("colsum", bias_sums)
            if allreduce:
                g.allreduce(prescaled=pre)
            return g
        # Data-parallel overlap: the audio weight gradient (2/3 of the bucket at cfg2) is all-reduced on a side stream
        # while the video weight gradient and the bias sums are still being computed on `comm_reserve_sms` fewer SMs;
        # the rest of the bucket follows.  Same arithmetic, same results as the single all-reduce.
        main = torch.cuda.current_stream()
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        comm = self._comm_stream
        self._timed("proj_bwd_dw", lambda: L.proj_bwd_dw(dy, [xa], [g["audio_connector.linear.weight"]], [ga],
                                                         dy_row_base=base))
        if bias is not None:   # overlapped NCCL schedule: the bias sums ride on the video-weight launch below
            pass
        elif side_cs:
            cs_done = self._on_side(fork, "colsum", bias_sums)
        e1 = torch.cuda.Event()
        e1.record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(e1)
            g.allreduce_span("audio_connector.linear.weight", "audio_connector.linear.weight", prescaled=pre)
        sms = self._num_sms - self.comm_reserve_sms
        self._timed("proj_bwd_dw_v", lambda: L.proj_bwd_dw(dy, [xv], [g["video_connector.linear.weight"]], [gv],
                                                           dy_row_base=base, max_sms=sms, bias=bias))
        if bias is not None:
            pass
        elif cs_done is not None:
            main.wait_event(cs_done)
        else:
            self._timed("colsum", bias_sums)
        e2 = torch.cuda.Event()
        e2.record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(e2)
            g.allreduce_span("video_connector.linear.weight", "video_connector.linear.bias", prescaled=pre)
            e3 = torch.cuda.Event()
            e3.record(comm)
        main.wait_event(e3)
        return g

    def _backward_fused_allreduce(self, dy, base, xa, xv, dba, dbv, ga, gv, cs, bias):
        """dW GEMM + db + all-reduce of the whole bucket in ONE launch: the bias gradients come from the launch's own
        bias work items, whose epilogue flags them ready; the comm warps reduce them together with the weight tiles.
        (AVC_BIAS_IN_GEMM=0: the stand-alone bias-sum kernel runs on the side stream under the GEMM and flags them.)"""
        g = self.bucket
        comm = g.peer.next_epoch()
        ex = [t for t in (dba, dbv) if t is not None]
        if bias is not None:
            xs = ([xa] if self.use_a else []) + ([xv] if self.use_v else [])
            dws = ([g["audio_connector.linear.weight"]] if self.use_a else []) + \
                  ([g["video_connector.linear.weight"]] if self.use_v else [])
            al = ([ga] if self.use_a else []) + ([gv] if self.use_v else [])
            self._timed("proj_bwd_dw", lambda: L.proj_bwd_dw_allreduce(dy, xs, dws, al, comm, dy_row_base=base,
                                                                       bias=bias))
            return g

        def bias_sums():  # one launch; its last CTA flags the sums ready for this epoch
            L.colsum(dy, dba, dbv, self.colsum_ws, alpha0=ga, alpha1=gv, comm=comm, **cs)

        xs = ([xa] if self.use_a else []) + ([xv] if self.use_v else [])
        dws = ([g["audio_connector.linear.weight"]] if self.use_a else []) + \
              ([g["video_connector.linear.weight"]] if self.use_v else [])
        al = ([ga] if self.use_a else []) + ([gv] if self.use_v else [])

        def gemm():
            self._timed("proj_bwd_dw", lambda: L.proj_bwd_dw_allreduce(
                dy, xs, dws, al, comm, extra0=ex[0], extra1=ex[1] if len(ex) > 1 else None, dy_row_base=base))

        # AVC_BIAS_IN_GEMM=0 (A/B knob): the stand-alone bias-sum kernel runs first and flags its sums, the fused launch
        # follows on the same stream (no co-residency of the two kernels is assumed)
        self._timed("colsum", bias_sums)
        gemm()
        return g

    def step(self, allreduce: bool = True):
        self.forward()
        return self.backward(allreduce)

    def capture_graph(self):
        """The whole step (every launch of forward + backward, side-stream work included) as ONE CUDA graph; replay it
        with `.replay()`.  For launch-bound shapes (BASELINE configs[0]: batch 2, 1000 fused tokens, ~7 launches of
        10 - 40 us) the host cost of the step drops from ~0.2 ms to one graph launch.  Single process only: the fused
        all-reduce needs a fresh epoch number per launch, which a captured kernel argument cannot carry."""
        if self.fused_allreduce:
            raise L.ConnectorError("capture_graph: the data-parallel step is not capturable (per-launch epochs)")
        self.events = None
        for _ in range(2):   # lazy module loading, workspace allocation, pack / plan caches: outside the capture
            self.step(allreduce=False)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.step(allreduce=False)
        return graph

    def load_inputs_from_host(self, audio_h: Optional[torch.Tensor], video_h: Optional[torch.Tensor],
                              ids_h: torch.Tensor, labels_h: torch.Tensor):
        """H2D copies of one step's inputs (pinned host tensors) into the resident buffers, on the current stream."""
        if self.use_a:
            self.audio.copy_(audio_h, non_blocking=True)
        if self.use_v:
            self.video.copy_(video_h, non_blocking=True)
        self.input_ids.copy_(ids_h, non_blocking=True)
        self.labels_in.copy_(labels_h, non_blocking=True)


class HostFeeder:
    """Double-buffered host -> device input pipeline for the connector step.

    The reference trainer copies each batch to the device synchronously right before the step
    (`.to(device)`, clip_whisper_trainer.py:655-657, pin_memory=False).  Here the batch for step i+1 is copied from
    PINNED host memory on a side stream while step i computes; a slot is only overwritten after the step that read
    it (forward and backward) has finished on the compute stream.
    """

    def __init__(self, device, slots: int = 2):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [dict(bufs=None, ready=torch.cuda.Event(), free=None) for _ in range(slots)]
        self.head = 0   # next slot to fill
        self.tail = 0   # next slot to hand out
        self.bytes_per_batch = 0

    def prefetch(self, host_tensors):
        """Start the H2D copy of one batch (a sequence of pinned CPU tensors, None entries allowed)."""
        slot = self.slots[self.head % len(self.slots)]
        self.head += 1
        if slot["bufs"] is None:
            slot["bufs"] = [None if t is None else torch.empty(t.shape, dtype=t.dtype, device=self.device)
                            for t in host_tensors]
        with torch.cuda.stream(self.copy_stream):
            if slot["free"] is not None:
                self.copy_stream.wait_event(slot["free"])  # the step that last read this slot is done
            nbytes = 0
            for dst, src in zip(slot["bufs"], host_tensors):
                if src is None:
                    continue
                if not src.is_pinned():
                    raise ValueError("HostFeeder needs pinned host tensors (torch.Tensor.pin_memory())")
                dst.copy_(src, non_blocking=True)
                nbytes += src.nbytes
            slot["ready"].record(self.copy_stream)
        self.bytes_per_batch = nbytes

    def take(self):
        """Device tensors of the oldest prefetched batch; the current stream waits for its copy."""
        slot = self.slots[self.tail % len(self.slots)]
        self.tail += 1
        torch.cuda.current_stream().wait_event(slot["ready"])
        self._last = slot
        return slot["bufs"]

    def release(self):
        """Call after the step (forward + backward) that consumed the last `take()` has been enqueued."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._last["free"] = ev


class GraphedEncoder:
    """Forward-only connector call (decode / generate, clip_whisper_model.py:1301-1340) as ONE CUDA-graph launch.

    Batch-1 decoding is launch-latency bound: eager `fused_connector` costs ~0.13 ms of host work (ctypes calls, tensor
    maps, allocator) for ~20 us of kernels.  The graph is captured once per input signature on static input buffers;
    a call copies the new inputs into them (device-to-device, stream-ordered) and replays the graph: the kernel
    parameters -- including the TMA tensor maps, which live in kernel parameter space -- are baked into the graph, so
    nothing is re-encoded.  Outputs are static tensors that the next call overwrites.

        enc = GraphedEncoder(lambda a, v, ids: fused_connector(a, v, wa, ba, wv, bv, plan, prompt_ids=ids, ...))
        emb, mask, _ = enc(audio_feats, video_feats, prompt_ids)

    The captured callable must be free of host synchronisation (`check=False`) and run under `torch.no_grad()`; weights
    it closes over are read at replay time (the bf16 pack is part of the graph when the pack cache is cold, so call
    `connector_ops.invalidate_pack_cache()` + `reset()` after the weights change)."""

    def __init__(self, fn, warmup: int = 2):
        self.fn = fn
        self.warmup = warmup
        self._graphs = {}

    @staticmethod
    def _key(inputs):
        return tuple(None if t is None else (tuple(t.shape), t.dtype, t.device) for t in inputs)

    def reset(self) -> None:
        self._graphs.clear()

    def __call__(self, *inputs):
        key = self._key(inputs)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = [None if t is None else t.clone() for t in inputs]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(self.warmup):   # lazy module loading, pack cache, allocator warm-up: outside the capture
                    self.fn(*static_in)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(graph):
                static_out = self.fn(*static_in)
            entry = (graph, static_in, static_out)
            self._graphs[key] = entry
        graph, static_in, static_out = entry
        for dst, src in zip(static_in, inputs):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        graph.replay()
        return static_out
