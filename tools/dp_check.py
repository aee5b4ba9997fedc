"""2+ GPU check of the data-parallel step (run under torchrun).  Three schedules of the same step must give the same
reduced gradients, and every rank must end with identical buckets:
  fused   -- the dW GEMM all-reduces the bucket itself over peer-mapped memory (avc_proj_bwd_dw_allreduce; default)
  nccl    -- one NCCL all-reduce after the backward
  overlap -- NCCL all-reduce of the audio-weight span under the video-weight dW launch
Also runs the fused step several times back to back (epochs) and prints its time next to the NCCL schedule's."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import audio_visual_llm_b200 as pkg  # noqa: E402
from audio_visual_llm_b200.engine import ConnectorStep, StepShape  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
big = "--big" in sys.argv
if big:   # BASELINE cfg2 per GPU
    shape = StepShape(batch=32, audio_frames=1500, video_frames=750, audio_dim=1024, video_dim=1024, hidden=4096)
else:
    shape = StepShape(batch=4, audio_frames=400, video_frames=200, audio_dim=256, video_dim=128, hidden=512,
                      prompt_len=8, vocab=1000)
plan = pkg.FusePlan(fusion="concat", audio_stride=4, video_stride=2, max_seq_len=4096)
ref = ConnectorStep(shape, plan, dev, seed=10, fused_allreduce=False)  # rank 0's parameters everywhere


def make(fused, overlap=False, multimem=False):
    os.environ["AVC_COMM_MULTIMEM"] = "1" if multimem else "0"
    eng = ConnectorStep(shape, plan, dev, seed=10 + rank, fused_allreduce=fused)  # different data per rank
    for n in ("wa", "wv", "ba", "bv"):
        getattr(eng, n).copy_(getattr(ref, n))
    eng.overlap_comm = overlap
    return eng


local_only = make(False)
local_only.step(allreduce=False)
torch.cuda.synchronize()
gathered = [torch.empty_like(local_only.bucket.flat) for _ in range(world)]
dist.all_gather(gathered, local_only.bucket.flat)
mean = torch.stack(gathered).double().mean(0)
del gathered
res = {}
variants = [("fused", True, False, False), ("nccl", False, False, False), ("overlap", False, True, False)]
if "--multimem" in sys.argv:  # the fused launch reducing through an NVSwitch multicast mapping (multimem.ld_reduce / st)
    variants.insert(1, ("fused_multimem", True, False, True))
if "--only-fused" in sys.argv:
    variants = [v for v in variants if v[1]]
for name, fused, overlap, mm in variants:
    eng = make(fused, overlap, mm)
    if mm and rank == 0:
        print(f"dp_check {name}: multicast transport active = {eng.bucket.peer.mc is not None}", file=sys.stderr)
    for _ in range(3):      # several epochs: flags are never reset between launches
        eng.bucket.flat.fill_(float("nan"))
        eng.bucket.zero_padding()
        eng.step()
    torch.cuda.synchronize()
    if eng.bucket.peer is not None:
        eng.bucket.peer.check()
    res[name] = eng.bucket.flat.clone()
    err = float((res[name].double() - mean).abs().max() / mean.abs().max())
    assert err < 1e-6, (name, err)
    chk = [torch.empty_like(res[name]) for _ in range(world)]
    dist.all_gather(chk, res[name])
    assert all(torch.equal(chk[0], c) for c in chk), f"{name}: ranks disagree on the reduced gradients"
    # timing, max over ranks
    dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    t0.record()
    for _ in range(n):
        eng.step()
    t1.record()
    torch.cuda.synchronize()
    t = torch.tensor([t0.elapsed_time(t1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if eng.bucket.peer is not None:
        eng.bucket.peer.check()
        eng.bucket.peer.close()
    if rank == 0:
        print(f"dp_check {name}: max rel err vs fp64 mean {err:.2e}, {float(t):.4f} ms / step", file=sys.stderr)
    del eng
    torch.cuda.empty_cache()
base_name = "nccl" if "nccl" in res else "fused"
scale = float(res[base_name].abs().max())
for a in [v[0] for v in variants if v[0] != base_name]:
    assert float((res[a] - res[base_name]).abs().max()) <= 1e-6 * scale, a
if rank == 0:
    print("dp_check ok: world", world)
dist.destroy_process_group()
