"""2+ GPU check of the data-parallel step (run under torchrun): the overlapped all-reduce schedule gives the same
reduced gradients as one all-reduce after the backward, and every rank ends with identical buckets."""
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
shape = StepShape(batch=4, audio_frames=400, video_frames=200, audio_dim=256, video_dim=128, hidden=512, prompt_len=8,
                  vocab=1000)
plan = pkg.FusePlan(fusion="concat", audio_stride=4, video_stride=2, max_seq_len=4096)
res = []
for overlap in (False, True):
    eng = ConnectorStep(shape, plan, dev, seed=10 + rank)   # different data per rank, same weights needed:
    ref = ConnectorStep(shape, plan, dev, seed=10)          # take rank 0's parameters everywhere
    for n in ("wa", "wv", "ba", "bv"):
        getattr(eng, n).copy_(getattr(ref, n))
    eng.overlap_comm = overlap
    eng.step()
    torch.cuda.synchronize()
    res.append(eng.bucket.flat.clone())
    local_only = ConnectorStep(shape, plan, dev, seed=10 + rank)
    for n in ("wa", "wv", "ba", "bv"):
        getattr(local_only, n).copy_(getattr(ref, n))
    local_only.step(allreduce=False)
    torch.cuda.synchronize()
    gathered = [torch.empty_like(local_only.bucket.flat) for _ in range(world)]
    dist.all_gather(gathered, local_only.bucket.flat)
    mean = torch.stack(gathered).double().mean(0)
    err = float((res[-1].double() - mean).abs().max() / mean.abs().max())
    assert err < 1e-6, (overlap, err)
assert torch.equal(res[0], res[1]) or float((res[0] - res[1]).abs().max()) <= 1e-6 * float(res[0].abs().max())
chk = [torch.empty_like(res[1]) for _ in range(world)]
dist.all_gather(chk, res[1])
assert all(torch.equal(chk[0], c) for c in chk), "ranks disagree on the reduced gradients"
if rank == 0:
    print("dp_check ok: world", world)
dist.destroy_process_group()
