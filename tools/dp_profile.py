"""Timeline of the fused dW + all-reduce launch on N GPUs (run under torchrun): %globaltimer stamps taken inside the kernel
(avc_debug_gemm_profile) on every rank -- last MMA commit, last epilogue, when the comm warps saw the last round's flags,
finished their units, passed the system fence and the done handshake -- next to the plain dW launch of the same rank."""
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

L = pkg._lib
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = StepShape(batch=32, audio_frames=1500, video_frames=750, audio_dim=1024, video_dim=1024, hidden=4096)
plan = pkg.FusePlan(fusion="concat", audio_stride=4, video_stride=2, max_seq_len=4096)


def timeline(eng, label):
    for _ in range(10):
        eng.step()
    torch.cuda.synchronize()
    dist.barrier()
    eng.forward()
    prof = torch.zeros(3 * 148, 8, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    dist.barrier()
    L.debug_gemm_profile(prof)
    eng.backward()
    torch.cuda.synchronize()
    L.debug_gemm_profile(None)
    ts = prof.cpu()[296:].double()
    t0 = ts[:, 0].min()

    def us(col, red="max"):
        v = ts[:, col]
        v = v[v > 0]
        if v.numel() == 0:
            return float("nan")
        return float(((v.max() if red == "max" else v.min()) - t0) / 1e3)

    p = prof.cpu().double()[:148]
    msg = ("%s rank %d: us since launch: last MMA commit %.1f | last epilogue done %.1f | comm saw last-round flags %.1f .. %.1f | "
           "units done %.1f | after fence %.1f | after handshake %.1f || MMA wait-full %.0f kc, wait-tempty %.0f kc, producer loop %.0f kc"
           % (label, rank, us(5), us(1), us(2, "min"), us(2), us(3), us(6), us(4), p[0::2, 1].mean() / 1e3,
              p[0::2, 2].mean() / 1e3, p[:, 6].mean() / 1e3))
    out = [None] * world
    dist.all_gather_object(out, msg)
    if rank == 0:
        for m in out:
            print(m, flush=True)


for label, fused, mm in (("plain+nccl", False, "0"), ("fused peer", True, "0"), ("fused multimem", True, "1")):
    os.environ["AVC_COMM_MULTIMEM"] = mm
    eng = ConnectorStep(shape, plan, dev, seed=10 + rank, fused_allreduce=fused)
    timeline(eng, label)
    if eng.bucket.peer is not None:
        eng.bucket.peer.check()
        eng.bucket.peer.close()
    del eng
    torch.cuda.empty_cache()
dist.destroy_process_group()
