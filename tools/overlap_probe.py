"""Timeline of the overlapped backward (torchrun, N >= 2): when does the first all-reduce actually run?"""
import os, sys, torch, torch.distributed as dist
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry; entry.build()
import audio_visual_llm_b200 as pkg
from audio_visual_llm_b200.engine import ConnectorStep, StepShape
from audio_visual_llm_b200 import _lib as L
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = StepShape(batch=32, audio_frames=1500, video_frames=750, audio_dim=1024, video_dim=1024, hidden=4096)
plan = pkg.FusePlan(fusion="concat", audio_stride=4, video_stride=2, max_seq_len=1536)
eng = ConnectorStep(shape, plan, dev, seed=rank)
g = eng.bucket
B, N, P = 32, eng.N, 16
xa, xv = eng.audio.view(B, N, eng.Ka), eng.video.view(B, N, eng.Kv)
main = torch.cuda.current_stream(); comm = torch.cuda.Stream()
reserve = int(os.environ.get("AVC_COMM_RESERVE_SMS", "16"))
def ev(): return torch.cuda.Event(enable_timing=True)
acc = {}
for it in range(30):
    eng.forward()
    t0 = ev(); t0.record(main)
    L.proj_bwd_dw(eng.d_emb, [xa], [g["audio_connector.linear.weight"]], [1.0], dy_row_base=P)
    t1 = ev(); t1.record(main)
    with torch.cuda.stream(comm):
        comm.wait_event(t1)
        a0 = ev(); a0.record(comm)
        g.allreduce_span("audio_connector.linear.weight", "audio_connector.linear.weight")
        a1 = ev(); a1.record(comm)
    L.proj_bwd_dw(eng.d_emb, [xv], [g["video_connector.linear.weight"]], [1.0], dy_row_base=P, max_sms=148 - reserve)
    t2 = ev(); t2.record(main)
    with torch.cuda.stream(comm):
        comm.wait_event(t2)
        g.allreduce_span("video_connector.linear.weight", "video_connector.linear.bias")
        a2 = ev(); a2.record(comm)
    main.wait_event(a2)
    t3 = ev(); t3.record(main)
    torch.cuda.synchronize()
    if it >= 10:
        for k, v in dict(dWa=t0.elapsed_time(t1), dWv=t1.elapsed_time(t2), ar1_start_after_dWa=t1.elapsed_time(a0),
                         ar1=a0.elapsed_time(a1), ar1_end_after_dWa=t1.elapsed_time(a1), ar2_end_after_dWv=t2.elapsed_time(a2),
                         total=t0.elapsed_time(t3)).items():
            acc[k] = acc.get(k, 0) + v / 20
if rank == 0:
    print("reserve", reserve, {k: round(v, 4) for k, v in acc.items()}, file=sys.stderr)
dist.destroy_process_group()
