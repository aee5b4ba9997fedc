"""Single-GPU cost of the comm warps: plain dW GEMM vs the fused dW + all-reduce kernel at world = 1 (the protocol runs
against itself: every tile is re-read and re-written locally), BASELINE cfg2 shapes.  AVC_COMM_POLL_NS is read per launch."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import audio_visual_llm_b200 as pkg  # noqa: E402

L = pkg._lib
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
B, R, P, H, Ka, Kv = 32, 375, 16, 4096, 4096, 2048
g = torch.Generator().manual_seed(0)
dy = torch.randn(B, P + R, H, generator=g).to(torch.bfloat16).to(dev)
xa = torch.randn(B, R, Ka, generator=g).to(torch.bfloat16).to(dev)
xv = torch.randn(B, R, Kv, generator=g).to(torch.bfloat16).to(dev)
n = H * (Ka + Kv) + 2 * H
bptr, fptr = L.comm_alloc(n * 4), L.comm_alloc(L.comm_flag_bytes())
bucket = L.as_tensor(bptr, n, torch.float32, dev)
dws = [bucket[:H * Ka].view(H, Ka), bucket[H * Ka:H * (Ka + Kv)].view(H, Kv)]
ex0, ex1 = bucket[H * (Ka + Kv):H * (Ka + Kv) + H], bucket[H * (Ka + Kv) + H:]
status = torch.zeros(1, dtype=torch.int32, device=dev)
epoch = 0


def comm():
    global epoch
    epoch += 1
    c = L.AvcComm()
    c.world, c.rank, c.epoch = 1, 0, epoch
    c.bucket[0], c.flags[0] = bptr, fptr
    c.status, c.timeout_ns, c.bucket_bytes = status.data_ptr(), int(5e9), n * 4
    return c


present = L.present_operand(B, R, dev)


def plain():
    L.proj_bwd_dw(dy, [xa, xv], dws, [0.5, 0.5], dy_row_base=P)


def plain_db():   # dW + db in one launch (bias work items)
    L.proj_bwd_dw(dy, [xa, xv], dws, [0.5, 0.5], dy_row_base=P, bias=(present, ex0, ex1, 0.5, 0.5))


def fused():      # round-1 protocol: the bias sums come from another kernel and are flagged by avc_comm_signal_extra
    c = comm()
    L.comm_signal_extra(c, H, H)
    L.proj_bwd_dw_allreduce(dy, [xa, xv], dws, [0.5, 0.5], c, extra0=ex0, extra1=ex1, dy_row_base=P)


def fused_db():   # dW + db + all-reduce in one launch
    L.proj_bwd_dw_allreduce(dy, [xa, xv], dws, [0.5, 0.5], comm(), dy_row_base=P, bias=(present, ex0, ex1, 0.5, 0.5))


def profile(name, fn):
    prof = torch.zeros(3 * 148, 8, dtype=torch.int64, device=dev)
    for _ in range(3):
        fn()
    L.debug_gemm_profile(prof)
    fn()
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

    print("  %-10s us since launch: last MMA commit %.1f | last epilogue done %.1f | comm: first last-round flag seen %.1f, "
          "last %.1f | units done %.1f | after fence %.1f | after handshake %.1f | kernel end %.1f"
          % (name, us(5), us(1), us(2, "min"), us(2), us(3), us(6), us(4), us(7)))
    p = prof.cpu().double()[:148]
    lead = p[0::2]
    print("  %-10s kernel %.0f kc | producer wait-empty %.0f | MMA wait-full %.0f wait-tempty %.0f | epilogue body %.1f kc / item"
          % (name, p[:, 6].mean() / 1e3, p[:, 0].mean() / 1e3, lead[:, 1].mean() / 1e3, lead[:, 2].mean() / 1e3,
             (p[:, 4] / p[:, 5].clamp_min(1)).mean() / 1e3))


def bench(fn, iters=60):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


if "--ncu" in sys.argv:   # short launch sequence for an ncu capture: colsum, plain dW, fused dW + all-reduce (world = 1)
    ws = L.colsum_workspace(H, dev)
    for _ in range(2):
        L.colsum(dy, ex0, ex1, ws, dy_row_base=P, sum_rows=R)
        plain()
        fused()
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    sys.exit(0)
for _ in range(3):   # let the GPU reach its sustained (power-capped) state first: cold numbers are ~20 % faster
    bench(plain, 200)
print("plain dW            : %.4f ms" % bench(plain))
print("plain dW + db       : %.4f ms" % bench(plain_db))
for ns in (200, 2000):
    os.environ["AVC_COMM_POLL_NS"] = str(ns)
    print("fused, poll %5d ns : %.4f ms" % (ns, bench(fused)))
    print("fused + db, poll %5d: %.4f ms" % (ns, bench(fused_db)))
os.environ["AVC_COMM_POLL_NS"] = "200"
print("plain dW            : %.4f ms" % bench(plain))
for name, fn in (("plain", plain), ("plain_db", plain_db), ("fused", fused), ("fused_db", fused_db)):
    profile(name, fn)
assert int(status.item()) == 0
