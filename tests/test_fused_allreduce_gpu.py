"""dW GEMM with the gradient all-reduce fused into the same kernel (avc_proj_bwd_dw_allreduce), on ONE GPU.

The protocol (per-item ready flags, owner = item % world, peer loads in rank order, peer stores to every rank, done
flags, epochs) is exercised with `world` virtual ranks: every rank has its own bucket + flag area in device memory and
its own launch on its own stream over a share of the SMs, so the launches are co-resident and talk to each other
exactly as the per-GPU launches of a real data-parallel job do (there the peers' blocks are mapped with CUDA IPC and
the loads / stores travel over NVLink: tools/dp_check.py runs that under torchrun on 2+ GPUs).

Bars: the reduced gradients are bit-identical on every rank (they are computed once, by the item's owner, in rank
order) and equal the fp64 sum of the per-rank references to fp32 rounding (tolerance at the assert).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def bf16_randn(gen, *shape):
    return torch.randn(*shape, generator=gen, dtype=torch.float32).to(torch.bfloat16)


class VirtualRanks:
    """`world` buckets / flag areas on one device and the avc_comm descriptor of each virtual rank."""

    def __init__(self, L, world, nfloats, dev, timeout_s=4.0):
        self.L, self.world, self.dev = L, world, dev
        self.bucket_ptrs = [L.comm_alloc(nfloats * 4) for _ in range(world)]
        self.flag_ptrs = [L.comm_alloc(L.comm_flag_bytes()) for _ in range(world)]
        self.buckets = [L.as_tensor(p, nfloats, torch.float32, dev) for p in self.bucket_ptrs]
        self.status = torch.zeros(world, dtype=torch.int32, device=dev)
        self.timeout_ns = int(timeout_s * 1e9)
        self.nfloats = nfloats
        self.epoch = 0

    def descriptors(self):
        self.epoch += 1
        out = []
        for r in range(self.world):
            c = self.L.AvcComm()
            c.world, c.rank, c.epoch = self.world, r, self.epoch
            for p in range(self.world):
                c.bucket[p] = self.bucket_ptrs[p]
                c.flags[p] = self.flag_ptrs[p]
            c.status = self.status[r:r + 1].data_ptr()
            c.timeout_ns = self.timeout_ns
            c.bucket_bytes = self.nfloats * 4
            out.append(c)
        return out

    def free(self):
        torch.cuda.synchronize()
        self.buckets = None
        for p in self.bucket_ptrs + self.flag_ptrs:
            self.L.comm_free(p)


def run_virtual(L, dev, world, B, R, H, Ka, Kv, base, epochs=2, seed=0, sms_per_rank=None, use_colsum=False):
    g = torch.Generator().manual_seed(seed)
    pad = 64
    offs = [0, H * Ka]
    off_b0 = offs[1] + H * Kv
    off_b0 = (off_b0 + pad - 1) // pad * pad
    off_b1 = off_b0 + (H + pad - 1) // pad * pad
    nfloats = off_b1 + (H + pad - 1) // pad * pad
    vr = VirtualRanks(L, world, nfloats, dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    side = [torch.cuda.Stream(device=dev) for _ in range(world)]
    cs_ws = [L.colsum_workspace(H, dev) for _ in range(world)]
    num_sms = torch.cuda.get_device_properties(dev).multi_processor_count
    max_sms = sms_per_rank if sms_per_rank is not None else (0 if world == 1 else (num_sms // world) & ~1)
    try:
        for ep in range(epochs):
            dys = [bf16_randn(g, B, base + R, H) for _ in range(world)]
            xas = [bf16_randn(g, B, R, Ka) for _ in range(world)]
            xvs = [bf16_randn(g, B, R, Kv) for _ in range(world)] if Kv else None
            biases = [torch.randn(2, H, generator=g) for _ in range(world)]
            al = [0.5, 0.25][: 2 if Kv else 1]
            def ref(scale, xs):  # fp64 on the device (cuBLAS), independent of the kernel under test
                return sum(scale * torch.einsum("brh,brk->hk", d[:, base:].to(dev).double(), x.to(dev).double())
                           for d, x in zip(dys, xs)).cpu()

            ref_a = ref(al[0], xas)
            ref_v = ref(al[1], xvs) if Kv else None
            ref_b = sum(b.double() for b in biases)
            if use_colsum:  # the extra ranges are the bias gradients, produced by avc_colsum_comm next to the GEMM
                ref_b = torch.stack([sum(a * d[:, base:].double().sum((0, 1)) for d in dys) for a in (0.5, 0.25)])
            comms = vr.descriptors()
            dev_in = []
            for r in range(world):
                bk = vr.buckets[r]
                bk.fill_(float("nan"))
                bk[off_b0:off_b0 + H].copy_(biases[r][0])   # stands for this rank's bias column sums
                bk[off_b1:off_b1 + H].copy_(biases[r][1])
                dev_in.append((dys[r].to(dev), xas[r].to(dev), xvs[r].to(dev) if Kv else None))
            torch.cuda.synchronize()
            for r in range(world):
                bk = vr.buckets[r]
                dy, xa, xv = dev_in[r]
                dws = [bk[:H * Ka].view(H, Ka)] + ([bk[offs[1]:offs[1] + H * Kv].view(H, Kv)] if Kv else [])
                with torch.cuda.stream(streams[r]):
                    if not use_colsum:
                        L.comm_signal_extra(comms[r], H, H)
                    L.proj_bwd_dw_allreduce(dy, [xa] + ([xv] if Kv else []), dws, al, comms[r],
                                            extra0=bk[off_b0:off_b0 + H], extra1=bk[off_b1:off_b1 + H],
                                            dy_row_base=base, max_sms=max_sms)
                if use_colsum:  # enqueued after the GEMM, on another stream: must run beside the GEMM's CTAs
                    with torch.cuda.stream(side[r]):
                        L.colsum(dy, bk[off_b0:off_b0 + H], bk[off_b1:off_b1 + H], cs_ws[r], alpha0=0.5, alpha1=0.25,
                                 dy_row_base=base, sum_rows=R, comm=comms[r])
            torch.cuda.synchronize()
            assert vr.status.tolist() == [0] * world, f"a virtual rank timed out waiting for its peers: {vr.status.tolist()}"
            got = [b.cpu() for b in vr.buckets]
            for r in range(1, world):
                for lo, hi in ((0, H * Ka), (offs[1], offs[1] + H * Kv), (off_b0, off_b0 + H), (off_b1, off_b1 + H)):
                    assert torch.equal(got[0][lo:hi], got[r][lo:hi]), f"rank {r} differs from rank 0 in [{lo}, {hi})"
            dwa = got[0][:H * Ka].view(H, Ka).double()
            assert torch.isfinite(dwa).all()
            # fp32 accumulation of the reduction over R*B rows, then `world` fp32 adds: 2e-5 of the largest entry
            assert float((dwa - ref_a).abs().max() / ref_a.abs().max()) <= 2e-5
            if Kv:
                dwv = got[0][offs[1]:offs[1] + H * Kv].view(H, Kv).double()
                assert float((dwv - ref_v).abs().max() / ref_v.abs().max()) <= 2e-5
            db = torch.stack([got[0][off_b0:off_b0 + H], got[0][off_b1:off_b1 + H]]).double()
            assert float((db - ref_b).abs().max()) <= (2e-5 * float(ref_b.abs().max()) if use_colsum else 1e-6 * world)
    finally:
        vr.free()


def test_fused_allreduce_single_rank_is_plain_dw(avc, cuda_dev):
    """world = 1: the comm warps only run the protocol against themselves; results equal the plain kernel's."""
    run_virtual(avc._lib, cuda_dev, 1, B=2, R=150, H=320, Ka=192, Kv=72, base=0)
    run_virtual(avc._lib, cuda_dev, 1, B=2, R=100, H=1024, Ka=2048, Kv=1024, base=16, epochs=3)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_fused_allreduce_virtual_ranks_small_grids(avc, cuda_dev, monkeypatch, world):
    """Few workers per rank (multi-round schedule + tail sub-tiles) so that all ranks are trivially co-resident."""
    monkeypatch.setenv("AVC_GEMM_MAX_WORKERS", "5")
    run_virtual(avc._lib, cuda_dev, world, B=2, R=130, H=1024, Ka=1536, Kv=520, base=8, epochs=3, seed=world)


def test_fused_allreduce_virtual_ranks_audio_only(avc, cuda_dev, monkeypatch):
    monkeypatch.setenv("AVC_GEMM_MAX_WORKERS", "3")
    run_virtual(avc._lib, cuda_dev, 2, B=1, R=257, H=640, Ka=1280, Kv=0, base=0, seed=11)


@pytest.mark.parametrize("world", [2, 4])
def test_fused_allreduce_virtual_ranks_share_the_gpu(avc, cuda_dev, world):
    """BASELINE cfg2 weight shapes (4096 x 4096 + 4096 x 2048), every virtual rank on 148 / world SMs; the bias
    gradients come from avc_colsum_comm running beside the GEMM CTAs (as in engine.ConnectorStep)."""
    run_virtual(avc._lib, cuda_dev, world, B=4, R=375, H=4096, Ka=4096, Kv=2048, base=16, epochs=2, seed=3,
                use_colsum=True)


def test_fused_allreduce_bias_sums_from_colsum_comm(avc, cuda_dev, monkeypatch):
    monkeypatch.setenv("AVC_GEMM_MAX_WORKERS", "5")
    run_virtual(avc._lib, cuda_dev, 2, B=3, R=130, H=1024, Ka=1536, Kv=520, base=8, epochs=3, seed=21,
                use_colsum=True)


def test_fused_allreduce_rejects_bad_descriptors(avc, cuda_dev):
    L = avc._lib
    vr = VirtualRanks(L, 1, 4096, cuda_dev)
    try:
        c = vr.descriptors()[0]
        dy = torch.zeros(1, 64, 64, dtype=torch.bfloat16, device=cuda_dev)
        x = torch.zeros(1, 64, 64, dtype=torch.bfloat16, device=cuda_dev)
        inside = vr.buckets[0][:4096].view(64, 64)
        outside = torch.zeros(64, 64, dtype=torch.float32, device=cuda_dev)  # a torch allocation, not the bucket
        with pytest.raises(L.ConnectorError, match="not inside the local bucket"):
            L.proj_bwd_dw_allreduce(dy, [x], [outside], [1.0], c)
        with pytest.raises(L.ConnectorError, match="extra range 0"):
            L.proj_bwd_dw_allreduce(dy, [x], [inside], [1.0], c, extra0=outside.view(-1)[:64])
        c.epoch = 0
        with pytest.raises(L.ConnectorError, match="epochs count from 1"):
            L.proj_bwd_dw_allreduce(dy, [x], [inside], [1.0], c)
        c.epoch, c.world = 1, 9
        with pytest.raises(L.ConnectorError, match="out of range"):
            L.proj_bwd_dw_allreduce(dy, [x], [inside], [1.0], c)
        torch.cuda.synchronize()
    finally:
        vr.free()
