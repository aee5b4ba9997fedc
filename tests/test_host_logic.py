"""CPU: host-side logic of the drop-in layer (no kernels): token planning, resampling taps, sharding, and the
projector-gradient all-reduce over a 2-rank gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import connector_oracle as O


def test_fuse_plan_matches_oracle_token_rule(avc):
    for ka, kv, cap, Ta, Tv in [(1, 1, 256, 1500, 750), (4, 2, 1536, 1500, 750), (4, 2, 100, 1500, 750),
                                (3, 1, 64, 40, 7), (1, 1, 16, 24, 12)]:
        plan = avc.FusePlan(audio_stride=ka, video_stride=kv, max_seq_len=cap)
        spec = O.ConnectorSpec(audio_stride=ka, video_stride=kv, max_seq_len=cap)
        assert plan.tokens(Ta, Tv) == O.token_counts(spec, Ta, Tv)
        assert plan.tokens(Ta, None) == O.token_counts(spec, Ta, None) == -(-Ta // ka)  # single modality: no cap
    assert avc.FusePlan(fusion="sum", fusion_scale=0.3).scales(True, True) == (0.3, 0.7)
    assert avc.FusePlan(fusion="concat").scales(True, True) == (1.0, 1.0)
    assert avc.FusePlan(fusion="sum", fusion_scale=0.3).scales(True, False) == (1.0, 1.0)
    with pytest.raises(ValueError):
        avc.FusePlan(modality="text")


@pytest.mark.parametrize("S,L", [(28, 10), (1500, 256), (8, 19), (3, 5), (5, 1)])
def test_resample_taps_reproduce_adaptive_projection(avc, S, L):
    """The CSR taps fed to the row-resample kernel, applied densely on CPU, equal the oracle (and the transpose
    is the autograd backward)."""
    from audio_visual_llm_b200.seq_adapt import _taps

    (ptr, col, wt), (tptr, tcol, twt) = _taps(S, L)
    mat = torch.zeros(L, S, dtype=torch.float64)
    for i in range(L):
        for t in range(ptr[i], ptr[i + 1]):
            mat[i, col[t]] += wt[t]
    x = torch.randn(2, S, 6, dtype=torch.float64)
    ref = O.reference_adaptive_projection(x.float(), L).double()
    assert torch.allclose(mat @ x, ref, atol=1e-5)
    assert torch.allclose(mat.sum(1), torch.ones(L, dtype=torch.float64), atol=1e-6)
    tmat = torch.zeros(S, L, dtype=torch.float64)
    for s in range(S):
        for t in range(tptr[s], tptr[s + 1]):
            tmat[s, tcol[t]] += twt[t]
    assert torch.allclose(tmat, mat.t())


def test_shard_and_balance(avc):
    from audio_visual_llm_b200.parallel import balance_ragged, shard_batch

    spans = [shard_batch(256, r, 8) for r in range(8)]
    assert spans[0] == (0, 32) and spans[-1] == (224, 256)
    spans = [shard_batch(10, r, 4) for r in range(4)]
    assert [hi - lo for lo, hi in spans] == [3, 3, 2, 2] and spans[-1][1] == 10
    counts = [400, 100, 250, 399, 120, 130, 101, 380]
    parts = balance_ragged(counts, 2)
    assert sorted(sum(parts, [])) == list(range(8))
    loads = [sum(counts[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= min(counts)  # greedy longest-first: within one small item


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audio_visual_llm_b200.parallel import GradBucket

    shapes = {"audio_connector.linear.weight": (6, 8), "video_connector.linear.weight": (6, 4),
              "audio_connector.linear.bias": (6,), "video_connector.linear.bias": (6,)}
    b = GradBucket(shapes, "cpu")
    g = torch.Generator().manual_seed(rank)
    for v in b.views.values():
        v.copy_(torch.randn(v.shape, generator=g))
    local = {k: v.clone() for k, v in b.views.items()}
    b.allreduce()
    lin = torch.nn.Linear(8, 6)
    b.attach([("audio_connector.linear.weight", lin.weight), ("audio_connector.linear.bias", lin.bias)])
    assert lin.weight.grad.data_ptr() == b["audio_connector.linear.weight"].data_ptr()
    out[rank] = ({k: v.clone() for k, v in b.views.items()}, local)
    dist.destroy_process_group()


def test_grad_bucket_allreduce_two_ranks_gloo(avc):
    """N > 1 path: the flat projector-gradient bucket is averaged across ranks with one collective."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (r0, l0), (r1, l1) = out[0], out[1]
    for k in r0:
        assert torch.equal(r0[k], r1[k])
        assert torch.allclose(r0[k], (l0[k] + l1[k]) / 2)
    views = list(r0.values())
    assert all(v.dtype == torch.float32 for v in views)


def test_reference_arm_and_cpu_baseline_run(avc):
    """bench.py's CPU legs execute the oracle port on a tiny workload."""
    from oracle import cpu_baseline

    wl = dict(modality="both", fusion="concat", fusion_scale=0.5, max_seq_len=64, audio_stride=4, video_stride=2,
              audio_frames=40, video_frames=20, audio_dim=16, video_dim=8, hidden=32, prompt_len=4)
    tok_s, dt, threads = cpu_baseline.time_cpu(wl, 2, 1, 0)
    assert tok_s > 0 and threads >= 1
    emb, mask, lab, grads = cpu_baseline.one_step(cpu_baseline.make_case(wl, 2))
    assert emb.shape == (2, 14, 32) and mask.dtype == torch.int64 and grads[0].shape == (32, 64)


def test_config_keys_follow_the_reference_yaml(avc, tmp_path):
    """The shipped reference yaml layout (nested model:/data:) loads; new keys default to reference behaviour."""
    from audio_visual_llm_b200.config import load_config, model_kwargs

    ref_yaml = """
data:
  max_seq_len: 512
model:
  llm_path: "checkpoints/Llama-3.2-1B"
  whisper_model: "openai/whisper-medium"
  modality: "both"
  use_fp16: false
  freeze_encoders: true
  fusion_scale: 0.5
"""
    p = tmp_path / "clip_whisper.yaml"
    p.write_text(ref_yaml)
    kw = load_config(str(p))
    assert kw["max_seq_len"] == 512 and kw["modality"] == "both" and kw["fusion_scale"] == 0.5
    assert kw["llm_path"] == "checkpoints/Llama-3.2-1B" and kw["connector_type"] == "simple"
    assert (kw["fusion"], kw["stride"], kw["align"], kw["mask_mode"], kw["label_mode"]) == ("sum", 1, "index", 0, 0)
    kw = model_kwargs({"model": {"fusion": "concat", "stride": 4, "align": "rate"}, "modality": "audio"})
    assert kw["fusion"] == "concat" and kw["stride"] == 4 and kw["align"] == "rate" and kw["modality"] == "audio"
    with pytest.raises(ValueError):
        model_kwargs({"model": {"fusion": "attention"}})


def test_optimizer_decay_groups_follow_the_reference(avc):
    """'bias' in the parameter name -> no weight decay (clip_whisper_trainer.py:188)."""
    from audio_visual_llm_b200.trainer_step import ConnectorAdamW

    assert ConnectorAdamW.decay_for(type("X", (), {"weight_decay": 0.01})(), "audio_connector.linear.bias") == 0.0
    assert ConnectorAdamW.decay_for(type("X", (), {"weight_decay": 0.01})(), "video_connector.linear.weight") == 0.01


def test_fd_passing_between_threads(tmp_path, monkeypatch):
    """The multicast object's POSIX fd travels from rank 0 to the other ranks over an AF_UNIX socket (SCM_RIGHTS):
    `parallel._serve_fd` / `_fetch_fd`.  Checked here with an ordinary file: the received descriptor is a NEW number
    that refers to the same open file."""
    import os
    import secrets

    from audio_visual_llm_b200 import parallel

    tag = "test_" + secrets.token_hex(8)  # random per-job token, as _setup_multicast broadcasts it
    assert parallel._fd_socket_address(tag).startswith("\0"), "abstract namespace: no file in a shared /tmp"
    payload = tmp_path / "payload.bin"
    payload.write_bytes(b"multicast-object")
    fd = os.open(payload, os.O_RDONLY)
    try:
        server = parallel._serve_fd(fd, 2, tag)
        got = [parallel._fetch_fd(tag, timeout_s=5.0) for _ in range(2)]
        server.join(timeout=5)
        assert not server.is_alive()
        for g in got:
            assert g != fd
            assert os.pread(g, 64, 0) == b"multicast-object"
            os.close(g)
        # a second server on the same name fails loudly (rank 0 turns that into an all-rank fallback) ...
        s2 = parallel._serve_fd(fd, 1, tag + "x")
        with pytest.raises(OSError):
            parallel._serve_fd(fd, 1, tag + "x")
        os.close(parallel._fetch_fd(tag + "x", timeout_s=5.0))
        s2.join(timeout=5)
        # ... and a client without a server times out instead of hanging
        with pytest.raises(TimeoutError):
            parallel._fetch_fd(tag + "_nobody", timeout_s=0.2)
    finally:
        os.close(fd)


def test_numa_helpers_topology_and_binding(avc):
    """numa.py: sysfs parsing, and -- where the kernel allows mbind -- that the pages really land on the node."""
    import ctypes

    from audio_visual_llm_b200 import numa

    assert numa._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert numa._parse_cpulist("") == set()
    ns = numa.nodes()
    assert ns == sorted(ns)
    if not ns:
        pytest.skip("no NUMA nodes in sysfs")
    assert all(len(numa.node_cpus(n)) > 0 for n in ns)
    t = torch.empty(1 << 22, dtype=torch.uint8)
    err = numa._mbind(t.data_ptr(), t.nbytes, ns[0])
    ctypes.memset(t.data_ptr(), 0, t.nbytes)
    placed = numa.pages_on_node(t)
    if err == 0 and placed:
        assert placed.get(ns[0], 0) >= (1 << 22) // 4096 and sum(placed.values()) == placed[ns[0]]
    with numa.cpus_of_node(ns[0]) as n:
        assert n >= 0
