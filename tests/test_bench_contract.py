"""bench.py's JSON line, produced by bench.main() with the GPU-facing pieces mocked (no GPU needed).

The timed numbers are fake; what is checked is that every branch of the line's assembly runs (N = 1 with e2e and the CPU
baseline, N > 1 with the fused / multicast / NCCL collective, strong scaling) and that the keys the driver reads exist.
"""
import contextlib
import io
import json
import sys
import types
from unittest import mock

import pytest
import torch


class FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, *a):
        pass

    def elapsed_time(self, other):
        return 100.0


class FakeEngine:
    fused, mc = False, None

    def __init__(self, shape, plan, dev, seed=0, fuse_gather=True, fused_allreduce=None, ragged_range=None):
        self.shape, self.direct, self.overlap_comm = shape, fuse_gather and ragged_range is None, False
        self.fused_allreduce = FakeEngine.fused and fused_allreduce is not False
        self.ragged = ragged_range is not None
        self.fused_tokens = self.M = shape.batch * 375
        self.S = 391
        self.K = 6144
        self.bias_in_gemm = True
        self.status = torch.zeros(1, dtype=torch.int32)
        peer = types.SimpleNamespace(mc=FakeEngine.mc, check=lambda: None, closed=False) if self.fused_allreduce else None
        if peer is not None:
            peer.close = lambda: setattr(peer, "closed", True)
        self.bucket = types.SimpleNamespace(peer=peer)
        self.launches_per_step = 5 if self.direct else 7
        self.events = None

    def step(self, allreduce=True):
        pass

    def enable_kernel_timing(self):
        names = ["proj_fwd", "splice_fwd", "proj_bwd_dw"] + ([] if self.direct else ["gather", "splice_bwd"])
        self.events = {n: [(FakeEvent(), FakeEvent())] for n in names}

    def gemm_flops(self):
        return 2 * self.M * 6144 * 4096

    def gather_bytes(self):
        return 4 * self.M * 6144

    def splice_bytes(self):
        return 4 * self.M * 4096


def fake_trainer_leg(torch_, dist_, eng, steps, world):
    # the leg all-reduces through the engine's own peer bucket: it has to run before anything closes that bucket
    assert eng.bucket.peer is None or not eng.bucket.peer.closed, "trainer leg after the peer bucket was closed"
    return {"ms_per_step": 120.0, "value": 1.5}


def run_bench(argv, world, fused, mc):
    import audio_visual_llm_b200.engine as E
    import bench
    import oracle.cpu_baseline  # noqa: F401  (patched below)

    FakeEngine.fused, FakeEngine.mc = fused, mc

    def all_gather(lst, t):
        for x in lst:
            x.copy_(t)

    env = {"RANK": "0", "WORLD_SIZE": str(world), "LOCAL_RANK": "0"}
    fake_e2e = {"value": 1.0, "unit": bench.UNIT, "h2d_bytes_per_step": 1, "d2h_bytes_per_step": 1}
    fake_check = {"ok": True, "output_max_rel": 1e-3, "dw_max_rel": 1e-5}
    fake_eager = {"ms_per_step": 200.0, "value": 1.0, "unit": bench.UNIT, "best": "two_linear",
                  "cublas": {"fwd_ms": 90.0, "dw_ms": 95.0}}
    with mock.patch.dict("os.environ", env), mock.patch.object(sys, "argv", ["bench.py"] + argv), \
            mock.patch.object(E, "ConnectorStep", FakeEngine), mock.patch("torch.cuda.set_device"), \
            mock.patch.object(bench, "self_check", lambda *a, **k: dict(fake_check)), \
            mock.patch.object(bench, "trainer_leg", fake_trainer_leg), \
            mock.patch.object(bench, "graphed_leg", lambda *a, **k: {"ms_per_step": 50.0, "value": 2.0}), \
            mock.patch.object(bench, "gpu_eager_leg", lambda *a, **k: dict(fake_eager, cublas=dict(fake_eager["cublas"]))), \
            mock.patch("torch.cuda.Event", FakeEvent), mock.patch("torch.cuda.synchronize"), \
            mock.patch("torch.cuda.empty_cache"), mock.patch("torch.device", lambda *a: "cpu"), \
            mock.patch.object(bench, "run_e2e", lambda *a, **k: dict(fake_e2e)), \
            mock.patch("torch.distributed.init_process_group"), mock.patch("torch.distributed.barrier"), \
            mock.patch("torch.distributed.all_reduce"), mock.patch("torch.distributed.all_gather", all_gather), \
            mock.patch("torch.distributed.destroy_process_group"), \
            mock.patch("oracle.cpu_baseline.time_cpu", lambda *a: (1000.0, 0.1, 8)):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            bench.main()
    lines = [ln for ln in buf.getvalue().splitlines() if ln.strip()]
    assert len(lines) == 1, "rank 0 prints exactly ONE line on stdout"
    return json.loads(lines[0])


BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches",
             "sustained", "self_check", "gpu_eager", "strong_scaling", "trainer_step"}


@pytest.mark.parametrize("argv,world,fused,mc,scaling,collective", [
    (["--steps", "5", "--warmup", "3"], 1, False, None, "weak", "none"),
    (["--steps", "5", "--warmup", "3", "--no-e2e"], 2, True, object(), "weak", "multimem.ld_reduce"),
    (["--steps", "5", "--warmup", "3", "--no-e2e"], 8, True, None, "weak", "peer loads + peer stores"),
    (["--steps", "5", "--warmup", "3", "--no-e2e", "--global-batch", "256"], 8, True, None, "strong", "fused into"),
    (["--steps", "5", "--warmup", "3", "--no-e2e"], 2, False, None, "weak", "NCCL"),
])
def test_bench_line(avc, argv, world, fused, mc, scaling, collective):
    d = run_bench(argv, world, fused, mc)
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["n_gpus"] == world and d["scaling"] == scaling and d["higher_is_better"] is True
    assert d["steps"] == 5 and d["warmup"] == 3 and d["dtype"] == "bf16" and d["vs_baseline"] is None
    assert collective in d["config"]["collective"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert d["roofline"]["bound"] == "tensor" and d["roofline"]["unit"] == "TFLOP/s"
    assert d["gpu_launches"] == 5 * 5
    assert d["self_check"]["ok"] is True and d["trainer_step"]["ms_per_step"] > 0
    assert {"steps", "ms_per_step", "value", "clocks", "roofline"} <= set(d["sustained"])
    assert d["gpu_eager"]["ours_over_eager_step"] > 0 and set(d["gpu_eager"]["ours_over_cublas"]) == {"fwd", "dw"}
    if world > 1 and scaling == "weak":
        assert d["strong_scaling"]["global_batch"] == 256 and d["strong_scaling"]["batch_per_gpu"] == 256 // world
    if world == 1:
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert d["config"]["global_batch"] == 32
    else:
        assert len(d["per_rank_kernel_ms"]["proj_bwd_dw_ms"]) == world
    if scaling == "strong":
        assert d["config"]["batch_per_gpu"] == 256 // world and d["config"]["global_batch"] == 256


@pytest.mark.parametrize("config", ["cfg1", "cfg3", "cfg4"])
def test_bench_line_other_configs(avc, config):
    d = run_bench(["--steps", "5", "--warmup", "3", "--config", config, "--no-cpu-baseline"], 1, False, None)
    assert BASE_KEYS <= set(d)
    assert d["config"]["workload"].startswith(config + ":")
    if config == "cfg4":
        assert "gather" in d["kernels"] and "splice_bwd" in d["kernels"], "ragged config: stand-alone kernels on the path"
