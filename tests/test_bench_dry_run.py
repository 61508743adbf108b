"""bench.py's GPU arm, dry: run_ours() end to end on the CPU with stand-ins for the CUDA runtime, the collectives and the library
(no compute, unit timings) — the control flow, the input / output rings and the assembly of the contract's JSON line are what is
tested here; the numbers are meaningless.  The real thing runs under `-m gpu` (tests/test_bench_configs_gpu.py) and on the box."""
import contextlib
import io
import json
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import bench  # noqa: E402


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self):
        pass

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return 1.0


class _Graph:
    replays = 0

    def replay(self):
        _Graph.replays += 1


@contextlib.contextmanager
def _capture(graph):
    yield


class _TorchProxy:
    """torch with a CUDA runtime that does nothing and every device mapped to the CPU"""
    cuda = types.SimpleNamespace(is_available=lambda: True, set_device=lambda i: None, synchronize=lambda *a: None, Event=_Event,
                                 CUDAGraph=_Graph, graph=_capture, empty_cache=lambda: None)

    def __getattr__(self, name):
        return getattr(torch, name)

    def device(self, *a):
        return torch.device("cpu")


def dry_run(argv, monkeypatch, world=1):
    import torch.distributed as dist

    import diffusion_uncertainty_b200 as pkg

    launches = {"plans": 0, "inputs": set()}

    class Plan:
        def __init__(self, scores, eps, sample, *a, **k):
            self.res = {"prev": None, "thr": torch.zeros(eps.shape[0])}
            self.eps = eps
            launches["plans"] += 1

        def set_map_out(self, u):
            pass

        def set_prev_out(self, p):
            self.res["prev"] = p

        def launch(self):
            ops.launch_count += 1
            launches["inputs"].add(self.eps.data_ptr())
            return self.res

        def launch_with_batch_sum(self, eps, S):
            assert eps is self.eps, "the batch sum must read the input copy the step reads"
            ops.launch_count += 2
            launches["inputs"].add(self.eps.data_ptr())
            return self.res

    ops = types.SimpleNamespace(FusedStep=Plan, launch_count=0, make_coeffs=lambda *a, **k: None, fused_supported=lambda n, dt: 1,
                                fused_last_kernel=lambda: "fused_pred_kernel", batch_sum=lambda e, out: None)

    class HostStep:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Event()

    def all_gather(out, t):
        for o in out:
            o.copy_(t)

    synth = bench.synth_host
    monkeypatch.setattr(bench, "torch", _TorchProxy())
    monkeypatch.setattr(bench, "synth_host", lambda B, C, H, W, M, dt, seed, pin: synth(B, C, H, W, M, dt, seed, False))
    monkeypatch.setattr(bench.StepBench, "parity", lambda self: {"mask_agreement": 1.0, "thr_bit_exact": True})
    monkeypatch.setattr(bench.StepBench, "time_eager_reference", lambda self, reps=5: 1.0)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setitem(sys.modules, "diffusion_uncertainty_b200.ops", ops)
    monkeypatch.setattr(pkg, "ops", ops, raising=False)
    monkeypatch.setitem(sys.modules, "diffusion_uncertainty_b200.host_step", types.SimpleNamespace(HostStreamedUncertaintyStep=HostStep))
    monkeypatch.setattr(dist, "init_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "destroy_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
    monkeypatch.setattr(dist, "all_reduce", lambda t, op=None: None)
    monkeypatch.setattr(dist, "all_gather", all_gather)
    monkeypatch.setenv("WORLD_SIZE", str(world))
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    lines = buf.getvalue().strip().splitlines()
    assert len(lines) == 1, "rank 0 prints ONE line"
    return json.loads(lines[0]), launches


CONTRACT_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                 "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "parity_checked"}


def test_gpu_arm_line_single_gpu(monkeypatch):
    d, seen = dry_run(["--steps", "6", "--warmup", "3", "--no-loop", "--no-cpu"], monkeypatch)
    assert CONTRACT_KEYS <= set(d)
    assert d["metric"] == "uncertainty_step_throughput" and d["unit"] == "Mpix/s" and d["n_gpus"] == 1 and d["scaling"] == "weak"
    assert d["steps"] == 6 and d["warmup"] == 3 and d["dtype"] == "f32" and d["vs_baseline"] is None and d["parity_checked"] is True
    assert d["config"] == bench.step_config("imagenet128_adm_b128_m5", 128, 1, True, "f32", "weak")       # = the reference arm's
    assert d["config"]["l2"].startswith("no flush") and d["run"]["input_copies"] == 1 and seen["plans"] >= 1
    assert d["gpu_launches"] == 12                                  # du_batch_sum + the fused step, per step
    assert d["e2e"]["h2d_bytes_per_step"] == 128 * 3 * 128 * 128 * (6 * 4 + 4) and d["e2e"]["d2h_bytes_per_step"] == 128 * 3 * 128 * 128 * 8
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["kernel"] == "fused_pred_kernel" and r["algorithmic_bytes"] == 128 * 3 * 128 * 128 * 36
    assert r["traffic"] is not None and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert set(d["other_workloads"]) == {"imagenet128_adm_b128_m5:fp16", "imagenet64_adm_b128_m5:fp32", "cifar10_ddpm_b16_m5:fp32",
                                         "uvit256_latent_b128_m5:fp32", "sd512_latent_b1_m16:fp32"}
    assert all("error" not in v for v in d["other_workloads"].values()), d["other_workloads"]
    assert "gpu_eager_baseline" in d and "cpu_baseline" not in d


@pytest.mark.parametrize("world", [2, 8])
def test_gpu_arm_line_strong_split(monkeypatch, world):
    d, seen = dry_run(["--gpus", str(world), "--steps", "20", "--warmup", "5", "--no-loop"], monkeypatch, world=world)
    assert CONTRACT_KEYS <= set(d)
    assert d["n_gpus"] == world and d["scaling"] == "strong" and d["config"]["images_per_gpu"] == 128 // world
    assert d["config"] == bench.step_config("imagenet128_adm_b128_m5", 128, world, True, "f32", "strong")
    assert "no collective" in d["config"]["parallelism"] and d["config"]["global_batch"] == 128
    # a shard below the L2 size: the steps rotate over resident copies of the input set, each with its own prepared launch
    R = d["run"]["input_copies"]
    assert R == bench.input_ring(36 * 3 * 128 * 128 * (128 // world)) and R > 1 and len(seen["inputs"]) >= R
    assert d["config"]["l2"].startswith("no flush: the steps rotate over %d resident copies" % R)
    assert d["roofline"]["traffic"] is None              # the committed ncu capture is of the 128-image launch
    o = d["other_scaling"]
    assert o["scaling"] == "weak" and o["images_per_gpu"] == 128 and o["l2"].startswith("no flush: per-step working set")
    assert len(d["ms_per_step_by_rank"]) == world and "cpu_baseline" not in d and "other_workloads" not in d


def test_gpu_arm_line_with_the_all_reduce_inside_the_step(monkeypatch):
    """--allreduce-batch-sum: every rank must replay the captured step (which then holds a collective) the same number of times —
    the warm-up count is agreed on through a collective, not timed per rank"""
    d, _ = dry_run(["--gpus", "2", "--steps", "5", "--warmup", "3", "--no-loop", "--no-extras", "--allreduce-batch-sum"], monkeypatch, world=2)
    assert "all-reduce" in d["config"]["parallelism"] and d["scaling"] == "strong" and d["gpu_launches"] == 10
    assert "other_scaling" not in d
