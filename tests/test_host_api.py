"""CPU-only checks of the host layer: the C-ABI library loads and exports every symbol include/du_b200.h declares, and
the drop-in scheduler classes mirror the reference's constructor / config / set_timesteps / error behaviour
(SURVEY.md §8b).  No kernel is launched here."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ----------------------------------------------------------------------------------------------- C ABI
def header_symbols():
    src = open(os.path.join(ROOT, "include", "du_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(du_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from diffusion_uncertainty_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/du_b200.h but not exported by libdu_b200.so"
    # and the ctypes prototypes of the host layer cover exactly the header
    assert sorted(_lib.PROTOTYPES) == names
    loaded = _lib.load()
    assert loaded.du_version() == 1
    assert isinstance(loaded.du_last_error(), bytes)
    assert loaded.du_fused_supported(3 * 128 * 128, _lib.F32) >= 1
    assert loaded.du_fused_supported(3 * 512 * 512, _lib.F32) == 0
    assert loaded.du_quantile_scratch_bytes(4, 1000) >= 0


def test_struct_layouts_match_the_header():
    """sizes the C compiler gives the parameter structs == sizes of the ctypes mirrors"""
    import subprocess
    import tempfile
    from diffusion_uncertainty_b200 import _lib
    prog = ('#include "du_b200.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu\\n", sizeof(du_ddim_coeffs), '
            'sizeof(du_guided_params), sizeof(du_fused_params)); return 0;}\n')
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.DdimCoeffs), ctypes.sizeof(_lib.GuidedParams), ctypes.sizeof(_lib.FusedParams)]


def test_no_cpu_fallback():
    from diffusion_uncertainty_b200 import ops
    x = torch.randn(2, 3, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.moments([x, x], mode="var")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.quantile_threshold(x, 0.9)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.ddim_step(x, x, ops.make_coeffs(1.0, 0.0, 1.0, 0.0))


def test_in_kernel_noise_has_no_cpu_path():
    """the draw is fused only for dense CUDA tensors; CPU tensors never reach a kernel and never get a silent torch fallback"""
    from diffusion_uncertainty_b200 import ops
    x = torch.randn(2, 3, 8, 8)
    assert ops.randn_fusable(x) is False
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.perturb_randn(x, 0.9, 0.1)
    with pytest.raises(RuntimeError):
        ops.randn_like(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.perturb_fresh(x, 0.9, 0.1)           # torch.randn_like, then du_perturb refuses the CPU tensors
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.dpm_solver_update(x, x, None, 1.0, 1.0)


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm: oracle port on the host cores) — one JSON line with the contract's keys"""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--workload", "cifar10_ddpm_b16_m5"], capture_output=True, text=True, check=True, timeout=300).stdout
    d = json.loads(out.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "cifar10_ddpm_b16_m5" and d["steps"] == 1
    loop = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "imagenet128_adm_loop"],
                          capture_output=True, text=True, check=True, timeout=120).stdout
    assert json.loads(loop.strip().splitlines()[-1])["impl"] == "reference"


def test_bench_config_is_the_same_dictionary_in_both_arms():
    """the reference arm prints the GPU arm's `config` (workload, rank count and scaling mode decide it, nothing of the run does),
    and every working set below the L2 size is rotated over enough copies of the input set or labelled L2-resident"""
    import bench
    for world in (1, 2, 4, 8):
        scaling = "strong" if world > 1 else "weak"       # the defaults of both arms
        ours = bench.step_config("imagenet128_adm_b128_m5", 128, world, 1, "f32", scaling, False)
        ref = bench.step_config("imagenet128_adm_b128_m5", 128, world, True, "f32", scaling)
        assert ours == ref and ours["global_batch"] == 128 and ours["images_per_gpu"] == 128 // world
        assert ("no collective" in ours["parallelism"]) and ours["l2"].startswith("no flush")
    weak = bench.step_config("imagenet128_adm_b128_m5", 128, 4, 1, "f32", "weak")
    assert weak["global_batch"] == 512 and weak["images_per_gpu"] == 128
    assert "all-reduce" in bench.step_config("imagenet128_adm_b128_m5", 128, 2, 1, "f32", "strong", True)["parallelism"]
    for alg in (226.5e6, 113.2e6, 56.6e6, 28.3e6, 20e6):
        R = bench.input_ring(alg)
        assert R == 1 if alg > bench.L2_BYTES else (R - 1) * alg >= 2 * bench.L2_BYTES
    assert bench.input_ring(1.3e6) == bench.IN_RING_MAX and bench.l2_note(1.3e6).startswith("L2-resident")


def test_host_quantile_stand_in_of_the_gpu_tests():
    """tests/conftest.py routes torch.quantile of the GPU parity tests through this wrapper: uncontracted lerp whatever the host does"""
    import numpy as np
    import torch

    from oracle import du_oracle as O
    from tests.helpers import host_quantile_without_contraction
    real = torch.quantile
    hq = host_quantile_without_contraction(real)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(64, 3072, generator=g) ** 3
    for q in (0.0, 0.333, 0.9, 0.95, 1.0):
        want = O.quantile_linear_rows(x, q, lerp_fma=False)[0]
        got = hq(x, q, dim=1)
        assert got.shape == (64,) and np.array_equal(got.numpy().view(np.int32), want.numpy().view(np.int32))
        assert hq(x, q, dim=1, keepdim=True).shape == (64, 1) and torch.equal(hq(x, q, dim=-1), got)
        fused = O.quantile_linear_rows(x, q, lerp_fma=True)[0]
        host = real(x, q, dim=1)
        assert all(h in (a, b) for h, a, b in zip(host.numpy().view(np.int32), want.numpy().view(np.int32), fused.numpy().view(np.int32)))
    x[3, 7] = float("nan")
    got = hq(x, 0.5, dim=1)
    assert torch.isnan(got[3]) and not torch.isnan(got[2])
    # everything else is torch's own function, untouched
    assert torch.equal(hq(x[0], 0.5), real(x[0], 0.5)) and torch.equal(hq(x.double()[:2], 0.5, dim=1), real(x.double()[:2], 0.5, dim=1))
    assert torch.equal(hq(x[:2], torch.tensor([0.25, 0.5]), dim=1), real(x[:2], torch.tensor([0.25, 0.5]), dim=1))
    assert torch.equal(hq(x[:2], 0.5, dim=0), real(x[:2], 0.5, dim=0))
    with pytest.raises(RuntimeError):
        hq(x, 1.5, dim=1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "diffusion-uncertainty_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"


# ----------------------------------------------------------------------------------------------- schedulers
SU = "diffusion_uncertainty_b200.schedulers_uncertainty."
MODULES = ["scheduling_ddim_uncertainty_zigzag_centered", "scheduling_ddim_uncertainty_zigzag", "scheduling_ddim_uncertainty_centered",
           "scheduling_ddim_uncertainty_centered_d", "scheduling_ddim_uncertainty", "scheduling_ddim_uncertainty_image",
           "scheduling_ddim_infer_noise", "scheduling_ddim_mc_dropout", "scheduling_ddim_uncertainty_threshold",
           "scheduling_ddim_infer_noise_multiscale_threshold", "scheduling_ddim_flip", "scheduling_ddim_flip_threshold",
           "scheduling_ddim_uncertainty_grad", "scheduling_ddim_mc_dropout_gradient"]
CLASSES = ["DDIMSchedulerUncertainty", "DDIMSchedulerUncertaintyImagenet", "DDIMSchedulerUncertaintyCifar10",
           "DDIMSchedulerUncertaintyImagenetClassConditioned"]


def imp(name):
    import importlib
    return importlib.import_module(SU + name)


@pytest.mark.parametrize("module", MODULES)
def test_every_reference_module_and_class_name_exists(module):
    from diffusion_uncertainty_b200.schedulers_uncertainty.mixin import (SchedulerUncertaintyClassConditionedMixin,
                                                                         SchedulerUncertaintyMixin)
    m = imp(module)
    for cls_name in CLASSES:
        cls = getattr(m, cls_name)
        s = cls(M=3, after_step=2, num_steps_uc=2)
        assert s.config.M == 3 and s.M == 3 and len(s) == 1000 and s.order == 1 and s.init_noise_sigma == 1.0
        assert s.alphas_cumprod.shape == (1000,) and s.alphas_cumprod.device.type == "cpu" and s.alphas_cumprod.dtype == torch.float32
        s.set_timesteps(10)
        assert isinstance(s, SchedulerUncertaintyMixin)
    assert isinstance(getattr(m, CLASSES[3])(), SchedulerUncertaintyClassConditionedMixin)
    assert getattr(m, CLASSES[3]).class_conditioned is True
    assert hasattr(m, "DDIMSchedulerUncertaintyOutput")


def test_config_is_readable_assignable_and_from_config_drops_unknown_keys():
    m = imp("scheduling_ddim_uncertainty_zigzag_centered")
    base = m.DDIMSchedulerUncertainty(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, clip_sample=False,
                                      set_alpha_to_one=False, steps_offset=1)
    s = m.DDIMSchedulerUncertaintyImagenetClassConditioned.from_config(
        base.config, M=5, after_step=40, num_steps_uc=10, num_zigzag=3, unet="model", y="labels", eta=0.0, not_a_parameter=1)
    assert s.config.beta_schedule == "scaled_linear" and s.config.num_zigzag == 3 and s.num_zigzag == 3
    assert s.unet == "model" and s.prompt_embeds == "labels" and s.predict_next is True
    assert "not_a_parameter" not in s.config and "eta" not in s.config
    s.set_timesteps(50)
    assert (s.timestep_after_step, s.timestep_end_step) == (181, 1)     # steps_offset = 1
    s.config.after_step = 0          # callers mutate the config between runs (guided_gradient.py:64-66)
    s.config.num_steps_uc = 3
    s.set_timesteps(50)
    assert (s.timestep_after_step, s.timestep_end_step) == (981, 941)
    assert s.uncertainty_timesteps() == [981, 961, 941]
    # a plain dict and a namespace work as config sources too
    s2 = m.DDIMSchedulerUncertainty.from_config(dict(base.config), M=2)
    assert s2.config.clip_sample is False and s2.M == 2
    s3 = m.DDIMSchedulerUncertainty.from_config(types.SimpleNamespace(num_train_timesteps=500, junk=3))
    assert len(s3) == 500


def test_set_timesteps_matches_the_golden_schedules(golden_dir):
    m = imp("scheduling_ddim_uncertainty_zigzag_centered")
    g = np.load(os.path.join(golden_dir, "sched_zigzag_centered.npz"))
    s = m.DDIMSchedulerUncertaintyImagenetClassConditioned(M=5, after_step=40, num_steps_uc=10, num_zigzag=3)
    s.set_timesteps(50)
    assert np.array_equal(s.timesteps.numpy(), g["timesteps"]) and s.timesteps.dtype == torch.int64
    assert s.timestep_after_step == int(g["after"]) and s.timestep_end_step == int(g["end"])
    g = np.load(os.path.join(golden_dir, "sched_zigzag_centered_eta.npz"))
    s = m.DDIMSchedulerUncertaintyImagenetClassConditioned(M=3, after_step=5, num_steps_uc=4, num_zigzag=2,
                                                           beta_schedule="squaredcos_cap_v2", clip_sample=False)
    s.set_timesteps(20)
    assert np.array_equal(s.timesteps.numpy(), g["timesteps"])
    assert s.timestep_after_step == int(g["after"]) and s.timestep_end_step == int(g["end"])
    for spacing, first in (("linspace", 999), ("trailing", 999), ("leading", 980)):
        s = m.DDIMSchedulerUncertainty(timestep_spacing=spacing)
        s.set_timesteps(50)
        assert int(s.timesteps[0]) == first and len(s.timesteps) == 50


def test_schedule_scalars_match_the_oracle():
    from oracle import du_oracle as O
    m = imp("scheduling_ddim_uncertainty_centered")
    for sched in ("linear", "scaled_linear", "squaredcos_cap_v2"):
        s = m.DDIMSchedulerUncertainty(beta_schedule=sched)
        assert torch.equal(s.betas, O.make_betas(sched)) and torch.equal(s.alphas_cumprod, torch.cumprod(1 - O.make_betas(sched), 0))
        s.set_timesteps(50)
        for t, eta in ((980, 0.0), (500, 0.3), (0, 0.0), (0, 1.0)):
            c, host = s._step_scalars(t, eta, False)
            o = O.DDIMCoeffs(s.alphas_cumprod, s.final_alpha_cumprod, t, t - 20, eta)
            got = (c.sqrt_alpha_t, c.sqrt_beta_t, c.sqrt_alpha_prev, c.dir_coef, c.sigma)
            want = tuple(np.float32(float(v)) for v in (o.sqrt_alpha_t, o.sqrt_beta_t, o.sqrt_alpha_prev, o.dir_coef, o.sigma))
            assert tuple(np.float32(v) for v in got) == want


def test_error_behaviour_matches_the_reference():
    m = imp("scheduling_ddim_uncertainty_zigzag_centered")
    s = m.DDIMSchedulerUncertainty(after_step=2, num_steps_uc=2)
    x = torch.zeros(1, 3, 4, 4)
    with pytest.raises(ValueError, match="set_timesteps"):
        s.step(x, 10, x)
    with pytest.raises(ValueError, match="cannot be larger"):
        s.set_timesteps(2000)
    with pytest.raises(NotImplementedError):
        m.DDIMSchedulerUncertainty(beta_schedule="nope")
    with pytest.raises(IndexError):              # window outside the schedule: indexing self.timesteps fails, as in the reference
        m.DDIMSchedulerUncertainty(after_step=10).set_timesteps(10)
    bad = m.DDIMSchedulerUncertainty(timestep_spacing="nope")
    with pytest.raises(ValueError, match="not supported"):
        bad.set_timesteps(10)
    s.set_timesteps(10)
    s.config.prediction_type = "nope"
    with pytest.raises(ValueError, match="prediction_type"):
        s.step(x, 900, x)
    s.config.prediction_type = "epsilon"
    with pytest.raises(RuntimeError, match="no CPU fallback"):     # CPU tensors never reach a kernel
        s.step(x, 900, x)
    thr = imp("scheduling_ddim_uncertainty_threshold").DDIMSchedulerUncertainty(prediction_type="sample", after_step=2, num_steps_uc=2)
    thr.set_timesteps(10)
    with pytest.raises(AssertionError, match="prediction type epsilon"):
        thr.step(x, 900, x)


def test_output_type_access_patterns():
    from diffusion_uncertainty_b200.outputs import DDIMSchedulerUncertaintyOutput
    a, b = torch.zeros(1), torch.ones(1)
    o = DDIMSchedulerUncertaintyOutput(prev_sample=a, pred_original_sample=b)
    assert o.prev_sample is a and o["prev_sample"] is a and o[0] is a and o.to_tuple() == (a, b)
    assert o.uncertainty is None and "uncertainty" not in o
    o.uncertainty = b
    o.pred_epsilon = a
    o.something_else = 3
    assert list(o.keys()) == ["prev_sample", "pred_original_sample", "uncertainty", "pred_epsilon", "something_else"]
    assert o["uncertainty"] is b and o.something_else == 3 and len(o) == 5
    with pytest.raises(KeyError):
        o["score"]
    with pytest.raises(AttributeError):
        o.missing


def test_factory_keys():
    from diffusion_uncertainty_b200.schedulers_uncertainty import get_uncertainty_scheduler
    from diffusion_uncertainty_b200.schedulers_uncertainty.get_uncertainty_scheduler import (instatiate_uc_scheduler,
                                                                                            instatiate_uncertainty_scheduler)
    base = imp("scheduling_ddim_uncertainty_centered").DDIMSchedulerUncertainty()
    args = types.SimpleNamespace(M=4, start_step_uc=3, num_steps_uc=2, predict_next=True, eta=0.0, uncertainty_distance=7, num_zigzag=3)
    expect = {"uncertainty": "ActivationNoise", "uncertainty_original": "ActivationNoise", "uncertainty_image": "UncertaintyImage",
              "uncertainty_centered": "Centered", "uncertainty_centered_d": "CenteredD", "uncertainty_zigzag_centered": "ZigZagCentered",
              "mc_dropout": "MCDropout", "anything_else": "MCDropout"}
    for key, variant in expect.items():
        args.scheduler_type = key
        s = get_uncertainty_scheduler(args, "y", "unet", base)
        assert variant in [c.__name__ for c in type(s).__mro__], key
        assert s.M == 4 and s.config.after_step == 3 and s.config.num_steps_uc == 2 and s.unet == "unet" and s.prompt_embeds == "y"
        assert s.class_conditioned is True
    assert s.__class__.__name__ == "DDIMSchedulerUncertaintyImagenetClassConditioned"
    args.scheduler_type = "uncertainty_original"
    assert get_uncertainty_scheduler(args, "y", "unet", base).predict_next is False
    args.scheduler_type = "uncertainty_centered_d"
    assert get_uncertainty_scheduler(args, "y", "unet", base).uncertainty_distance == 7
    args.scheduler_type = "flip"
    fl = get_uncertainty_scheduler(args, "y", "unet", base)
    assert "Flip" in [c.__name__ for c in type(fl).__mro__] and fl.config.after_step == 3 and fl.prompt_embeds == "y"
    args.scheduler_type = "flip_grad"
    with pytest.raises(NotImplementedError):
        get_uncertainty_scheduler(args, "y", "unet", base)
    args.scheduler_type = "dpm_2_uncertainty_centered"
    dp = get_uncertainty_scheduler(args, "y", "unet", base)
    assert dp.__class__.__name__ == "KDPM2SchedulerUncertaintyImagenetClassConditioned" and dp.M == 4 and dp.unet == "unet"
    assert dp.config.after_step == 3 and dp.config.beta_schedule == "linear" and dp.config.timestep_spacing == "leading"
    assert dp.prompt_embeds is None and dp.class_conditioned is True      # y= is dropped, as in the reference (the loop assigns it)
    dp.set_timesteps(10)
    assert dp.timesteps.tolist() == [900, 810, 720, 630, 540, 450, 360, 270, 180, 90] and len(dp.sigmas) == 11
    assert dp.timestep_after_step == 630 and dp.timestep_end_step == 540 and float(dp.sigmas[-1]) == 0.0
    from diffusion_uncertainty_b200.schedulers_uncertainty.mixin import SchedulerUncertaintyClassConditionedMixin, SchedulerUncertaintyMixin
    assert isinstance(dp, SchedulerUncertaintyMixin) and isinstance(dp, SchedulerUncertaintyClassConditionedMixin)
    import diffusion_uncertainty_b200.schedulers_uncertainty.scheduling_dpm_2_uncertainty_centered as dpm
    for bad in ("dpmsolver", "sde-dpmsolver++"):
        with pytest.raises(NotImplementedError):
            dpm.KDPM2DiscreteSchedulerUncertainty(algorithm_type=bad)
    with pytest.raises(ValueError, match="set_timesteps"):
        dpm.KDPM2DiscreteSchedulerUncertainty().step(None, 0, None)
    assert instatiate_uc_scheduler is get_uncertainty_scheduler and instatiate_uncertainty_scheduler is get_uncertainty_scheduler


def test_model_dispatch_trait():
    from diffusion_uncertainty_b200.schedulers_uncertainty.traits import PredictorClassConditionedTrait

    class UViT(torch.nn.Module):
        def forward(self, x, t, y):
            return ("uvit", t.dtype, y)

    class UNet2DModel(torch.nn.Module):
        def forward(self, x, t):
            return types.SimpleNamespace(sample=("unet2d", t.shape))

    class ADM(torch.nn.Module):
        def forward(self, x, t, y=None):
            return torch.zeros(x.shape[0], 6, 2, 2)

    host = PredictorClassConditionedTrait()
    host.prompt_embeds = "emb"
    x = torch.zeros(5, 3, 2, 2)
    host.unet = UViT()
    assert host.predict_model(x, 7.4) == ("uvit", torch.int64, "emb")
    host.unet = UNet2DModel()
    assert host.predict_model(x, 3) == ("unet2d", (5,))
    host.unet = ADM()
    out = host.predict_model(x, 3)
    assert out.shape == (5, 3, 2, 2) and not out.is_contiguous()


def test_threshold_map_argument_checks():
    from diffusion_uncertainty_b200.pipeline_uncertainty import calculate_threshold_map
    u = torch.rand(2, 3, 4, 4)
    with pytest.raises(TypeError):
        calculate_threshold_map(1, None, u, "higher")          # int is neither Tensor nor float (beartype in the reference)
    with pytest.raises(TypeError):
        calculate_threshold_map(0.9, None, u, "sideways")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        calculate_threshold_map(0.9, None, u, "higher")
