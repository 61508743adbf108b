"""The drop-in pipeline classes and threshold-guided loops on the GPU against fixtures recorded from the UNMODIFIED reference
(tests/golden/make_golden.py `pipelines`): DiffusionClassConditionalWithUncertainty, DiffusionClassConditionalGuidedPosteriorDistribution
(percentile -> the FUSED single-launch step; tensor threshold -> the kernel chain), DiffusionClassConditionalGuidedSecondOrder and
the three `..._with_threshold / _with_percentile` loops.

Bars: x_{t-1} of the last batch within 1e-5 relative of max(|x|, 0.4) (the DDIM update cancels, `x - sqrt(1-abar) eps`, so a pure
relative bound cannot hold near zero crossings; 4e-6 absolute is the bound for operands of magnitude <= 4); a pixel whose map value
sits within an ulp of its threshold may take the other branch, which is allowed for < 0.5 % of the pixels; uint8 images may differ
by one level where the float value rounds at .5.
"""
import numpy as np
import pytest
import torch

from tests.test_oracle_golden import T, load
from tests.toy_models import ToyADM, ToyADMWithParameter, UViTAE, seeded_noise

pytestmark = pytest.mark.gpu

BASE = dict(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", clip_sample=True, set_alpha_to_one=True,
            steps_offset=0, prediction_type="epsilon", timestep_spacing="leading")
UVIT_CFG = dict(BASE, beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, clip_sample=False, set_alpha_to_one=False,
                steps_offset=1)


def dev():
    return torch.device("cuda:0")


class StepRecorder:
    def __init__(self, sched):
        self.inner, self.last = sched.step, None
        sched.step = self

    def __call__(self, *a, **kw):
        out = self.inner(*a, **kw)
        self.last = out.prev_sample.detach().clone()
        return out


def close_frac(got, want, rtol=1e-5, floor=0.4):
    got, want = torch.as_tensor(got).double().cpu(), torch.as_tensor(want).double().cpu()
    assert got.shape == want.shape
    bad = (got - want).abs() > rtol * want.abs().clamp_min(floor)
    bad &= ~(torch.isnan(got) & torch.isnan(want))
    return float(bad.float().mean())


def images_close(got, want, max_frac=0.01):
    d = (torch.as_tensor(got).int() - torch.as_tensor(want).int()).abs()
    assert int(d.max()) <= 1 or float((d > 1).float().mean()) < 0.005, f"images differ by up to {int(d.max())} levels"
    assert float((d > 0).float().mean()) < max_frac


def plain_ddim(n_steps, cfg=BASE):
    from diffusion_uncertainty_b200.schedulers_uncertainty.scheduling_ddim import DDIMScheduler
    s = DDIMScheduler.from_config(cfg)
    s.set_timesteps(n_steps)
    return s


def test_pipeline_with_uncertainty_matches_reference(golden_dir):
    from diffusion_uncertainty_b200.pipeline_uncertainty import DiffusionClassConditionalWithUncertainty
    from diffusion_uncertainty_b200.schedulers_uncertainty.scheduling_ddim_uncertainty_zigzag_centered import \
        DDIMSchedulerUncertaintyImagenetClassConditioned as Sched
    g = load(golden_dir, "pipe_with_uncertainty")
    model = ToyADM(3, seed=50).eval().to(dev())
    sched = Sched.from_config(BASE, unet=model, M=3, after_step=12, num_steps_uc=4, num_zigzag=2)
    sched.set_timesteps(20)
    pipe = DiffusionClassConditionalWithUncertainty(model, sched, 16, dev(), 4, 0)
    with seeded_noise(80):
        res = pipe(X_T=T(g["x_T"]), y=T(g["y"]))
    assert res["uncertainty"].shape == g["uncertainty"].shape and res["score"].shape == g["score"].shape
    assert np.array_equal(res["score"].numpy(), g["score"]), "scores must replay bit for bit"
    assert np.array_equal(res["gen_images"].numpy(), g["gen_images"])
    u, w = res["uncertainty"].double(), T(g["uncertainty"]).double()
    assert float(((u - w).abs() / w.abs().clamp_min(1e-30)).max()) < 1e-5
    assert np.array_equal(res["y"].numpy(), g["y"]) and np.array_equal(res["x_t"].numpy(), g["x_T"])
    # the `sample` dispatch of the reference class (:65-75)
    assert pipe.sample.__name__ == "sample"


@pytest.mark.parametrize("tag", ["q", "t"])
def test_posterior_pipeline_matches_reference(golden_dir, tag):
    from diffusion_uncertainty_b200 import ops
    from diffusion_uncertainty_b200.pipeline_uncertainty import DiffusionClassConditionalGuidedPosteriorDistribution as Pipe
    g = load(golden_dir, f"pipe_posterior_{tag}")
    model = ToyADM(3, seed=51).eval().to(dev())
    thr = float(g["q"]) if tag == "q" else T(g["threshold"]).to(dev())
    sched = plain_ddim(8)
    pipe = Pipe(model, sched, thr, 16, dev(), 3, 0, M=4)
    with seeded_noise(81):
        rec = StepRecorder(sched)
        res = pipe(X_T=T(g["x_T"]), y=T(g["y"]), start_step=2, num_steps=3)
    if tag == "q":
        # 2 batches x window steps 2..5: every one of them ran as ONE fused launch behind the reference's class API
        assert pipe.fused_steps == 8, pipe.fused_steps
        assert ops.fused_last_kernel() in ("fused_step_kernel", "fused_pred_kernel")
    else:
        assert pipe.fused_steps == 0
    assert close_frac(rec.last, g["final_last_batch"]) < 0.005
    images_close(res["gen_images"], g["gen_images"])


def test_second_order_pipeline_matches_reference(golden_dir):
    from diffusion_uncertainty_b200.pipeline_uncertainty import DiffusionClassConditionalGuidedSecondOrder as Pipe
    g = load(golden_dir, "pipe_second_order")
    model = ToyADM(3, seed=52).eval().to(dev())
    sched = plain_ddim(8)
    pipe = Pipe(model, sched, float(g["q"]), 16, dev(), 3, 0, M=4, threshold_type="higher")
    with seeded_noise(82):
        rec = StepRecorder(sched)
        res = pipe(X_T=T(g["x_T"]), y=T(g["y"]), start_step=2, num_steps=4)
    assert close_frac(rec.last, g["final_last_batch"]) < 0.005
    images_close(res["gen_images"], g["gen_images"])


def test_second_order_momentum_kernel():
    """du_ema_update against the reference expressions (…guided_second_order.py:212-218) in torch on the same device."""
    from diffusion_uncertainty_b200.pipeline_uncertainty.pipeline_sampler_class_conditional_uncertainty_guided_second_order import \
        second_order_momentum_update
    gen = torch.Generator().manual_seed(3)
    u = (torch.rand(3, 3, 16, 16, generator=gen) ** 2).to(dev())
    mom = torch.rand(3, 3, 16, 16, generator=gen).to(dev())
    beta, i = 0.99, 7
    new, corrected, root = second_order_momentum_update(mom, u, i, beta)
    want = beta * mom + (1 - beta) * u
    assert torch.equal(new, want)
    # (torch divides a tensor by a Python scalar as a multiplication by its reciprocal; the kernel divides: last-bit differences)
    wc = want.double().cpu() / (1 - beta ** i + 1e-5)
    assert float(((corrected.double().cpu() - wc).abs() / wc.abs().clamp_min(1e-30)).max()) < 2e-7
    assert float(((root.double().cpu() - wc.sqrt()).abs() / wc.sqrt().clamp_min(1e-30)).max()) < 2e-7
    first, _, _ = second_order_momentum_update(None, u, 0, beta)
    assert torch.equal(first, u)


def test_threshold_loop_adm_matches_reference(golden_dir):
    from diffusion_uncertainty_b200.pipeline_uncertainty.uncertainty_guidance import \
        generate_samples_model_scheduler_class_conditioned_with_threshold as loop
    from diffusion_uncertainty_b200.schedulers_uncertainty.scheduling_ddim_uncertainty_centered import \
        DDIMSchedulerUncertaintyImagenetClassConditioned as Sched
    g = load(golden_dir, "loop_threshold_adm")
    model = ToyADMWithParameter(3, seed=53, scale=3.0).eval().to(dev())
    sched = Sched.from_config(BASE, unet=model, M=2, after_step=2, num_steps_uc=3)
    sched.set_timesteps(8)
    with seeded_noise(83):
        rec = StepRecorder(sched)
        res = loop(5, 3, 16, model, sched, 10, T(g["threshold"]).to(dev()), device=dev(), x_T=T(g["x_T"]), y=T(g["y"]), start_step=2,
                   num_steps=3)
    # downstream of a gradient through the model: 1e-4 (the reduction's backward has torch.var's formula, not its rounding order)
    assert close_frac(rec.last, g["final_last_batch"], rtol=1e-4) < 0.005
    images_close(res["gen_images"], g["gen_images"])


def test_threshold_loop_uvit_matches_reference(golden_dir):
    """BASELINE config 4's loop: U-ViT latent, the scheduler's own map against a fitted tensor threshold."""
    from diffusion_uncertainty_b200.generate_samples import generate_samples_uvit_scheduler_class_conditioned_with_threshold as loop
    from diffusion_uncertainty_b200.schedulers_uncertainty.scheduling_ddim_uncertainty_zigzag_centered import \
        DDIMSchedulerUncertaintyImagenetClassConditioned as Sched
    g = load(golden_dir, "loop_threshold_uvit")
    model = UViTAE(54).eval().to(dev())
    sched = Sched.from_config(UVIT_CFG, unet=model, M=2, after_step=2, num_steps_uc=3, num_zigzag=2)
    sched.set_timesteps(8)
    with seeded_noise(84):
        rec = StepRecorder(sched)
        res = loop(5, 3, 8, model, sched, 10, T(g["threshold"]).to(dev()), device=dev(), x_T=T(g["x_T"]), y=T(g["y"]), start_step=2,
                   num_steps=3)
    assert close_frac(rec.last, g["final_last_batch"], rtol=1e-4) < 0.005
    images_close(res["gen_images"], g["gen_images"])


def test_percentile_loop_adm_matches_reference(golden_dir):
    from diffusion_uncertainty_b200.generate_samples import generate_samples_model_scheduler_class_conditioned_with_percentile as loop
    from diffusion_uncertainty_b200.schedulers_uncertainty.scheduling_ddim_uncertainty_centered import \
        DDIMSchedulerUncertaintyImagenetClassConditioned as Sched
    g = load(golden_dir, "loop_percentile_adm")
    model = ToyADMWithParameter(3, seed=55, scale=3.0).eval().to(dev())
    sched = Sched.from_config(BASE, unet=model, M=2, after_step=2, num_steps_uc=3)
    sched.set_timesteps(8)
    with seeded_noise(85):
        rec = StepRecorder(sched)
        res = loop(4, 4, 16, model, sched, T(g["y"]).to(dev()), float(g["q"]), device=dev(), x_T=T(g["x_T"]), start_step=2, num_steps=3)
    assert close_frac(rec.last, g["final_last_batch"], rtol=1e-4) < 0.005
    images_close(res["gen_images"], g["gen_images"])


def test_guidance_function_posterior_is_one_fused_launch():
    """get_uncertainty_guided_score_with_percentile (posterior mode) on the SD latent shape: everything after the M forwards is
    a single du_fused_uncertainty_step launch (skip_ddim), and it equals the three-kernel chain on the same predictions."""
    import diffusion_uncertainty_b200.uncertainty_guidance as ug
    from diffusion_uncertainty_b200 import ops
    from tests.toy_models import ToySDUNet
    sd = ToySDUNet(4, seed=12).eval().to(dev())
    gen = torch.Generator().manual_seed(12)
    lat = torch.randn(1, 4, 64, 64, generator=gen).to(dev())
    lat2 = torch.cat([lat] * 2)
    emb = torch.randn(2, 8, 16, generator=gen).to(dev())
    t_tensor = torch.tensor(501, device=dev())
    a_hat = torch.cumprod(1 - torch.linspace(1e-4, 0.02, 1000), 0)[501]
    un, tx = sd(lat2, t_tensor, emb)[0].chunk(2)
    eps = (un + 7.5 * (tx - un)).detach().clone()
    torch.manual_seed(5)
    n0 = ops.launch_count
    out = ug.get_uncertainty_guided_score_with_percentile(eps, lat2.clone(), t_tensor, emb.clone(), sd, a_hat, 0.9, "stable-diffusion",
                                                          num_uncertainty_samples=16, guidance_scale=7.5)
    launches = ops.launch_count - n0
    assert ops.last_step_path == "fused" and ops.fused_last_kernel() == "fused_step_kernel"
    assert launches == 1 + 16 + 1, launches          # x0, 16 perturbations (noise drawn in the kernel), ONE fused step
    # the same predictions through the unfused chain
    torch.manual_seed(5)
    sa, sb = float(torch.sqrt(a_hat)), float(torch.sqrt(1 - a_hat))
    x0 = ops.ddim_step(eps.expand(lat2.shape), lat2, ops.make_coeffs(sa, sb, 0.0, 0.0, clip_sample=False), want_prev=False, want_x0=True)[1]
    preds = []
    for _ in range(16):
        o = sd(ops.perturb_fresh(x0, sa, sb, noise_like=eps), t_tensor, emb)[0]
        u_, t_ = o.chunk(2)
        preds.append(u_ + 7.5 * (t_ - u_))
    ref = ops.uncertainty_step(preds, eps, None, 0.9, None, a_hat, fused=False, want_mask=True)
    fus = ops.uncertainty_step(preds, eps, None, 0.9, None, a_hat, fused=True, want_mask=True)
    assert close_frac(fus["u"], ref["u"], rtol=2e-6, floor=0.0) == 0.0
    # threshold and mask are exact functions of the kernel's own map
    assert np.array_equal(fus["thr"].cpu().numpy(), torch.quantile(fus["u"].cpu().flatten(1), 0.9, dim=1).numpy())
    assert torch.equal(fus["mask"], (fus["u"] > fus["thr"].view(-1, 1, 1, 1)).float())
    assert close_frac(fus["eps"], ref["eps"]) < 0.002          # (a pixel on the threshold may flip with the map's last bit)
    assert torch.equal(out.view(torch.int32), fus["eps"].view(torch.int32))     # (bit patterns: u = 0 gives NaN, as in the reference)
