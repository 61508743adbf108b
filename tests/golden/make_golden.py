"""Generate the golden fixtures under tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, read-only) in the build container.

    python tests/golden/make_golden.py

The reference needs `diffusers` base classes only; tests/golden/_standin provides
them.  All noise is routed through a seeded CPU generator (tests/toy_models.seeded_noise) and the toy
score models are bit-reproducible across devices, so the fixtures replay exactly on the GPU box, where
/root/reference does not exist.  Fixtures are small (< 1.5 MB total) and committed.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "_standin"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from tests.helpers import l4_sampling_loop  # noqa: E402
from tests.toy_models import ToyADM, ToyADMWithParameter, ToySDUNet, seeded_noise  # noqa: E402

SU = "diffusion_uncertainty.schedulers_uncertainty."


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def base_config(**over):
    cfg = dict(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear",
               clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
               timestep_spacing="leading")
    cfg.update(over)
    return cfg


def run_scheduler_case(name, module, cls_name, ctor_kw, n_steps, B=4, C=3, H=16, seed=0, eta=0.0,
                       dropout=False, cfg_over=None, model_scale=1.0):
    import importlib
    mod = importlib.import_module(SU + module)
    cls = getattr(mod, cls_name)
    model = ToyADM(C, seed=seed, dropout=dropout, scale=model_scale).eval()
    g = torch.Generator().manual_seed(100 + seed)
    x_T = torch.randn(B, C, H, H, generator=g)
    y = torch.randint(0, 10, (B,), generator=g)
    with quiet():
        sched = cls.from_config(base_config(**(cfg_over or {})), unet=model, **ctor_kw)
        sched.set_timesteps(n_steps)
        with seeded_noise(1000 + seed):
            res = l4_sampling_loop(sched, model, x_T, y, eta=eta)
    keep = [0, len(res["prevs"]) // 2, len(res["prevs"]) - 1]
    save(name, x_T=x_T, y=y, final=res["final"], uncertainty=res["uncertainty"], score=res["score"],
         prev_first=res["prevs"][keep[0]], prev_mid=res["prevs"][keep[1]],
         timesteps=sched.timesteps, after=sched.timestep_after_step, end=sched.timestep_end_step)


def main_widen():
    """Fixtures of the rows added after the core path: flip / flip_threshold, the gradient schedulers and the
    per-pixel threshold fitting.  `python tests/golden/make_golden.py widen` writes only these."""
    torch.manual_seed(0)
    torch.set_num_threads(1)
    run_scheduler_case("sched_flip", "scheduling_ddim_flip", "DDIMSchedulerUncertaintyImagenet",
                       dict(after_step=10, num_steps_uc=5), n_steps=20, seed=10)
    run_scheduler_case("sched_flip_threshold", "scheduling_ddim_flip_threshold",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(after_step=10, num_steps_uc=5, uncertainty_threshold=0.5, uncertainty_threshold_mode="max"),
                       n_steps=20, seed=11)
    run_scheduler_case("sched_uncertainty_grad", "scheduling_ddim_uncertainty_grad",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=4, after_step=10, num_steps_uc=5, predict_next=False), n_steps=20, seed=12)
    run_scheduler_case("sched_mc_dropout_gradient", "scheduling_ddim_mc_dropout_gradient",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=4, after_step=10, num_steps_uc=5), n_steps=20, seed=13, dropout=True)
    # N2: the expressions of scripts/compute_threshold_pixel_wise.py:89-100 (the script has no importable function for them)
    g = torch.Generator().manual_seed(21)
    unc = torch.rand(37, 3, 3, 8, 8, generator=g) ** 3
    unc[5, 1, 0, 0, 0] = float("nan")
    out = {}
    for perc in (0.15, 0.9):
        thr = []
        for i in range(unc.shape[1]):
            uncertainties_timestep = unc[:, i]
            i_uncertaintities = uncertainties_timestep.argsort(dim=0)
            i_perc_th = i_uncertaintities[int(unc.shape[0] * perc)].unsqueeze(0)
            thr.append(uncertainties_timestep.gather(dim=0, index=i_perc_th).squeeze(0))
        out[f"thr_{perc}"] = torch.stack(thr, dim=0)
    save("pixel_thresholds", unc=unc, **out)

    # F6: DiffusionClassConditionalGuidedGradient.estimate_score_update (both gradient targets) + the blend lines :114-118
    from diffusion_uncertainty.pipeline_uncertainty import pipeline_sampler_class_conditional_uncertainty_guided_gradient as gg
    from diffusion_uncertainty.pipeline_uncertainty import \
        pipeline_sampler_class_conditional_uncertainty_guided_posterior_distribution as pd
    model = ToyADMWithParameter(3, seed=14).eval()
    g = torch.Generator().manual_seed(14)
    B = 4
    x = torch.randn(B, 3, 16, 16, generator=g)
    y = torch.randint(0, 10, (B,), generator=g)
    t_tensor = torch.full((B,), 300, dtype=torch.long)
    a_hat = torch.cumprod(1 - torch.linspace(1e-4, 0.02, 1000), 0)[30]
    out = {}
    for wrt in ("input", "score"):
        pipe = gg.DiffusionClassConditionalGuidedGradient(model, None, 0.9, 16, torch.device("cpu"), B, 0, M=4, gradient_wrt=wrt,
                                                          lambda_update=0.1)
        with torch.no_grad():
            eps = model(x, t_tensor, y=y)[:, :3].clone()
        with seeded_noise(14), quiet():
            u, upd = pipe.estimate_score_update(x.clone(), y, 7, t_tensor, eps, x.clone(), a_hat)
        m = pd.calculate_threshold_map(0.9, None, u.detach(), "higher").float()
        with torch.no_grad():
            post = eps + 0.1 * upd
            new = eps * (1 - m) + post * m
        out.update({f"{wrt}_u": u.detach(), f"{wrt}_update": upd.detach(), f"{wrt}_mask": m, f"{wrt}_eps_new": new.detach()})
    save("gradient_update", x=x, y=y, a_hat=a_hat, t=300, M=4, **out)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)

    # ---- A: the BASELINE scheduler (uncertainty_zigzag_centered), through the reference's own L4 loop too
    run_scheduler_case("sched_zigzag_centered", "scheduling_ddim_uncertainty_zigzag_centered",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=5, after_step=40, num_steps_uc=10, num_zigzag=3), n_steps=50)
    run_scheduler_case("sched_zigzag_centered_eta", "scheduling_ddim_uncertainty_zigzag_centered",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=3, after_step=5, num_steps_uc=4, num_zigzag=2), n_steps=20, seed=1, eta=0.3,
                       cfg_over=dict(beta_schedule="squaredcos_cap_v2", clip_sample=False))
    import diffusion_uncertainty.generate_samples as gs
    mod = __import__(SU + "scheduling_ddim_uncertainty_zigzag_centered", fromlist=["x"])
    model = ToyADM(3, seed=0).eval()
    g = torch.Generator().manual_seed(100)
    x_T = torch.randn(6, 3, 16, 16, generator=g)
    y = torch.randint(0, 10, (6,), generator=g)
    with quiet():
        sched = mod.DDIMSchedulerUncertaintyImagenetClassConditioned.from_config(
            base_config(), unet=model, M=5, after_step=40, num_steps_uc=10, num_zigzag=3)
        sched.set_timesteps(50)
        with seeded_noise(77):
            res = gs.generate_samples_model_scheduler_class_conditioned_from_tensor(
                X_T=x_T, y=y, batch_size=4, device=torch.device("cpu"), model=model, scheduler=sched)
    save("l4_zigzag_centered", x_T=x_T, y=y, gen_images=res["gen_images"], uncertainty=res["uncertainty"],
         score=res["score"])

    # ---- other scheduler variants
    run_scheduler_case("sched_zigzag", "scheduling_ddim_uncertainty_zigzag",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=4, after_step=10, num_steps_uc=5, num_zigzag=2), n_steps=20, seed=2)
    run_scheduler_case("sched_centered", "scheduling_ddim_uncertainty_centered",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=5, after_step=10, num_steps_uc=5, predict_next=False), n_steps=20, seed=3)
    run_scheduler_case("sched_centered_next", "scheduling_ddim_uncertainty_centered",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=5, after_step=10, num_steps_uc=5, predict_next=True), n_steps=20, seed=4)
    run_scheduler_case("sched_infer_noise", "scheduling_ddim_infer_noise",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=5, after_step=10, num_steps_uc=5, predict_next=False), n_steps=20, seed=5)
    run_scheduler_case("sched_mc_dropout", "scheduling_ddim_mc_dropout",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=6, after_step=10, num_steps_uc=5), n_steps=20, seed=6, dropout=True)
    run_scheduler_case("sched_threshold_max", "scheduling_ddim_uncertainty_threshold",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=5, after_step=10, num_steps_uc=5, uncertainty_threshold=1.0,
                            uncertainty_threshold_mode="max"), n_steps=20, seed=7)
    run_scheduler_case("sched_threshold_min", "scheduling_ddim_uncertainty_threshold",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=5, after_step=10, num_steps_uc=5, uncertainty_threshold=0.5,
                            uncertainty_threshold_mode="min", predict_next=True), n_steps=20, seed=8)
    run_scheduler_case("sched_multiscale", "scheduling_ddim_infer_noise_multiscale_threshold",
                       "DDIMSchedulerUncertaintyImagenetClassConditioned",
                       dict(M=5, after_step=10, num_steps_uc=5), n_steps=20, seed=9)

    # ---- F2a / F2b: calculate_threshold_map
    from diffusion_uncertainty.pipeline_uncertainty import \
        pipeline_sampler_class_conditional_uncertainty_guided_posterior_distribution as pd
    g = torch.Generator().manual_seed(5)
    arrays = {}
    cases = [("a", (4, 3, 16, 16), 0.9, "higher"), ("b", (4, 3, 16, 16), 0.95, "lower"),
             ("c", (2, 4, 32, 32), 0.9, "higher"), ("d", (3, 100, 7), 0.5, "higher"),
             ("e", (5, 7), 0.99, "higher"), ("f", (2, 3, 64, 64), 0.9, "higher")]
    for tag, shape, q, kind in cases:
        u = torch.rand(shape, generator=g) ** 3
        if tag == "c":  # heavy ties: quantise
            u = (u * 16).round() / 16
        if tag == "f":  # a NaN row
            u[1, 0, 0, 0] = float("nan")
        arrays[f"{tag}_u"] = u
        arrays[f"{tag}_q"] = q
        arrays[f"{tag}_higher"] = kind == "higher"
        arrays[f"{tag}_mask"] = pd.calculate_threshold_map(float(q), None, u, kind)
        arrays[f"{tag}_thr"] = torch.quantile(u.flatten(1).float(), q, dim=1)
    thr_map = torch.rand(6, 3, 16, 16, generator=g) ** 3
    u = torch.rand(4, 3, 16, 16, generator=g) ** 3
    arrays["t_u"], arrays["t_thr"] = u, thr_map
    arrays["t_mask_hi"] = pd.calculate_threshold_map(thr_map, 2, u, "higher")
    arrays["t_mask_lo"] = pd.calculate_threshold_map(thr_map, 3, u, "lower")
    save("threshold_map", **arrays)

    # ---- F1c + F2a + F5: estimate_score_update_posterior (batch-sum quirk at B=4)
    model = ToyADM(3, seed=11).eval()
    g = torch.Generator().manual_seed(11)
    B = 4
    x = torch.randn(B, 3, 16, 16, generator=g)
    y = torch.randint(0, 10, (B,), generator=g)
    t_tensor = torch.full((B,), 300, dtype=torch.long)
    alphas_cumprod = torch.cumprod(1 - torch.linspace(1e-4, 0.02, 1000), 0)
    a_hat = alphas_cumprod[30]
    with torch.no_grad(), seeded_noise(11):
        eps = model(x, t_tensor, y=y)[:, :3]
        u, post = pd.estimate_score_update_posterior(5, model, None, x, y, t_tensor, eps, x.clone(), a_hat)
    mask = pd.calculate_threshold_map(0.9, None, u, "higher")
    eps_new = post * mask + eps * (1 - mask)
    save("posterior_update", x=x, y=y, eps=eps, u=u, post=post, mask=mask, eps_new=eps_new, a_hat=a_hat, t=300, M=5)

    # ---- SD percentile guidance fn (posterior and gradient modes), B=1 with CFG-doubled input
    import diffusion_uncertainty.uncertainty_guidance as ug
    sd = ToySDUNet(4, seed=12).eval()
    g = torch.Generator().manual_seed(12)
    lat = torch.randn(1, 4, 32, 32, generator=g)
    lat2 = torch.cat([lat] * 2)
    emb = torch.randn(2, 8, 16, generator=g)
    t_tensor = torch.tensor(501)
    a_hat = alphas_cumprod[501]
    out = {}
    for mode, use_post in (("post", True), ("grad", False)):
        ug.use_posterior = use_post
        with seeded_noise(12), quiet():
            un, tx = sd(lat2, t_tensor, emb)[0].chunk(2)
            eps = (un + 7.5 * (tx - un)).detach().clone()
            res = ug.get_uncertainty_guided_score_with_percentile(
                eps, lat2.clone(), t_tensor, emb.clone(), sd, a_hat, 0.9, "stable-diffusion",
                num_uncertainty_samples=5, guidance_scale=7.5, lr=0.7)
        out[f"{mode}_eps"] = eps.detach()
        out[f"{mode}_out"] = res.detach()
    ug.use_posterior = True
    save("sd_percentile_guidance", lat=lat, emb=emb, t=501, a_hat=a_hat, **out)


def main_dpm():
    """`dpm_2_uncertainty_centered` (SU/scheduling_dpm_2_uncertainty_centered.py): the factory builds it from the DDPM
    scheduler's config (get_uncertainty_scheduler.py:31-32).  `python tests/golden/make_golden.py dpm` writes only these."""
    torch.manual_seed(0)
    torch.set_num_threads(1)
    cls = "KDPM2SchedulerUncertaintyImagenetClassConditioned"
    run_scheduler_case("sched_dpm2", "scheduling_dpm_2_uncertainty_centered", cls,
                       dict(M=4, after_step=10, num_steps_uc=5), n_steps=20, seed=20,
                       cfg_over=dict(variance_type="fixed_small"))
    run_scheduler_case("sched_dpm2_heun_short", "scheduling_dpm_2_uncertainty_centered", cls,
                       dict(M=3, after_step=4, num_steps_uc=4, solver_type="heun", final_sigmas_type="sigma_min"), n_steps=12, seed=21,
                       cfg_over=dict(variance_type="learned_range", timestep_spacing="linspace", beta_schedule="squaredcos_cap_v2"))
    run_scheduler_case("sched_dpm2_order1", "scheduling_dpm_2_uncertainty_centered", cls,
                       dict(M=2, after_step=2, num_steps_uc=3, solver_order=1), n_steps=8, seed=22,
                       cfg_over=dict(variance_type="fixed_small", timestep_spacing="trailing", beta_schedule="scaled_linear"))


def main_uncond():
    """The unconditioned (CIFAR-10) L4 loop, generate_samples.py:366-463, with the Cifar10 scheduler class of
    `uncertainty_centered` (the zig-zag module of the reference has no Cifar10 class).  `python tests/golden/make_golden.py uncond`."""
    import diffusion_uncertainty.generate_samples as gs
    from tests.toy_models import ToyUNet2D3
    torch.manual_seed(0)
    torch.set_num_threads(1)
    mod = __import__(SU + "scheduling_ddim_uncertainty_centered", fromlist=["x"])
    model = ToyUNet2D3(3, seed=30).eval()
    g = torch.Generator().manual_seed(130)
    x_T = torch.randn(5, 3, 16, 16, generator=g)
    with quiet():
        sched = mod.DDIMSchedulerUncertaintyCifar10.from_config(base_config(), unet=model, M=3, after_step=14, num_steps_uc=5)
        sched.set_timesteps(20)
        with seeded_noise(78):
            res = gs.generate_samples_model_scheduler_unconditioned_from_tensor(
                X_T=x_T, batch_size=2, device=torch.device("cpu"), model=model, scheduler=sched)
    save("l4_unconditioned", x_T=x_T, gen_images=res["gen_images"], uncertainty=res["uncertainty"], score=res["score"])


def main_uvit():
    """The U-ViT latent loop, generate_samples.py:469-571 (`uvit_ae(x, t, y)`, `uvit_ae.decode` before the uint8 epilogue), with
    a toy subclass of the reference's UViTAE so that the reference's isinstance dispatch (traits.py:12) takes the U-ViT
    branch.  `python tests/golden/make_golden.py uvit`."""
    import diffusion_uncertainty.generate_samples as gs
    from diffusion_uncertainty.uvit.uvit_ae import UViTAE as RefUViTAE
    from tests.toy_models import ToyUViTMixin

    class ToyRefUViTAE(ToyUViTMixin, RefUViTAE):
        def __init__(self, seed):
            torch.nn.Module.__init__(self)
            self.toy_init(seed)

    torch.manual_seed(0)
    torch.set_num_threads(1)
    mod = __import__(SU + "scheduling_ddim_uncertainty_zigzag_centered", fromlist=["x"])
    model = ToyRefUViTAE(40).eval()
    g = torch.Generator().manual_seed(140)
    x_T = torch.randn(5, 4, 8, 8, generator=g)
    y = torch.randint(0, 10, (5,), generator=g)
    with quiet():
        sched = mod.DDIMSchedulerUncertaintyImagenetClassConditioned.from_config(
            base_config(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, clip_sample=False, set_alpha_to_one=False,
                        steps_offset=1), unet=model, M=3, after_step=14, num_steps_uc=5, num_zigzag=2)
        sched.set_timesteps(20)
        with seeded_noise(79):
            res = gs.generate_samples_model_scheduler_class_conditioned_uvit_from_tensor(
                X_T=x_T, y=y, batch_size=2, uvit_ae=model, scheduler=sched, device=torch.device("cpu"))
    save("l4_uvit", x_T=x_T, y=y, gen_images=res["gen_images"], uncertainty=res["uncertainty"], score=res["score"],
         timestep=res["timestep"])


class _StepRecorder:
    """Wraps scheduler.step: keeps the last x_{t-1} (the pipelines return uint8 images only) and, with neutral=True, rewinds the
    seeded noise generator after every step so that the scheduler's own draws (`best_noise`, every step) do not consume the
    stream — what diffusers' DDIMScheduler, the scheduler the reference drives these pipelines with
    (scripts/generate_images_with_uncertainty_threshold.py:202-203), does not draw at all."""

    def __init__(self, sched, gen=None, neutral=False):
        self.sched, self.gen, self.neutral, self.last = sched, gen, neutral, None
        self.inner = sched.step
        sched.step = self

    def __call__(self, *a, **kw):
        state = self.gen.get_state() if self.neutral else None
        out = self.inner(*a, **kw)
        if self.neutral:
            self.gen.set_state(state)
        self.last = out.prev_sample.detach().clone()
        return out


def _ref_plain_ddim(n_steps, **cfg):
    """The reference's own DDIM arithmetic as a plain scheduler: the centred uncertainty scheduler with an EMPTY window."""
    mod = __import__(SU + "scheduling_ddim_uncertainty_centered", fromlist=["x"])

    class DDIMScheduler(mod.DDIMSchedulerUncertainty):
        def set_timesteps(self, n, device=None):
            self.config.after_step, self.config.num_steps_uc = 0, 1
            super().set_timesteps(n, device)
            self.timestep_after_step, self.timestep_end_step = -1, 10 ** 9

    with quiet():
        sched = DDIMScheduler.from_config(base_config(**cfg), unet=None, M=1)
        sched.set_timesteps(n_steps)
    return sched


def main_pipelines():
    """Fixtures of the pipeline classes and threshold-guided loops (VERDICT r1 missing 1-3): `python tests/golden/make_golden.py pipelines`."""
    torch.manual_seed(0)
    torch.set_num_threads(1)
    cpu = torch.device("cpu")
    from diffusion_uncertainty.pipeline_uncertainty import pipeline_sampler_class_conditional_uncertainty as pu
    from diffusion_uncertainty.pipeline_uncertainty import \
        pipeline_sampler_class_conditional_uncertainty_guided_posterior_distribution as pd
    from diffusion_uncertainty.pipeline_uncertainty import pipeline_sampler_class_conditional_uncertainty_guided_second_order as so
    from diffusion_uncertainty.pipeline_uncertainty import uncertainty_guidance as pug
    import diffusion_uncertainty.generate_samples as gs
    from diffusion_uncertainty.uvit.uvit_ae import UViTAE as RefUViTAE
    from tests.toy_models import ToyUViTMixin

    # ---- DiffusionClassConditionalWithUncertainty (:9-147) with the BASELINE scheduler, two batches
    mod = __import__(SU + "scheduling_ddim_uncertainty_zigzag_centered", fromlist=["x"])
    model = ToyADM(3, seed=50).eval()
    g = torch.Generator().manual_seed(150)
    x_T = torch.randn(6, 3, 16, 16, generator=g)
    y = torch.randint(0, 10, (6,), generator=g)
    with quiet():
        sched = mod.DDIMSchedulerUncertaintyImagenetClassConditioned.from_config(base_config(), unet=model, M=3, after_step=12,
                                                                                 num_steps_uc=4, num_zigzag=2)
        sched.set_timesteps(20)
        pipe = pu.DiffusionClassConditionalWithUncertainty(model, sched, 16, cpu, 4, 0)
        with seeded_noise(80):
            res = pipe(X_T=x_T, y=y)
    save("pipe_with_uncertainty", x_T=x_T, y=y, gen_images=res["gen_images"], uncertainty=res["uncertainty"], score=res["score"])

    # ---- DiffusionClassConditionalGuidedPosteriorDistribution (:71-243), percentile and tensor thresholds.  Its __call__ passes
    # three arguments to the four-argument module function (:159): the missing threshold_type is supplied here, nothing else changes.
    orig = pd.calculate_threshold_map
    pd.calculate_threshold_map = lambda thr, i, u, kind="higher": orig(thr, i, u, kind)
    try:
        for tag, thr in (("q", 0.9), ("t", None)):
            model = ToyADM(3, seed=51).eval()
            g = torch.Generator().manual_seed(151)
            x_T = torch.randn(5, 3, 16, 16, generator=g)
            y = torch.randint(0, 10, (5,), generator=g)
            if thr is None:
                thr = torch.rand(8, 3, 16, 16, generator=g) * 2e-3
            sched = _ref_plain_ddim(8)
            pipe = pd.DiffusionClassConditionalGuidedPosteriorDistribution(model, sched, thr, 16, cpu, 3, 0, M=4)
            with seeded_noise(81) as gen, quiet():
                rec = _StepRecorder(sched, gen, neutral=True)
                finals = []
                orig_cat = None
                res = pipe(X_T=x_T, y=y, start_step=2, num_steps=3)
            save(f"pipe_posterior_{tag}", x_T=x_T, y=y, gen_images=res["gen_images"], final_last_batch=rec.last,
                 **({"threshold": thr} if torch.is_tensor(thr) else {"q": thr}))
    finally:
        pd.calculate_threshold_map = orig

    # ---- DiffusionClassConditionalGuidedSecondOrder (:71-330)
    model = ToyADM(3, seed=52).eval()
    g = torch.Generator().manual_seed(152)
    x_T = torch.randn(5, 3, 16, 16, generator=g)
    y = torch.randint(0, 10, (5,), generator=g)
    sched = _ref_plain_ddim(8)
    pipe = so.DiffusionClassConditionalGuidedSecondOrder(model, sched, 0.85, 16, cpu, 3, 0, M=4, threshold_type="higher")
    with seeded_noise(82) as gen, quiet():
        rec = _StepRecorder(sched, gen, neutral=True)
        res = pipe(X_T=x_T, y=y, start_step=2, num_steps=4)
    save("pipe_second_order", x_T=x_T, y=y, gen_images=res["gen_images"], final_last_batch=rec.last, q=0.85)

    # ---- PU/uncertainty_guidance.generate_samples_model_scheduler_class_conditioned_with_threshold (:12-125)
    modc = __import__(SU + "scheduling_ddim_uncertainty_centered", fromlist=["x"])
    model = ToyADMWithParameter(3, seed=53, scale=3.0).eval()   # scale 3: the guided update moves ~23 % of the pixels
    g = torch.Generator().manual_seed(153)
    x_T = torch.randn(5, 3, 16, 16, generator=g)
    y = torch.randint(0, 10, (5,), generator=g)
    thr = torch.rand(8, 3, 16, 16, generator=g) * 0.45
    with quiet():
        sched = modc.DDIMSchedulerUncertaintyImagenetClassConditioned.from_config(base_config(), unet=model, M=2, after_step=2, num_steps_uc=3)
        sched.set_timesteps(8)
        with seeded_noise(83):
            rec = _StepRecorder(sched)
            res = pug.generate_samples_model_scheduler_class_conditioned_with_threshold(
                5, 3, 16, model, sched, 10, thr, device=cpu, x_T=x_T, y=y, start_step=2, num_steps=3)
    save("loop_threshold_adm", x_T=x_T, y=y, threshold=thr, gen_images=res["gen_images"], final_last_batch=rec.last)

    # ---- generate_samples_uvit_scheduler_class_conditioned_with_threshold (generate_samples.py:721-860; BASELINE config 4)
    class ToyRefUViTAE(ToyUViTMixin, RefUViTAE):
        def __init__(self, seed):
            torch.nn.Module.__init__(self)
            self.toy_init(seed)

    model = ToyRefUViTAE(54).eval()
    g = torch.Generator().manual_seed(154)
    x_T = torch.randn(5, 4, 8, 8, generator=g)
    y = torch.randint(0, 10, (5,), generator=g)
    thr = torch.rand(8, 4, 8, 8, generator=g) * 0.02
    with quiet():
        sched = mod.DDIMSchedulerUncertaintyImagenetClassConditioned.from_config(
            base_config(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, clip_sample=False, set_alpha_to_one=False,
                        steps_offset=1), unet=model, M=2, after_step=2, num_steps_uc=3, num_zigzag=2)
        sched.set_timesteps(8)
        with seeded_noise(84):
            rec = _StepRecorder(sched)
            res = gs.generate_samples_uvit_scheduler_class_conditioned_with_threshold(
                5, 3, 8, model, sched, 10, thr, device=cpu, x_T=x_T, y=y, start_step=2, num_steps=3)
    save("loop_threshold_uvit", x_T=x_T, y=y, threshold=thr, gen_images=res["gen_images"], final_last_batch=rec.last)

    # ---- generate_samples_model_scheduler_class_conditioned_with_percentile (generate_samples.py:861-983; legacy)
    model = ToyADMWithParameter(3, seed=55, scale=3.0).eval()
    g = torch.Generator().manual_seed(155)
    x_T = torch.randn(4, 3, 16, 16, generator=g)
    labels = torch.randint(0, 10, (4,), generator=g)
    with quiet():
        sched = modc.DDIMSchedulerUncertaintyImagenetClassConditioned.from_config(base_config(), unet=model, M=2, after_step=2, num_steps_uc=3)
        sched.set_timesteps(8)
        with seeded_noise(85):
            rec = _StepRecorder(sched)
            res = gs.generate_samples_model_scheduler_class_conditioned_with_percentile(
                4, 4, 16, model, sched, labels, 0.9, device=cpu, x_T=x_T, start_step=2, num_steps=3)
    save("loop_percentile_adm", x_T=x_T, y=labels, gen_images=res["gen_images"], final_last_batch=rec.last, q=0.9)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pipelines":
        main_pipelines()
    elif len(sys.argv) > 1 and sys.argv[1] == "uvit":
        main_uvit()
    elif len(sys.argv) > 1 and sys.argv[1] == "uncond":
        main_uncond()
    elif len(sys.argv) > 1 and sys.argv[1] == "dpm":
        main_dpm()
    elif len(sys.argv) > 1 and sys.argv[1] == "widen":
        main_widen()
    else:
        main()
