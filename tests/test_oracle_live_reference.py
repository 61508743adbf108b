"""The oracle against the LIVE reference, where the reference is present (the build container: /root/reference, read-only, imported
with the diffusers base-class stand-in of tests/golden/_standin).  The committed fixtures pin the oracle at a handful of recorded
configurations; here the same comparison runs on configurations and seeds that are NOT in the fixtures — other step counts,
windows, M, zig-zag counts, eta, beta schedules, batch shapes — bit for bit on the CPU.  Skipped on the GPU box (no reference there);
nothing under `-m gpu`, smoke() or bench.py depends on this file."""
import contextlib
import importlib
import io
import os
import sys

import pytest
import torch

from oracle import du_oracle as O
from tests.helpers import l4_sampling_loop
from tests.toy_models import ToyADM, seeded_noise

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "diffusion_uncertainty")),
                                reason="the reference tree is only present in the build container")
SU = "diffusion_uncertainty.schedulers_uncertainty."


@pytest.fixture(scope="module")
def reference_on_path():
    added = [os.path.join(HERE, "golden", "_standin"), REF]
    for p in added:
        sys.path.insert(0, p)
    yield
    for p in added:
        sys.path.remove(p)
    for name in [m for m in sys.modules if m == "diffusers" or m.startswith("diffusers.") or m.startswith("diffusion_uncertainty.")
                 or m == "diffusion_uncertainty"]:
        del sys.modules[name]


def base_config(**over):
    cfg = dict(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", clip_sample=True,
               set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon", timestep_spacing="leading")
    cfg.update(over)
    return cfg


def bits_equal(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype == torch.float32:
        # NaNs in the same places, everything else bit for bit (the sign of a zero included)
        return bool(torch.equal(torch.isnan(a), torch.isnan(b))) and \
            bool(torch.equal(a.nan_to_num(0.0).view(torch.int32), b.nan_to_num(0.0).view(torch.int32)))
    return bool(torch.equal(a, b))


# variant (oracle name), reference module, class, ctor kwargs, n_steps, seed, eta, dropout, config overrides, (B, H)
LIVE_CASES = [
    ("zigzag_centered", "scheduling_ddim_uncertainty_zigzag_centered", dict(M=2, after_step=3, num_steps_uc=6, num_zigzag=4), 12, 101, 0.0,
     False, {}, (3, 8)),
    ("zigzag_centered", "scheduling_ddim_uncertainty_zigzag_centered", dict(M=7, after_step=0, num_steps_uc=3, num_zigzag=1), 10, 102, 0.7,
     False, dict(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, clip_sample=False, set_alpha_to_one=False, steps_offset=1),
     (2, 12)),
    ("zigzag", "scheduling_ddim_uncertainty_zigzag", dict(M=3, after_step=2, num_steps_uc=4, num_zigzag=3), 10, 103, 0.0, False,
     dict(beta_schedule="squaredcos_cap_v2"), (5, 8)),
    ("centered", "scheduling_ddim_uncertainty_centered", dict(M=2, after_step=6, num_steps_uc=3, predict_next=False), 15, 104, 0.2, False, {},
     (2, 8)),
    ("centered", "scheduling_ddim_uncertainty_centered", dict(M=4, after_step=1, num_steps_uc=8, predict_next=True), 10, 105, 0.0, False,
     dict(clip_sample=False), (3, 8)),
    ("infer_noise", "scheduling_ddim_infer_noise", dict(M=3, after_step=4, num_steps_uc=4, predict_next=True), 12, 106, 0.0, False, {}, (2, 8)),
    ("mc_dropout", "scheduling_ddim_mc_dropout", dict(M=3, after_step=2, num_steps_uc=5), 10, 107, 0.0, True, {}, (4, 8)),
    ("threshold", "scheduling_ddim_uncertainty_threshold",
     dict(M=3, after_step=2, num_steps_uc=5, uncertainty_threshold=0.2, uncertainty_threshold_mode="max"), 10, 108, 0.0, False, {}, (3, 8)),
    ("threshold", "scheduling_ddim_uncertainty_threshold",
     dict(M=4, after_step=3, num_steps_uc=4, uncertainty_threshold=-0.3, uncertainty_threshold_mode="min"), 10, 109, 0.0, False, {}, (2, 12)),
    ("multiscale", "scheduling_ddim_infer_noise_multiscale_threshold", dict(M=4, after_step=2, num_steps_uc=6), 12, 110, 0.0, False, {}, (3, 8)),
    ("flip_threshold", "scheduling_ddim_flip_threshold",
     dict(M=1, after_step=2, num_steps_uc=6, uncertainty_threshold=0.0, uncertainty_threshold_mode="min"), 12, 111, 0.0, False, {}, (3, 8)),
    ("uncertainty_grad", "scheduling_ddim_uncertainty_grad", dict(M=2, after_step=3, num_steps_uc=3, predict_next=False), 10, 112, 0.0, False, {},
     (2, 8)),
    ("mc_dropout_gradient", "scheduling_ddim_mc_dropout_gradient", dict(M=3, after_step=1, num_steps_uc=4), 8, 113, 0.0, True, {}, (2, 8)),
]


@pytest.mark.parametrize("case", LIVE_CASES, ids=[f"{c[0]}-{c[5]}" for c in LIVE_CASES])
def test_oracle_scheduler_equals_live_reference(reference_on_path, case):
    variant, module, kw, n_steps, seed, eta, dropout, cfg, (B, H) = case
    g = torch.Generator().manual_seed(500 + seed)
    x_T = torch.randn(B, 3, H, H, generator=g)
    y = torch.randint(0, 10, (B,), generator=g)

    def reference():
        model = ToyADM(3, seed=seed, dropout=dropout).eval()
        cls = getattr(importlib.import_module(SU + module), "DDIMSchedulerUncertaintyImagenetClassConditioned")
        with contextlib.redirect_stdout(io.StringIO()):
            sched = cls.from_config(base_config(**cfg), unet=model, **kw)
            sched.set_timesteps(n_steps)
            with seeded_noise(2000 + seed):
                return sched, l4_sampling_loop(sched, model, x_T, y, eta=eta)

    def oracle():
        model = ToyADM(3, seed=seed, dropout=dropout).eval()
        sched = O.OracleScheduler(variant, None, unet=model, **kw, **cfg)
        sched.predict = lambda x, t: model(x, t, y=sched.prompt_embeds)[:, :3]
        sched.set_timesteps(n_steps)
        with seeded_noise(2000 + seed):
            return sched, l4_sampling_loop(sched, model, x_T, y, eta=eta)

    rs, r = reference()
    os_, o = oracle()
    assert torch.equal(rs.timesteps, os_.timesteps)
    assert int(rs.timestep_after_step) == int(os_.timestep_after_step) and int(rs.timestep_end_step) == int(os_.timestep_end_step)
    assert r["uncertainty"].shape[1] > 0, "the case must have a window"
    for key in ("uncertainty", "score", "final"):
        assert bits_equal(r[key], o[key]), key
    assert len(r["prevs"]) == len(o["prevs"]) and all(bits_equal(a, b) for a, b in zip(r["prevs"], o["prevs"]))


@pytest.mark.parametrize("seed", range(6))
def test_oracle_threshold_map_equals_live_reference(reference_on_path, seed):
    """calculate_threshold_map (PU/...posterior_distribution.py:10-30): scalar percentile (both threshold types, ties, a NaN row,
    ranks >= 2 incl. the [B, L, C] layout) and the tensor-threshold branch"""
    mod = importlib.import_module("diffusion_uncertainty.pipeline_uncertainty."
                                  "pipeline_sampler_class_conditional_uncertainty_guided_posterior_distribution")
    g = torch.Generator().manual_seed(900 + seed)
    shapes = [(3, 3, 8, 8), (2, 4, 16, 16), (5, 7), (2, 33, 6), (1, 4, 64, 64), (4, 1, 5, 9)]
    shape = shapes[seed % len(shapes)]
    u = torch.rand(shape, generator=g) ** 3
    if seed % 2:
        u = (u * 8).round() / 8               # heavy ties
    if seed == 3:
        u[0].view(-1)[1] = float("nan")       # NaN row -> NaN threshold -> all-false mask
    for q in (0.05, 0.5, 0.9, 0.99):
        for kind in ("higher", "lower"):
            assert bits_equal(mod.calculate_threshold_map(q, None, u, kind), O.calculate_threshold_map(q, None, u, kind)), (q, kind)
    if len(shape) == 4:
        thr = torch.rand((4,) + shape[1:], generator=g).half()
        for i in (0, 3):
            for kind in ("higher", "lower"):
                assert bits_equal(mod.calculate_threshold_map(thr, i, u, kind), O.calculate_threshold_map(thr, i, u, kind))


@pytest.mark.parametrize("seed", range(4))
def test_oracle_posterior_update_equals_live_reference(reference_on_path, seed):
    """estimate_score_update_posterior (PU/...posterior_distribution.py:32-68: F7 re-noising, M forwards, F1c, F5 with the batch-axis
    sum of the LAST perturbed prediction) on random shapes, M and alpha_hat — against the oracle's restatement of the same lines"""
    from oracle import du_oracle_pipelines as P
    mod = importlib.import_module("diffusion_uncertainty.pipeline_uncertainty."
                                  "pipeline_sampler_class_conditional_uncertainty_guided_posterior_distribution")
    g = torch.Generator().manual_seed(700 + seed)
    B, H = [(2, 8), (4, 16), (1, 8), (3, 4)][seed]
    M = [2, 5, 16, 3][seed]
    t = [900, 500, 180, 20][seed]
    model = ToyADM(3, seed=40 + seed).eval()
    x = torch.randn(B, 3, H, H, generator=g)
    y = torch.randint(0, 10, (B,), generator=g)
    t_tensor = torch.full((B,), t, dtype=torch.long)
    ac = torch.cumprod(1.0 - O.make_betas(), dim=0)
    a = ac[[3, 20, 41, 49][seed]]                    # the pipeline indexes alphas_cumprod by the STEP number (:151)
    with torch.no_grad():
        eps = model(x, t_tensor, y=y)[:, :3]
        with seeded_noise(3000 + seed):
            u_r, post_r = mod.estimate_score_update_posterior(M, model, None, x, y, t_tensor, eps, x, a)
        with seeded_noise(3000 + seed):
            preds = P.perturbed_predictions(lambda z: model(z, t_tensor, y=y)[:, :3], x, eps, x, a, M)
        u_o = O.variance_with_center(preds, eps)
        post_o = O.posterior_blend(eps, u_o, torch.ones_like(u_o), M, a, sum_source=preds[-1], batch_sum=True)
    assert bits_equal(u_r, u_o)
    assert bits_equal(post_r, post_o)


class _Recorder:
    """keeps the last x_{t-1} of a wrapped scheduler.step and rewinds the seeded generator after every step (the plain scheduler the
    pipelines are driven with draws nothing; the reference's uncertainty scheduler, used here with an empty window, draws every step)"""

    def __init__(self, sched, gen):
        self.gen, self.inner, self.last = gen, sched.step, None
        sched.step = self

    def __call__(self, *a, **kw):
        state = self.gen.get_state()
        out = self.inner(*a, **kw)
        self.gen.set_state(state)
        self.last = out.prev_sample.detach().clone()
        return out


def _plain_reference_ddim(n_steps):
    """the reference's own DDIM arithmetic as a plain scheduler: its centred uncertainty scheduler with an EMPTY window"""
    mod = importlib.import_module(SU + "scheduling_ddim_uncertainty_centered")

    class DDIMScheduler(mod.DDIMSchedulerUncertainty):
        def set_timesteps(self, n, device=None):
            self.config.after_step, self.config.num_steps_uc = 0, 1
            super().set_timesteps(n, device)
            self.timestep_after_step, self.timestep_end_step = -1, 10 ** 9

    with contextlib.redirect_stdout(io.StringIO()):
        sched = DDIMScheduler.from_config(base_config(), unet=None, M=1)
        sched.set_timesteps(n_steps)
    return sched


# kind, N images, batch size, n_steps, start_step, num_steps, M, threshold (float percentile or "tensor"), seed
PIPE_CASES = [
    ("posterior", 4, 4, 10, 0, 2, 2, 0.5, 201),
    ("posterior", 7, 3, 5, 1, 3, 6, 0.95, 202),        # ragged last batch; window reaching the last step (`start + num >= i`)
    ("posterior", 3, 2, 10, 4, 1, 3, "tensor", 203),
    ("second_order", 4, 4, 10, 0, 3, 2, 0.5, 204),
    ("second_order", 5, 2, 5, 2, 5, 5, 0.99, 205),     # ragged last batch; window longer than the loop
    ("second_order", 3, 3, 10, 3, 2, 3, "tensor", 206),
]


@pytest.mark.parametrize("case", PIPE_CASES, ids=[f"{c[0]}-{c[-1]}" for c in PIPE_CASES])
def test_oracle_pipeline_equals_live_reference(reference_on_path, case):
    """DiffusionClassConditionalGuidedPosteriorDistribution / ...GuidedSecondOrder.__call__ (PU/...posterior_distribution.py:117-188,
    PU/...second_order.py:110-194) against oracle/du_oracle_pipelines.py, bit for bit.  The posterior class calls the four-argument
    module function with three arguments (:159): the missing threshold_type is supplied, nothing else is changed."""
    from oracle import du_oracle_pipelines as P
    kind, N, bs, n_steps, start, num, M, thr, seed = case
    name = "pipeline_sampler_class_conditional_uncertainty_guided_" + ("posterior_distribution" if kind == "posterior" else "second_order")
    mod = importlib.import_module("diffusion_uncertainty.pipeline_uncertainty." + name)
    g = torch.Generator().manual_seed(seed)
    x_T = torch.randn(N, 3, 8, 8, generator=g)
    y = torch.randint(0, 10, (N,), generator=g)
    if thr == "tensor":
        thr = torch.rand(n_steps, 3, 8, 8, generator=g) * (2e-3 if kind == "posterior" else 1e-3)
    cpu = torch.device("cpu")
    ac = torch.cumprod(1 - O.make_betas(), 0)

    model = ToyADM(3, seed=seed).eval()
    sched = _plain_reference_ddim(n_steps)
    orig = mod.calculate_threshold_map
    if kind == "posterior":
        mod.calculate_threshold_map = lambda t, i, u, threshold_type="higher": orig(t, i, u, threshold_type)
    try:
        if kind == "posterior":
            pipe = mod.DiffusionClassConditionalGuidedPosteriorDistribution(model, sched, thr, 8, cpu, bs, 0, M=M)
        else:
            pipe = mod.DiffusionClassConditionalGuidedSecondOrder(model, sched, thr, 8, cpu, bs, 0, M=M, threshold_type="higher")
        with seeded_noise(4000 + seed) as gen, contextlib.redirect_stdout(io.StringIO()):
            rec = _Recorder(sched, gen)
            res = pipe(X_T=x_T, y=y, start_step=start, num_steps=num)
    finally:
        mod.calculate_threshold_map = orig

    model = ToyADM(3, seed=seed).eval()
    fn = P.posterior_pipeline if kind == "posterior" else P.second_order_pipeline
    with seeded_noise(4000 + seed):
        imgs, last = fn(model, x_T, y, thr, batch_size=bs, n_steps=n_steps, start_step=start, num_steps=num, M=M, ac=ac)
    assert bits_equal(rec.last, last)
    assert bits_equal(res["gen_images"], imgs)


DPM_LIVE_CASES = [
    # ctor kwargs, n_steps, seed, config overrides, (B, H)
    (dict(M=2, after_step=1, num_steps_uc=6), 9, 301, dict(variance_type="fixed_small"), (3, 8)),
    (dict(M=5, after_step=0, num_steps_uc=2, solver_type="heun"), 6, 302,
     dict(variance_type="fixed_small", beta_schedule="scaled_linear", timestep_spacing="trailing"), (2, 12)),
    (dict(M=3, after_step=3, num_steps_uc=3, solver_order=1, final_sigmas_type="sigma_min"), 10, 303,
     dict(variance_type="learned_range", timestep_spacing="linspace"), (4, 8)),
]


@pytest.mark.parametrize("case", DPM_LIVE_CASES, ids=[str(c[2]) for c in DPM_LIVE_CASES])
def test_oracle_dpm2_scheduler_equals_live_reference(reference_on_path, case):
    """SU/scheduling_dpm_2_uncertainty_centered.py (DPM-Solver++ multistep with the centred map) against oracle/du_oracle_dpm.py"""
    from oracle.du_oracle_dpm import OracleDPM2Scheduler
    kw, n_steps, seed, cfg, (B, H) = case
    g = torch.Generator().manual_seed(500 + seed)
    x_T = torch.randn(B, 3, H, H, generator=g)
    y = torch.randint(0, 10, (B,), generator=g)

    model = ToyADM(3, seed=seed).eval()
    cls = getattr(importlib.import_module(SU + "scheduling_dpm_2_uncertainty_centered"), "KDPM2SchedulerUncertaintyImagenetClassConditioned")
    with contextlib.redirect_stdout(io.StringIO()):
        rs = cls.from_config(base_config(**cfg), unet=model, **kw)
        rs.set_timesteps(n_steps)
        with seeded_noise(2000 + seed):
            r = l4_sampling_loop(rs, model, x_T, y)

    model = ToyADM(3, seed=seed).eval()
    full = base_config(**cfg)        # (the oracle scheduler reads the default beta range; the cases keep it)
    os_ = OracleDPM2Scheduler(None, **kw, **{k: full[k] for k in ("timestep_spacing", "steps_offset", "beta_schedule", "prediction_type",
                                                                  "num_train_timesteps")}, variance_type=cfg.get("variance_type"))
    os_.predict = lambda x, t: model(x, t, y=os_.prompt_embeds)[:, :3]
    os_.set_timesteps(n_steps)
    with seeded_noise(2000 + seed):
        o = l4_sampling_loop(os_, model, x_T, y)
    assert torch.equal(rs.timesteps, os_.timesteps) and r["uncertainty"].shape[1] > 0
    for key in ("uncertainty", "score", "final"):
        assert bits_equal(r[key], o[key]), key


@pytest.mark.parametrize("seed", range(3))
def test_oracle_sd_guidance_equals_live_reference(reference_on_path, seed):
    """get_uncertainty_guided_score_with_percentile, posterior mode (uncertainty_guidance.py:61-131; the Stable Diffusion call site with
    the CFG-doubled latent) for other latent sizes, M, percentiles and guidance scales than the fixture's"""
    from tests.toy_models import ToySDUNet
    import diffusion_uncertainty.uncertainty_guidance as ug
    H, M, q, gs, t = [(16, 3, 0.8, 5.0, 801), (64, 16, 0.9, 7.5, 401), (8, 2, 0.5, 1.5, 21)][seed]
    sd = ToySDUNet(4, seed=60 + seed).eval()
    g = torch.Generator().manual_seed(600 + seed)
    lat2 = torch.cat([torch.randn(1, 4, H, H, generator=g)] * 2)
    emb = torch.randn(2, 8, 16, generator=g)
    t_tensor = torch.tensor(t)
    a_hat = torch.cumprod(1 - O.make_betas(), 0)[t]
    saved = ug.use_posterior
    ug.use_posterior = True
    try:
        with seeded_noise(5000 + seed), contextlib.redirect_stdout(io.StringIO()):
            un, tx = sd(lat2, t_tensor, emb)[0].chunk(2)
            eps = (un + gs * (tx - un)).detach().clone()
            ref = ug.get_uncertainty_guided_score_with_percentile(eps, lat2.clone(), t_tensor, emb.clone(), sd, a_hat, q, "stable-diffusion",
                                                                  num_uncertainty_samples=M, guidance_scale=gs).detach()
    finally:
        ug.use_posterior = saved
    with torch.no_grad(), seeded_noise(5000 + seed):
        un, tx = sd(lat2, t_tensor, emb)[0].chunk(2)
        eps_o = un + gs * (tx - un)
        x0 = (lat2 - torch.sqrt(1 - a_hat) * eps_o) / torch.sqrt(a_hat)
        scores = []
        for _ in range(M):
            x_hat = O.perturb_add_noise(x0, torch.randn_like(eps_o), a_hat)
            un, tx = sd(x_hat, t_tensor, emb)[0].chunk(2)
            scores.append(un + gs * (tx - un))
        u = O.variance_with_center(scores, eps_o)
        out = O.posterior_blend(eps_o, u, O.calculate_threshold_map(q, None, u, "higher"), M, a_hat, batch_sum=True)
    assert bits_equal(eps, eps_o) and bits_equal(ref, out)
