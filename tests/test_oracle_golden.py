"""Pin the CPU oracle (oracle/du_oracle.py) against the golden vectors recorded from the UNMODIFIED
reference (tests/golden/make_golden.py) and against known-answer facts of torch.quantile / torch.var
(SURVEY.md §8c).  CPU-only."""
import os

import numpy as np
import pytest
import torch

from oracle import du_oracle as O
from tests.helpers import l4_sampling_loop
from tests.toy_models import ToyADM, seeded_noise


def load(golden_dir, name):
    return {k: v for k, v in np.load(os.path.join(golden_dir, name + ".npz")).items()}


def T(a):
    return torch.from_numpy(np.asarray(a))


def same(a, b):
    """bit-for-bit, NaNs in the same places"""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def restated_threshold_is(thr_torch, x2d, q):
    """torch.quantile's threshold equals the restatement row by row, with the lerp rounded once per operation or fused — which one is
    a property of the host's torch build / CPU (oracle.lerp_torch); everything in front of the lerp is exact"""
    sep, _, _ = O.quantile_linear_rows(x2d, q, lerp_fma=False)
    fus, _, _ = O.quantile_linear_rows(x2d, q, lerp_fma=True)
    t = np.asarray(thr_torch, dtype=np.float32).reshape(-1)
    ok = [(np.isnan(t[b]) and np.isnan(sep[b].item())) or t[b].view(np.int32) in (sep[b].numpy().view(np.int32), fus[b].numpy().view(np.int32))
          for b in range(t.shape[0])]
    return all(ok)


SCHED_CASES = [
    # name, variant, ctor kwargs, n_steps, seed, eta, dropout, cfg
    ("sched_zigzag_centered", "zigzag_centered", dict(M=5, after_step=40, num_steps_uc=10, num_zigzag=3), 50, 0, 0.0, False, {}),
    ("sched_zigzag_centered_eta", "zigzag_centered", dict(M=3, after_step=5, num_steps_uc=4, num_zigzag=2), 20, 1, 0.3, False,
     dict(beta_schedule="squaredcos_cap_v2", clip_sample=False)),
    ("sched_zigzag", "zigzag", dict(M=4, after_step=10, num_steps_uc=5, num_zigzag=2), 20, 2, 0.0, False, {}),
    ("sched_centered", "centered", dict(M=5, after_step=10, num_steps_uc=5, predict_next=False), 20, 3, 0.0, False, {}),
    ("sched_centered_next", "centered", dict(M=5, after_step=10, num_steps_uc=5, predict_next=True), 20, 4, 0.0, False, {}),
    ("sched_infer_noise", "infer_noise", dict(M=5, after_step=10, num_steps_uc=5, predict_next=False), 20, 5, 0.0, False, {}),
    ("sched_mc_dropout", "mc_dropout", dict(M=6, after_step=10, num_steps_uc=5), 20, 6, 0.0, True, {}),
    ("sched_threshold_max", "threshold", dict(M=5, after_step=10, num_steps_uc=5, uncertainty_threshold=1.0,
                                              uncertainty_threshold_mode="max"), 20, 7, 0.0, False, {}),
    ("sched_threshold_min", "threshold", dict(M=5, after_step=10, num_steps_uc=5, uncertainty_threshold=0.5,
                                              uncertainty_threshold_mode="min", predict_next=True), 20, 8, 0.0, False, {}),
    ("sched_multiscale", "multiscale", dict(M=5, after_step=10, num_steps_uc=5), 20, 9, 0.0, False, {}),
    # rows added after the core path (tests/golden/make_golden.py widen)
    ("sched_flip", "flip", dict(M=1, after_step=10, num_steps_uc=5), 20, 10, 0.0, False, {}),
    ("sched_flip_threshold", "flip_threshold", dict(M=1, after_step=10, num_steps_uc=5, uncertainty_threshold=0.5,
                                                    uncertainty_threshold_mode="max"), 20, 11, 0.0, False, {}),
    ("sched_uncertainty_grad", "uncertainty_grad", dict(M=4, after_step=10, num_steps_uc=5, predict_next=False), 20, 12, 0.0, False, {}),
    ("sched_mc_dropout_gradient", "mc_dropout_gradient", dict(M=4, after_step=10, num_steps_uc=5), 20, 13, 0.0, True, {}),
]
# the reference's `flip` fixture was recorded with DDIMSchedulerUncertaintyImagenet, whose predict_model passes no class label
NO_LABEL_VARIANTS = ("flip",)


@pytest.mark.parametrize("case", SCHED_CASES, ids=[c[0] for c in SCHED_CASES])
def test_oracle_scheduler_matches_reference(golden_dir, case):
    name, variant, kw, n_steps, seed, eta, dropout, cfg = case
    g = load(golden_dir, name)
    model = ToyADM(3, seed=seed, dropout=dropout).eval()
    x_T, y = T(g["x_T"]), T(g["y"])
    sched = O.OracleScheduler(variant, None, unet=model, **kw, **cfg)
    sched.predict = (lambda x, t: model(x, t)[:, :3]) if variant in NO_LABEL_VARIANTS else \
        (lambda x, t: model(x, t, y=sched.prompt_embeds)[:, :3])
    sched.set_timesteps(n_steps)
    assert sched.timestep_after_step == int(g["after"]) and sched.timestep_end_step == int(g["end"])
    assert same(sched.timesteps.numpy(), g["timesteps"])
    with seeded_noise(1000 + seed):
        res = l4_sampling_loop(sched, model, x_T, y, eta=eta)
    # the oracle repeats the reference's fp32 operations one for one -> bit-exact on CPU
    assert same(res["uncertainty"].numpy(), g["uncertainty"])
    assert same(res["score"].numpy(), g["score"])
    assert same(res["final"].numpy(), g["final"])


DPM_CASES = [
    # name, ctor kwargs, n_steps, seed, config overrides (tests/golden/make_golden.py dpm)
    ("sched_dpm2", dict(M=4, after_step=10, num_steps_uc=5), 20, 20, dict(variance_type="fixed_small", timestep_spacing="leading")),
    ("sched_dpm2_heun_short", dict(M=3, after_step=4, num_steps_uc=4, solver_type="heun", final_sigmas_type="sigma_min"), 12, 21,
     dict(variance_type="learned_range", timestep_spacing="linspace", beta_schedule="squaredcos_cap_v2")),
    ("sched_dpm2_order1", dict(M=2, after_step=2, num_steps_uc=3, solver_order=1), 8, 22,
     dict(variance_type="fixed_small", timestep_spacing="trailing", beta_schedule="scaled_linear")),
]


@pytest.mark.parametrize("case", DPM_CASES, ids=[c[0] for c in DPM_CASES])
def test_oracle_dpm2_scheduler_matches_reference(golden_dir, case):
    from oracle.du_oracle_dpm import OracleDPM2Scheduler
    name, kw, n_steps, seed, cfg = case
    g = load(golden_dir, name)
    model = ToyADM(3, seed=seed).eval()
    x_T, y = T(g["x_T"]), T(g["y"])
    sched = OracleDPM2Scheduler(None, **kw, **cfg)
    sched.predict = lambda x, t: model(x, t, y=sched.prompt_embeds)[:, :3]
    sched.set_timesteps(n_steps)
    assert sched.timestep_after_step == int(g["after"]) and sched.timestep_end_step == int(g["end"])
    assert same(sched.timesteps.numpy(), g["timesteps"])
    with seeded_noise(1000 + seed):
        res = l4_sampling_loop(sched, model, x_T, y)
    assert same(res["uncertainty"].numpy(), g["uncertainty"])
    assert same(res["score"].numpy(), g["score"])
    assert same(res["final"].numpy(), g["final"])


def test_oracle_unconditioned_loop_matches_reference(golden_dir):
    """generate_samples_model_scheduler_unconditioned_from_tensor (generate_samples.py:366-463) with the Cifar10 class of
    `uncertainty_centered`: oracle scheduler + restated loop, bit for bit"""
    from tests.helpers import l4_unconditioned_loop
    from tests.toy_models import ToyUNet2D3
    g = load(golden_dir, "l4_unconditioned")
    model = ToyUNet2D3(3, seed=30).eval()
    sched = O.OracleScheduler("centered", None, unet=model, M=3, after_step=14, num_steps_uc=5)
    sched.predict = lambda x, t: model(x, t).sample
    sched.set_timesteps(20)
    with seeded_noise(78):
        res = l4_unconditioned_loop(sched, model, T(g["x_T"]), 2)
    assert same(res["gen_images"].numpy(), g["gen_images"])
    assert same(res["uncertainty"].numpy(), g["uncertainty"]) and same(res["score"].numpy(), g["score"])


UVIT_CFG = dict(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012, clip_sample=False, set_alpha_to_one=False,
                steps_offset=1)


def test_oracle_uvit_loop_matches_reference(golden_dir):
    """generate_samples_model_scheduler_class_conditioned_uvit_from_tensor (generate_samples.py:469-571), U-ViT scheduler
    config of uvit/load_pretrained_models.py:45-57: oracle scheduler + restated loop, bit for bit"""
    from tests.helpers import l4_uvit_loop
    from tests.toy_models import UViTAE
    g = load(golden_dir, "l4_uvit")
    model = UViTAE(40).eval()
    sched = O.OracleScheduler("zigzag_centered", None, unet=model, M=3, after_step=14, num_steps_uc=5, num_zigzag=2, **UVIT_CFG)
    sched.predict = lambda x, t: model(x, t if torch.is_tensor(t) else torch.full((x.shape[0],), int(t)), sched.prompt_embeds)
    sched.set_timesteps(20)
    assert same(sched.timesteps.numpy(), g["timestep"])
    with seeded_noise(79):
        res = l4_uvit_loop(sched, model, T(g["x_T"]), T(g["y"]), 2)
    assert same(res["gen_images"].numpy(), g["gen_images"])
    assert same(res["uncertainty"].numpy(), g["uncertainty"]) and same(res["score"].numpy(), g["score"])


def test_quantile_restatement_is_torch_quantile(golden_dir):
    g = load(golden_dir, "threshold_map")
    for tag in "abcdef":
        u, q = T(g[f"{tag}_u"]), float(g[f"{tag}_q"])
        thr, ranks, vals = O.quantile_linear_rows(u.flatten(1), q)
        assert restated_threshold_is(g[f"{tag}_thr"], u.flatten(1), q), tag       # (recorded from torch.quantile on the build host)
        kind = "higher" if bool(g[f"{tag}_higher"]) else "lower"
        assert same(O.calculate_threshold_map(q, None, u, kind).numpy(), g[f"{tag}_mask"]), tag
        # mask rebuilt from the restated threshold
        thr_b = thr.view(-1, *([1] * (u.dim() - 1)))
        m = (u > thr_b) if kind == "higher" else (u < thr_b)
        assert same(m.float().numpy(), g[f"{tag}_mask"]), tag
    assert same(O.calculate_threshold_map(T(g["t_thr"]), 2, T(g["t_u"]), "higher").numpy(), g["t_mask_hi"])
    assert same(O.calculate_threshold_map(T(g["t_thr"]), 3, T(g["t_u"]), "lower").numpy(), g["t_mask_lo"])


@pytest.mark.parametrize("n,q,lo,w", [(3072, .95, 2917, .44995), (12288, .9, 11058, .29980), (49152, .9, 44235, .8984375),
                                      (16384, .9, 14744, .69921875), (4096, .95, 3890, .25)])
def test_quantile_rank_known_answers(n, q, lo, w):
    """SURVEY.md §8c known-answer facts, and bit-equality with torch.quantile on random rows."""
    l, h, ww = O.quantile_rank(n, q)
    assert l == lo and h == lo + 1 and abs(float(ww) - w) < 1e-4
    x = torch.rand(3, n, generator=torch.Generator().manual_seed(n)) ** 2
    thr, _, _ = O.quantile_linear_rows(x, q)
    assert restated_threshold_is(torch.quantile(x, q, dim=1).numpy(), x, q)
    assert int((x[0] > thr[0]).sum()) == n - 1 - lo  # tie-free row


def test_quantile_edge_cases():
    x = torch.rand(2, 7)
    thr, _, _ = O.quantile_linear_rows(x, 0.5)
    assert restated_threshold_is(torch.quantile(x, 0.5, dim=1).numpy(), x, 0.5)
    x[0, 3] = float("nan")
    thr, _, _ = O.quantile_linear_rows(x, 0.5)
    assert np.isnan(thr[0]) and not np.isnan(thr[1])
    assert O.calculate_threshold_map(0.5, None, x, "higher")[0].sum() == 0  # NaN threshold -> all False
    with pytest.raises(RuntimeError):
        torch.quantile(torch.zeros(2, 8, dtype=torch.float16), 0.5, dim=1)
    assert torch.isnan(O.variance_unbiased([torch.ones(3)])).all()  # M == 1
    assert float(torch.var(torch.tensor([1.0, 2.0, 4.0]))) == pytest.approx(7.0 / 3.0)  # unbiased by default


def test_posterior_update_matches_reference(golden_dir):
    g = load(golden_dir, "posterior_update")
    model = ToyADM(3, seed=11).eval()
    x, y, eps = T(g["x"]), T(g["y"]), T(g["eps"])
    t = int(g["t"]); M = int(g["M"]); a_hat = T(g["a_hat"])
    t_tensor = torch.full((x.shape[0],), t, dtype=torch.long)
    from math import sqrt
    with torch.no_grad(), seeded_noise(11):
        eps2 = model(x, t_tensor, y=y)[:, :3]
        assert same(eps2.numpy(), eps.numpy())
        # PU/...posterior_distribution.py:52-57: math.sqrt on a 0-dim tensor -> python float scalars
        x0 = (x - sqrt(1 - a_hat) * eps) / sqrt(a_hat)
        scores = []
        for _ in range(M):
            x_hat = sqrt(a_hat) * x0 + sqrt(1 - a_hat) * torch.randn_like(x)
            scores.append(model(x_hat, t_tensor, y=y)[:, :3])
    u = O.variance_with_center(scores, eps)
    assert same(u.numpy(), g["u"])
    mask = O.calculate_threshold_map(0.9, None, u, "higher")
    assert same(mask.numpy(), g["mask"])
    eps_new = O.posterior_blend(eps, u, mask, M, a_hat, sum_source=scores[-1], batch_sum=True)
    # reference blends as post*mask + eps*(1-mask) (posterior_distribution.py:160); addition commutes
    assert same(eps_new.numpy(), g["eps_new"])


def test_sd_percentile_guidance_matches_reference(golden_dir):
    from tests.toy_models import ToySDUNet
    g = load(golden_dir, "sd_percentile_guidance")
    sd = ToySDUNet(4, seed=12).eval()
    lat, emb, a_hat = T(g["lat"]), T(g["emb"]), T(g["a_hat"])
    lat2 = torch.cat([lat] * 2)
    t_tensor = torch.tensor(int(g["t"]))
    eps = T(g["post_eps"])
    with torch.no_grad(), seeded_noise(12):
        un, tx = sd(lat2, t_tensor, emb)[0].chunk(2)
        assert same((un + 7.5 * (tx - un)).numpy(), eps.numpy())
        x0 = (lat2 - torch.sqrt(1 - a_hat) * eps) / torch.sqrt(a_hat)   # uncertainty_guidance.py:86
        scores = []
        for _ in range(5):
            x_hat = O.perturb_add_noise(x0, torch.randn_like(eps), a_hat)
            un, tx = sd(x_hat, t_tensor, emb)[0].chunk(2)
            scores.append(un + 7.5 * (tx - un))
    u = O.variance_with_center(scores, eps)
    mask = O.calculate_threshold_map(0.9, None, u, "higher")
    out = O.posterior_blend(eps, u, mask, 5, a_hat, batch_sum=True)
    assert same(out.numpy(), g["post_out"])


@pytest.mark.parametrize("wrt", ["input", "score"])
def test_gradient_score_update_matches_reference(golden_dir, wrt):
    """estimate_score_update of the guided-gradient pipeline + its blend lines, recorded from the unmodified reference"""
    from tests.toy_models import ToyADMWithParameter
    g = load(golden_dir, "gradient_update")
    model = ToyADMWithParameter(3, seed=14).eval()
    x, y = T(g["x"]), T(g["y"])
    t_tensor = torch.full((x.shape[0],), int(g["t"]), dtype=torch.long)
    a_hat = T(g["a_hat"])
    predict = lambda inp: model(inp, t_tensor, y=y)[:, :3]   # noqa: E731
    with torch.no_grad():
        eps = predict(x).clone()
    with seeded_noise(14):
        u, upd = O.gradient_score_update(predict, x, eps, x, int(g["M"]), a_hat, wrt)
    assert same(u.numpy(), g[f"{wrt}_u"]) and same(upd.numpy(), g[f"{wrt}_update"])
    m = O.calculate_threshold_map(0.9, None, u, "higher")
    assert same(m.numpy(), g[f"{wrt}_mask"])
    assert same(O.gradient_blend_masked(eps, upd, m, 0.1).numpy(), g[f"{wrt}_eps_new"])


def test_pixel_threshold_restatement_matches_reference_script(golden_dir):
    g = load(golden_dir, "pixel_thresholds")
    for perc in (0.15, 0.9):
        assert same(O.fit_pixel_thresholds(T(g["unc"]), perc).numpy(), g[f"thr_{perc}"])


# ---- round 2: the pipeline classes, restated in oracle/du_oracle_pipelines.py, against fixtures recorded from the reference classes
@pytest.mark.parametrize("tag", ["q", "t"])
def test_oracle_posterior_pipeline_replays_the_reference(golden_dir, tag):
    from oracle import du_oracle_pipelines as P
    from tests.toy_models import ToyADM, seeded_noise
    g = load(golden_dir, f"pipe_posterior_{tag}")
    model = ToyADM(3, seed=51).eval()
    thr = float(g["q"]) if tag == "q" else T(g["threshold"])
    ac = torch.cumprod(1 - O.make_betas(), 0)
    with seeded_noise(81):
        imgs, last = P.posterior_pipeline(model, T(g["x_T"]), T(g["y"]), thr, batch_size=3, n_steps=8, start_step=2, num_steps=3, M=4, ac=ac)
    assert same(last.numpy(), g["final_last_batch"]) and same(imgs.numpy(), g["gen_images"])


def test_oracle_second_order_pipeline_replays_the_reference(golden_dir):
    from oracle import du_oracle_pipelines as P
    from tests.toy_models import ToyADM, seeded_noise
    g = load(golden_dir, "pipe_second_order")
    model = ToyADM(3, seed=52).eval()
    ac = torch.cumprod(1 - O.make_betas(), 0)
    with seeded_noise(82):
        imgs, last = P.second_order_pipeline(model, T(g["x_T"]), T(g["y"]), float(g["q"]), batch_size=3, n_steps=8, start_step=2, num_steps=4,
                                             M=4, ac=ac)
    assert same(last.numpy(), g["final_last_batch"]) and same(imgs.numpy(), g["gen_images"])


def test_quantile_restatement_equals_torch_quantile_property():
    """hypothesis: the sort -> rank -> two-branch-lerp restatement (what the select kernels implement) reproduces torch.quantile(dim=1)
    on the CPU for arbitrary row lengths, quantiles, value scales and ties: the two order statistics exactly, the threshold bit for bit
    with the lerp either rounded per operation or fused (a property of the host: both forms are restated exactly)"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(n=st.integers(1, 3000), B=st.integers(1, 3), q=st.floats(0.0, 1.0, width=32), seed=st.integers(0, 2 ** 31 - 1),
           scale=st.sampled_from([1e-12, 1e-4, 1.0, 1e6]), levels=st.sampled_from([0, 3, 17, 256]))
    def check(n, B, q, seed, scale, levels):
        g = torch.Generator().manual_seed(seed)
        x = torch.rand(B, n, generator=g) * scale
        if levels:
            x = (torch.rand(B, n, generator=g) * levels).floor() * (scale / levels)      # ties (and exact zeros)
        _, ranks, vals = O.quantile_linear_rows(x, q)
        assert restated_threshold_is(torch.quantile(x, q, dim=1).numpy(), x, q), (n, q)
        srt = torch.sort(x, dim=1).values
        assert torch.equal(vals[:, 0], srt[:, int(ranks[0, 0])]) and torch.equal(vals[:, 1], srt[:, int(ranks[0, 1])])

    check()


def test_fused_multiply_add_restatement():
    """oracle.fma_f32 (exact rational arithmetic, one rounding) against cases whose single-rounded value is known"""
    f = np.float32
    assert O.fma_f32(2, 3, -6) == 0 and O.fma_f32(3e38, 10, 0) == f(np.inf) and np.isnan(O.fma_f32(np.nan, 1, 1))
    a = f(1.0 + 2.0 ** -12)                              # a^2 - 1 = 2^-11 + 2^-24: separate roundings lose the 2^-24, one rounding keeps it
    assert O.fma_f32(a, a, -1.0) == f(2.0 ** -11 + 2.0 ** -24) and f(f(a * a) - f(1.0)) == f(2.0 ** -11)
    assert f(2.0 ** -11 + 2.0 ** -24) != f(2.0 ** -11)
    assert O.fma_f32(f(2.0 ** -100), f(2.0 ** -49), 0.0) == f(2.0 ** -149)      # smallest subnormal, exact
    assert O.fma_f32(f(2.0 ** -100), f(2.0 ** -50), 0.0) == 0.0                 # half of it: ties to even
    assert O.fma_f32(f(3 * 2.0 ** -100), f(2.0 ** -50), 0.0) == f(2.0 ** -148)  # 1.5 subnormal quanta: rounds to even (2)
    rng = np.random.default_rng(3)
    for _ in range(300):                                 # where the exact result fits a double, the double computation is the answer
        x, y, z = f(rng.normal()), f(rng.normal()), f(rng.normal() * 1e-3)
        d = float(x) * float(y) + float(z)
        if float(f(d)) == d:
            assert O.fma_f32(x, y, z) == f(d)
