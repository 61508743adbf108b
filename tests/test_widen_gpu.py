"""The rows added either side of the core path, through the C ABI on the GPU against the CPU oracle /
the reference's own torch expressions: flip maps (N4), the backward of the M-axis reduction (F6 gradient schedulers),
per-pixel threshold fitting (N2), per-image reductions (N3), and the broadcast-mask / linear-combination forms of
du_guided_step those schedulers use.  Bit-exact wherever the reference expression is one rounding per operation."""
import os

import numpy as np
import pytest
import torch

from oracle import du_oracle as O
from tests.test_ops_gpu import assert_close_rel, bits_equal, coeffs_for, dev, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from diffusion_uncertainty_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------------------- N4 flips
@pytest.mark.parametrize("shape", [(3, 3, 16, 16), (2, 4, 8, 12), (2, 3, 5, 7), (1, 1, 1, 4), (0, 3, 8, 8)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_flip_h_is_torch_flip(ops, shape, dtype):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g).to(dtype)
    assert bits_equal(ops.flip_h(x.to(dev())), torch.flip(x, dims=[2]))
    if shape[0] > 0:   # channel-slice view of a wider tensor (the ADM [:, :3] view), read in place
        wide = torch.randn(shape[0], 2 * shape[1], shape[2], shape[3], generator=g).to(dtype)
        assert bits_equal(ops.flip_h(wide.to(dev())[:, :shape[1]]), torch.flip(wide[:, :shape[1]], dims=[2]))


@pytest.mark.parametrize("shape", [(3, 3, 16, 16), (2, 4, 8, 12), (2, 3, 5, 7)])
@pytest.mark.parametrize("amax", [False, True])
def test_flip_sqdiff_matches_reference_expression(ops, shape, amax):
    g = torch.Generator().manual_seed(2)
    eps, f = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    f[0, 0, 1, 1] = float("nan")
    want = O.flip_uncertainty(eps, torch.flip(f, dims=[2]), channel_amax=amax)
    got = ops.flip_sqdiff(eps.to(dev()), f.to(dev()), channel_amax=amax)
    assert got.shape == want.shape and bits_equal(got, want)
    buf = torch.zeros((shape[0], 3) + tuple(want.shape[1:]), device=dev())     # straight into an accumulation slot
    ops.flip_sqdiff(eps.to(dev()), f.to(dev()), channel_amax=amax, out=buf[:, 1])
    assert bits_equal(buf[:, 1], want) and float(buf[:, 0].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------- F6 backward of the reduction
@pytest.mark.parametrize("mode", ["var", "centered", "var_with_center"])
@pytest.mark.parametrize("M", [2, 5, 9])
def test_moments_backward_matches_torch_autograd(ops, mode, M):
    eps, scores, _ = synth(3, 3, 16, M, seed=M, spread=0.3)
    gu = torch.randn(eps.shape, generator=torch.Generator().manual_seed(3))
    sc = [s.clone().requires_grad_(True) for s in scores]
    c = eps.clone().requires_grad_(True)
    if mode == "var":
        u = torch.var(torch.stack(sc, 0), dim=0)
    elif mode == "centered":
        u = (torch.stack(sc, 0) - c.unsqueeze(0)).pow(2).mean(dim=0)
    else:
        u = torch.var(torch.stack(sc + [c], 0), dim=0)
    u.backward(gu)
    grads, gc = ops.moments_backward([s.to(dev()) for s in scores], gu.to(dev()), mode, center=None if mode == "var" else eps.to(dev()),
                                     need_center=mode != "var")
    for a, b in zip(grads, sc):
        assert_close_rel(a, b.grad, 1e-5, atol=1e-6)
    if mode != "var":
        assert_close_rel(gc, c.grad, 1e-5, atol=1e-6)


def test_moments_autograd_through_a_model(ops):
    """`u.mean(dim=0).sum().backward()` through du_moments / du_moments_backward and a torch model equals pure torch"""
    from tests.toy_models import ToyADM
    model = ToyADM(3, seed=4).eval()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 3, 16, 16, generator=g)
    noises = [torch.randn(3, 3, 16, 16, generator=g) for _ in range(4)]

    def run(device, reduce):
        m = model.to(device)
        xi = x.to(device).clone().requires_grad_(True)
        sc = [m(xi + 0.1 * n.to(device), 300)[:, :3] for n in noises]
        u = reduce(sc)
        u.mean(dim=0).sum().backward()
        return u.detach().cpu(), xi.grad.cpu()

    u_ref, g_ref = run(torch.device("cpu"), lambda sc: torch.var(torch.stack(sc, 0), dim=0))
    u_k, g_k = run(dev(), lambda sc: ops.moments_autograd(sc, "var"))
    assert_close_rel(u_k, u_ref, 1e-5, atol=1e-12)
    assert_close_rel(g_k, g_ref, 1e-4, atol=1e-7)


# ------------------------------------------------------------------------------------------- N2 / N3
def test_pixel_threshold_fitting_matches_the_reference_script(ops, golden_dir):
    g = {k: v for k, v in np.load(os.path.join(golden_dir, "pixel_thresholds.npz")).items()}
    unc = torch.from_numpy(g["unc"]).to(dev())
    for perc in (0.15, 0.9):
        got = ops.fit_pixel_thresholds(unc, perc)
        assert bits_equal(got, torch.from_numpy(g[f"thr_{perc}"])), perc
    # every k on one timestep slice (a strided [:, i] view of the accumulated maps), fp32 and the script's .half() variant
    for k in (0, 1, 17, 35, 36):
        assert bits_equal(ops.column_kth(unc[:, 1], k), O.column_kth(unc[:, 1].cpu(), k))
    h = unc.half()
    assert bits_equal(ops.column_kth(h[:, 2], 11), O.column_kth(h[:, 2].cpu(), 11))
    with pytest.raises(IndexError):
        ops.column_kth(unc[:, 0], 37)


@pytest.mark.parametrize("shape", [(5, 4, 3, 16, 16), (3, 2, 3, 5, 7), (0, 2, 3, 4, 4)])
def test_per_image_reductions(ops, shape):
    g = torch.Generator().manual_seed(6)
    x = torch.rand(shape, generator=g) ** 3
    assert_close_rel(ops.row_sum(x.to(dev())), x.sum(dim=(1, 2, 3, 4)), 1e-5, atol=1e-6)
    assert_close_rel(ops.slot_sum(x.to(dev())), x.sum(dim=1), 1e-6, atol=1e-7)
    assert ops.slot_sum(x.to(dev())).shape == x.sum(dim=1).shape


# ------------------------------------------------------------------------------------------- guided-step forms of those schedulers
def test_guided_step_broadcast_mask_over_channels(ops):
    """flip_threshold: a [B,1,H,W] mask multiplies a [B,C,H,W] prediction (SU/scheduling_ddim_flip_threshold.py:541-561)"""
    eps, _, sample = synth(4, 3, 16, 1, seed=8)
    w = (torch.rand(4, 1, 16, 16, generator=torch.Generator().manual_seed(8)) > 0.4).float()
    c, k = coeffs_for(ops, 300, 280)
    prev, x0, e2 = O.masked_restep(eps, sample, w, c)
    r = ops.guided_step(eps.to(dev()), sample.to(dev()), k, guidance="weights", mask=w.to(dev()), want_x0=True)
    assert bits_equal(r["prev"], prev) and bits_equal(r["x0"], x0) and bits_equal(r["eps"], e2)
    with pytest.raises(ValueError):
        ops.guided_step(eps.to(dev()), sample.to(dev()), k, guidance="weights", mask=torch.ones(4, 1, 16, 15, device=dev()))


@pytest.mark.parametrize("form", ["grad_add", "lincomb"])
def test_guided_step_gradient_forms_with_unguided_x0(ops, form):
    """eps' = eps + g*abar (uncertainty_grad.py:551) / 0.9 eps + 0.1 g (mc_dropout_gradient.py:514); x0 from the unguided eps"""
    eps, _, sample = synth(3, 3, 16, 1, seed=9)
    gr = torch.randn(eps.shape, generator=torch.Generator().manual_seed(9))
    clip = form == "grad_add"
    c, k = coeffs_for(ops, 300, 280, clip_sample=clip)
    if form == "grad_add":
        e2 = eps + gr * c.alpha_prod_t
        r = ops.guided_step(eps.to(dev()), sample.to(dev()), k, guidance="grad_add", aux=gr.to(dev()), lam=float(c.alpha_prod_t),
                            x0_unguided=True, want_x0=True)
    else:
        e2 = 0.9 * eps + 0.1 * gr
        r = ops.guided_step(eps.to(dev()), sample.to(dev()), k, guidance="lincomb", aux=gr.to(dev()), post_M=0.9, lam=0.1,
                            x0_unguided=True, want_x0=True)
    x0 = (sample - c.sqrt_beta_t * eps) / c.sqrt_alpha_t
    if clip:
        x0 = x0.clamp(-1, 1)
    prev = c.sqrt_alpha_prev * x0 + c.dir_coef * e2
    assert bits_equal(r["eps"], e2) and bits_equal(r["x0"], x0) and bits_equal(r["prev"], prev)


# ------------------------------------------------------------------------------------------- the guided pipelines' pieces
@pytest.mark.parametrize("wrt", ["input", "score"])
def test_guided_gradient_pipeline_score_update(ops, golden_dir, wrt):
    """DiffusionClassConditionalGuidedGradient.estimate_score_update + blend against the fixture recorded from the reference
    class (gradients: fp32 tolerance, the reduction's backward runs in du_moments_backward)"""
    from diffusion_uncertainty_b200.pipeline_uncertainty.pipeline_sampler_class_conditional_uncertainty_guided_gradient import (
        DiffusionClassConditionalGuidedGradient, guided_gradient_blend)
    from tests.toy_models import ToyADMWithParameter, seeded_noise
    g = {k: v for k, v in np.load(os.path.join(golden_dir, "gradient_update.npz")).items()}
    model = ToyADMWithParameter(3, seed=14).eval().to(dev())
    x, y = torch.from_numpy(g["x"]).to(dev()), torch.from_numpy(g["y"]).to(dev())
    t_tensor = torch.full((x.shape[0],), int(g["t"]), dtype=torch.long, device=dev())
    a_hat = torch.from_numpy(g["a_hat"])
    pipe = DiffusionClassConditionalGuidedGradient(model, None, 0.9, 16, dev(), x.shape[0], 0, M=int(g["M"]), gradient_wrt=wrt,
                                                   lambda_update=0.1)
    with torch.no_grad():
        eps = model(x, t_tensor, y=y)[:, :3].clone()
    with seeded_noise(14):
        u, upd = pipe.estimate_score_update(x.clone(), y, 7, t_tensor, eps, x.clone(), a_hat)
    assert_close_rel(u, torch.from_numpy(g[f"{wrt}_u"]), 1e-5, atol=1e-12)
    assert_close_rel(upd, torch.from_numpy(g[f"{wrt}_update"]), 1e-4, atol=1e-7)
    # the blend is bit-exact given the recorded inputs
    new = guided_gradient_blend(eps, torch.from_numpy(g[f"{wrt}_update"]).to(dev()), torch.from_numpy(g[f"{wrt}_mask"]).to(dev()), 0.1)
    assert bits_equal(new, torch.from_numpy(g[f"{wrt}_eps_new"]))


def test_second_order_blend(ops):
    from diffusion_uncertainty_b200.pipeline_uncertainty.pipeline_sampler_class_conditional_uncertainty_guided_second_order import (
        second_order_blend, second_order_momentum_update)
    from tests.toy_models import seeded_noise
    eps, scores, _ = synth(3, 3, 16, 3, seed=5)
    u = O.centered_second_moment(scores, eps)
    m = O.calculate_threshold_map(0.8, None, u, "higher")
    with seeded_noise(5):
        want = eps + u * torch.sign(torch.randn_like(eps)) * m      # second_order.py:249
    with seeded_noise(5):
        got = second_order_blend(eps.to(dev()), u.to(dev()), m.to(dev()))
    assert bits_equal(got, want)
    mom, corr, root = second_order_momentum_update(None, u.to(dev()), 3)
    assert bits_equal(mom, u)
    assert_close_rel(root, torch.sqrt(u / (1 - 0.99 ** 3 + 1e-5)), 1e-6)   # plain torch ops on either device


def test_guided_gradient_pipeline_call(ops):
    """the whole __call__ loop on the GPU: deterministic, uint8 images; with percentile 1.0 nothing exceeds the threshold, so
    the guided run equals plain sampling with the same scheduler"""
    from diffusion_uncertainty_b200.pipeline_uncertainty.pipeline_sampler_class_conditional_uncertainty_guided_gradient import \
        DiffusionClassConditionalGuidedGradient
    from diffusion_uncertainty_b200.schedulers_uncertainty.scheduling_ddim_flip import DDIMSchedulerUncertaintyImagenetClassConditioned
    from tests.toy_models import ToyADMWithParameter, seeded_noise
    model = ToyADMWithParameter(3, seed=3).eval().to(dev())
    g = torch.Generator().manual_seed(3)
    X_T, y = torch.randn(5, 3, 16, 16, generator=g), torch.randint(0, 10, (5,), generator=g)

    def run(threshold, lam):
        sched = DDIMSchedulerUncertaintyImagenetClassConditioned(unet=model, after_step=0, num_steps_uc=1)
        sched.set_timesteps(10)
        pipe = DiffusionClassConditionalGuidedGradient(model, sched, threshold, 16, dev(), 3, 0, M=3, lambda_update=lam)
        with seeded_noise(30):
            return pipe(X_T=X_T, y=y, start_step=4, num_steps=3)

    a, b = run(0.9, 0.5), run(0.9, 0.5)
    assert a["gen_images"].dtype == torch.uint8 and a["gen_images"].shape == (5, 3, 16, 16) and torch.equal(a["gen_images"], b["gen_images"])
    assert torch.equal(a["x_t"], X_T) and torch.equal(a["y"], y)
    off, plain = run(1.0, 0.5), run(0.9, 0.0)
    assert torch.equal(off["gen_images"], plain["gen_images"])
    assert not torch.equal(a["gen_images"], plain["gen_images"])


def test_threshold_fitting_on_disk_contract(tmp_path):
    """N2 with the reference's file layout: scripts/compute_threshold_pixel_wise.py:86-116, 143-157 on the same files."""
    import yaml
    from diffusion_uncertainty_b200.threshold_fitting import fit_and_save_thresholds
    g = torch.Generator().manual_seed(9)
    folders = []
    for k, n in enumerate((120, 150)):
        d = tmp_path / f"run{k}"
        d.mkdir()
        torch.save(torch.rand(n, 3, 3, 8, 8, generator=g) ** 3, d / "uncertainty_0.pth")
        torch.save(torch.rand(40, 3, 3, 8, 8, generator=g), d / "uncertainty_1.pth")        # < 100 samples: skipped
        torch.save(torch.zeros(1, dtype=torch.uint8), d / "gen_images_0.pth")
        (d / "args.yaml").write_text(yaml.safe_dump({"dataset": "imagenet64", "scheduler_type": "uncertainty_centered", "generation_steps": 50}))
        folders.append(str(d))
    perc = 0.9

    def ref_fit(u):          # the reference's expressions, on the CPU
        out = []
        for i in range(u.shape[1]):
            ut = u[:, i]
            idx = ut.argsort(dim=0)[int(u.shape[0] * perc)].unsqueeze(0)
            out.append(ut.gather(dim=0, index=idx).squeeze(0))
        return torch.stack(out, dim=0)

    for used in (folders[:1], folders):
        paths = fit_and_save_thresholds(used, perc, str(tmp_path / "results"), extra_args={"on_cpu": False})
        assert paths["thresholds"].endswith("results/thresholds/imagenet64/thresholds_uncertainty_centered_perc=0.9.pth")
        got = torch.load(paths["thresholds"])
        per_file = [ref_fit(torch.load(os.path.join(f, "uncertainty_0.pth")).half()).unsqueeze(0) for f in used]
        want = per_file[0].squeeze(0) if len(per_file) == 1 else ref_fit(torch.cat(per_file, dim=0))
        assert got.dtype == torch.float16 and got.shape == (3, 3, 8, 8) and torch.equal(got, want)
        cfg = yaml.safe_load(open(paths["config"]))
        assert cfg["dataset_config"]["generation_steps"] == 50 and cfg["perc"] == perc and cfg["dataset_folders"] == used


# ---- round 2: the former eager-torch leftovers as kernels ------------------------------------------------------------------
def test_second_order_blend_sign_add_is_bit_exact(ops):
    """eps + u * sign(n) * mask (PU/..._guided_second_order.py:249), one du_guided_step launch, against the torch expression."""
    g = torch.Generator().manual_seed(2)
    eps, n = torch.randn(3, 3, 16, 16, generator=g), torch.randn(3, 3, 16, 16, generator=g)
    u = torch.rand(3, 3, 16, 16, generator=g) ** 2
    mask = (torch.rand(3, 3, 16, 16, generator=g) > 0.7).float()
    n[0, 0, 0, 0] = 0.0
    n[0, 0, 0, 1] = float("nan")
    want = eps + u * torch.sign(n) * mask
    got = ops.guided_step(eps.to(dev()), None, None, guidance="sign_add", u=u.to(dev()), mask=mask.to(dev()), aux=n.to(dev()), want_eps=True)["eps"]
    assert bits_equal(got, want)


def test_legacy_mul_blend_is_bit_exact(ops):
    """eps (1 - m) + eps m g (generate_samples.py:953)."""
    g = torch.Generator().manual_seed(3)
    eps, grad = torch.randn(2, 3, 16, 16, generator=g), torch.randn(2, 3, 16, 16, generator=g)
    mask = (torch.rand(2, 3, 16, 16, generator=g) > 0.5).float()
    want = eps * (1 - mask) + eps * mask * grad
    got = ops.guided_step(eps.to(dev()), None, None, guidance="mul_blend", mask=mask.to(dev()), aux=grad.to(dev()), want_eps=True)["eps"]
    assert bits_equal(got, want)


def test_std_map_and_its_backward(ops):
    """pred_epsilons.std(dim=0) and d(std.mean(0).sum())/d scores (generate_samples.py:941-943) against torch autograd."""
    g = torch.Generator().manual_seed(4)
    scores = [torch.randn(2, 3, 8, 8, generator=g) for _ in range(5)]
    ref = [s.clone().requires_grad_(True) for s in scores]
    torch.stack(ref, 0).std(dim=0).mean(dim=0).sum().backward()
    mine = [s.to(dev()).requires_grad_(True) for s in scores]
    u = ops.moments_autograd(mine, "std")
    u.mean(dim=0).sum().backward()
    assert_close_rel(u.detach(), torch.stack(scores, 0).std(dim=0), 1e-5)
    for a, b in zip(mine, ref):
        assert_close_rel(a.grad, b.grad, 1e-4, atol=1e-7)


def test_perturb_rows_and_per_sample_add_noise(ops):
    """add_noise / get_velocity with a VECTOR of timesteps (SU/...zigzag_centered.py:606-646): one du_perturb_rows launch."""
    from diffusion_uncertainty_b200.schedulers_uncertainty.scheduling_ddim import DDIMScheduler
    g = torch.Generator().manual_seed(5)
    x, nz = torch.randn(4, 3, 8, 8, generator=g), torch.randn(4, 3, 8, 8, generator=g)
    t = torch.tensor([10, 500, 999, 0])
    s = DDIMScheduler()
    ac = s.alphas_cumprod
    sa, sb = (ac[t] ** 0.5).view(-1, 1, 1, 1), ((1 - ac[t]) ** 0.5).view(-1, 1, 1, 1)
    assert bits_equal(s.add_noise(x.to(dev()), nz.to(dev()), t.to(dev())), sa * x + sb * nz)
    assert bits_equal(s.get_velocity(x.to(dev()), nz.to(dev()), t.to(dev())), sa * nz - sb * x)
    assert bits_equal(ops.scale(x.to(dev()), 0.37), 0.37 * x)
    # the differentiable forms used by the threshold-guided loops
    e = x.to(dev()).requires_grad_(True)
    x0 = ops.x0_autograd(e, nz.to(dev()), 0.8, 0.6)
    out = ops.perturb_autograd(x0, nz.to(dev()), 0.8, 0.6)
    out.sum().backward()
    er = x.clone().requires_grad_(True)
    (0.8 * ((nz - 0.6 * er) / 0.8) + 0.6 * nz).sum().backward()
    assert bits_equal(x0.detach(), (nz - 0.6 * x) / 0.8)
    assert_close_rel(e.grad, er.grad, 1e-6)
