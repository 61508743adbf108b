"""GPU parity tests proper: every C-ABI entry point (called through diffusion_uncertainty_b200.ops) against the
CPU oracle on the same seeded inputs.  Bars (BASELINE.json north_star): masks / thresholds / rank indices
bit-exact; mean, variance, x_{t-1} within 1e-5 relative in fp32 (1e-2 with 16-bit inputs)."""
import numpy as np
import pytest
import torch

from oracle import du_oracle as O

pytestmark = pytest.mark.gpu

RTOL32 = 1e-5


def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from diffusion_uncertainty_b200 import ops as _ops
    return _ops


def synth(B, C, H, M, seed=0, spread=0.05, dtype=torch.float32):
    """SURVEY.md §8d synthetic inputs: eps ~ N(0,1), eps_hat_m = eps + spread*N(0,1) (variance << mean^2)."""
    g = torch.Generator().manual_seed(seed)
    eps = torch.randn(B, C, H, H, generator=g)
    scores = [(eps + spread * torch.randn(B, C, H, H, generator=g)).to(dtype) for _ in range(M)]
    sample = torch.randn(B, C, H, H, generator=g)
    return eps.to(dtype), scores, sample


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    denom = b.abs().clamp_min(1e-30)
    return ((a - b).abs() / denom).max().item()


def assert_close_rel(a, b, rtol, atol=0.0):
    a, b = a.double().cpu(), b.double().cpu()
    bad = (a - b).abs() > (rtol * b.abs() + atol)
    assert not bad.any(), f"max rel err {rel_err(a, b):.3e}, {int(bad.sum())} elements out of tolerance"


def bits_equal(a, b):
    return np.array_equal(a.detach().cpu().numpy(), b.detach().cpu().numpy(), equal_nan=True)


# ------------------------------------------------------------------------------------------- F1
@pytest.mark.parametrize("M", [1, 2, 5, 8, 13, 30])
@pytest.mark.parametrize("mode", ["var", "centered", "var_with_center", "raw", "std"])
def test_moments_fp32(ops, M, mode):
    eps, scores, _ = synth(4, 3, 16, M, seed=M)
    d = dev()
    got = ops.moments([s.to(d) for s in scores], center=eps.to(d), mode=mode)
    want = {"var": lambda: O.variance_unbiased(scores), "centered": lambda: O.centered_second_moment(scores, eps),
            "var_with_center": lambda: O.variance_with_center(scores, eps), "raw": lambda: O.raw_second_moment(scores),
            "std": lambda: O.std_over_m(scores)}[mode]()
    if M == 1 and mode in ("var", "std"):
        assert torch.isnan(got).all() and torch.isnan(want).all()
        return
    assert got.dtype == torch.float32 and got.shape == eps.shape
    assert_close_rel(got, want, RTOL32)


def test_moments_mean_and_cancellation(ops):
    """variance 1e-6 of mean^2: the case E[x^2]-E[x]^2 fails; shifted single pass must hold 1e-5."""
    g = torch.Generator().manual_seed(3)
    base = 10.0 + torch.randn(2, 3, 32, 32, generator=g)
    scores = [base + 1e-3 * torch.randn(2, 3, 32, 32, generator=g) for _ in range(5)]
    want = torch.stack([s.double() for s in scores]).var(0)
    got, mean = ops.moments([s.to(dev()) for s in scores], mode="var", return_mean=True)
    assert_close_rel(got, want, RTOL32)
    assert_close_rel(mean, O.mean_over_m(scores), 1e-6)


@pytest.mark.parametrize("dtype,rtol", [(torch.float16, 1e-2), (torch.bfloat16, 1e-2)])
def test_moments_16bit_inputs(ops, dtype, rtol):
    eps, scores, _ = synth(4, 3, 16, 5, seed=9, spread=0.3, dtype=dtype)
    d = dev()
    got = ops.moments([s.to(d) for s in scores], center=eps.to(d), mode="centered")
    want = O.centered_second_moment(scores, eps)  # oracle upcasts the same 16-bit values to fp32
    assert got.dtype == torch.float32
    assert_close_rel(got, want, 1e-5, atol=1e-9)  # fp32 arithmetic on identical inputs: far inside the 1e-2 bar
    got16 = ops.moments([s.to(d) for s in scores], mode="var", out_dtype=dtype)
    assert got16.dtype == dtype
    assert_close_rel(got16.float(), O.variance_unbiased(scores), rtol, atol=1e-6)


def test_moments_channel_slice_view_and_ragged(ops):
    """ADM `model(...)[:, :3]` strided view; odd sizes take the scalar path."""
    d = dev()
    g = torch.Generator().manual_seed(5)
    full = [torch.randn(3, 6, 8, 8, generator=g) for _ in range(5)]
    views_cpu = [f[:, :3] for f in full]
    views_gpu = [f.to(d)[:, :3] for f in full]
    assert not views_gpu[0].is_contiguous()
    assert_close_rel(ops.moments(views_gpu, mode="var"), O.variance_unbiased(views_cpu), RTOL32)
    odd = [torch.randn(3, 1, 7, 5, generator=g) for _ in range(4)]
    ce = torch.randn(3, 1, 7, 5, generator=g)
    assert_close_rel(ops.moments([o.to(d) for o in odd], center=ce.to(d), mode="centered"),
                     O.centered_second_moment(odd, ce), RTOL32)
    # empty batch
    e = ops.moments([torch.empty(0, 3, 4, 4, device=d)] * 2, mode="var")
    assert e.shape == (0, 3, 4, 4)


def test_moments_into_accumulation_slot(ops):
    """F8 fused: write straight into slot [:, t] of the [B,T,C,H,W] buffer."""
    d = dev()
    eps, scores, _ = synth(4, 3, 16, 5, seed=2)
    buf = torch.zeros(4, 6, 3, 16, 16, device=d)
    for t in (0, 3, 5):
        ops.moments([s.to(d) for s in scores], center=eps.to(d), mode="centered", out=buf[:, t])
    want = O.centered_second_moment(scores, eps)
    for t in (0, 3, 5):
        assert_close_rel(buf[:, t], want, RTOL32)
    assert float(buf[:, 1].abs().max()) == 0.0


def test_moments_partial_merge(ops):
    """M-sharding: per-shard (count, mean, M2) merged == the unsharded variance."""
    d = dev()
    eps, scores, _ = synth(2, 4, 16, 16, seed=4)
    shards = [scores[0:4], scores[4:8], scores[8:13], scores[13:16]]
    means, m2s = [], []
    for sh in shards:
        m2, mu = ops.moments([s.to(d) for s in sh], mode="partial", return_mean=True)
        means.append(mu); m2s.append(m2)
    got, mean = ops.moments_merge(means, m2s, [len(s) for s in shards], mode="var", return_mean=True)
    assert_close_rel(got, O.variance_unbiased(scores), RTOL32)
    assert_close_rel(mean, O.mean_over_m(scores), 1e-6, atol=1e-7)
    # centred second moment: plain sum of per-shard sums / M
    cs = [ops.moments([s.to(d) for s in sh], center=eps.to(d), mode="partial") for sh in shards]
    got_c = ops.moments_merge(None, cs, [len(s) for s in shards], mode="centered")
    assert_close_rel(got_c, O.centered_second_moment(scores, eps), RTOL32)


def test_moments_errors(ops):
    d = dev()
    with pytest.raises(RuntimeError):
        ops.moments([torch.zeros(2, 3)], mode="var")  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        ops.moments([torch.zeros(2, 4, device=d)], mode="centered")  # centre missing
    with pytest.raises(ValueError):
        ops.moments([torch.zeros(2, 4, device=d)] * 65, mode="var")  # M > DU_MAX_M


# ------------------------------------------------------------------------------------------- F2a
@pytest.mark.parametrize("shape,q", [((4, 3, 16, 16), 0.9), ((3, 3072), 0.95), ((2, 12288), 0.9), ((2, 49152), 0.9),
                                     ((2, 16384), 0.9), ((3, 4096), 0.95), ((5, 1000), 0.99), ((4, 7), 0.5),
                                     ((3, 1), 0.3), ((2, 5000), 0.0), ((2, 5000), 1.0), ((2, 4097), 0.37)])
def test_quantile_threshold_bit_exact(ops, shape, q):
    g = torch.Generator().manual_seed(int(q * 1000) + shape[-1])
    u = torch.rand(shape, generator=g) ** 3
    thr, ranks, vals = ops.quantile_threshold(u.to(dev()), q, return_details=True)
    want = torch.quantile(u.flatten(1), q, dim=1)
    assert bits_equal(thr, want)
    othr, oranks, ovals = O.quantile_linear_rows(u.flatten(1), q)
    assert bits_equal(thr, othr)
    assert np.array_equal(ranks.cpu().numpy(), oranks.numpy().astype(np.int32))
    assert bits_equal(vals, ovals)


def test_quantile_ties_negatives_nan(ops):
    g = torch.Generator().manual_seed(1)
    u = (torch.randn(6, 2048, generator=g) * 4).round() / 4   # heavy ties, negative values, signed zeros
    u[1] = 0.25                                                # one constant row
    u[2, 100] = float("nan")
    u[3, :5] = torch.tensor([-0.0, 0.0, float("inf"), -float("inf"), 1e-45])
    for q in (0.1, 0.5, 0.9, 0.999):
        thr = ops.quantile_threshold(u.to(dev()), q)
        want = torch.quantile(u, q, dim=1)
        assert bits_equal(thr.abs(), want.abs()) and bits_equal(torch.isnan(thr), torch.isnan(want))
        m = ops.threshold_mask(u.to(dev()), thr, higher=True)
        assert bits_equal(m, O.calculate_threshold_map(q, None, u, "higher"))
        m = ops.threshold_mask(u.to(dev()), thr, higher=False)
        assert bits_equal(m, O.calculate_threshold_map(q, None, u, "lower"))


def test_quantile_errors(ops):
    d = dev()
    with pytest.raises(RuntimeError):
        ops.quantile_threshold(torch.zeros(2, 8, device=d), 1.5)
    with pytest.raises(RuntimeError):
        ops.quantile_threshold(torch.zeros(2, 0, device=d), 0.5)
    with pytest.raises(RuntimeError):
        ops.quantile_threshold(torch.zeros(1, 2 ** 24 + 1, device=d), 0.5)
    assert ops.quantile_threshold(torch.zeros(0, 8, device=d), 0.5).shape == (0,)


def test_quantile_matches_torch_cuda_with_fma_flag(ops):
    """Which lerp does torch's own CUDA kernel use?  lerp_fma=True must reproduce it bit for bit."""
    g = torch.Generator().manual_seed(11)
    u = (torch.rand(64, 3 * 32 * 32, generator=g) ** 3).to(dev())
    for q in (0.9, 0.95, 0.333):
        want = torch.quantile(u, q, dim=1)
        got_fma = ops.quantile_threshold(u, q, lerp_fma=True)
        got_cpu = ops.quantile_threshold(u, q, lerp_fma=False)
        assert bits_equal(got_fma, want) or bits_equal(got_cpu, want)
        assert bits_equal(got_cpu, torch.quantile(u.cpu(), q, dim=1))


def test_tensor_threshold_mask(ops):
    g = torch.Generator().manual_seed(2)
    u = torch.rand(4, 3, 16, 16, generator=g)
    thr = (torch.rand(6, 3, 16, 16, generator=g)).half()
    d = dev()
    for i, kind in ((2, "higher"), (5, "lower")):
        got = ops.tensor_threshold_mask(u.to(d), thr[i].to(d), higher=(kind == "higher"))
        assert bits_equal(got, O.calculate_threshold_map(thr, i, u, kind))


# ------------------------------------------------------------------------------------------- F2c
def test_znorm(ops):
    g = torch.Generator().manual_seed(3)
    u = torch.rand(4, 3, 16, 16, generator=g) ** 2 * 0.01
    d = dev()
    stats = ops.znorm_stats(u.to(d))
    assert abs(stats[0].item() - u.double().mean().item()) <= 1e-6 * abs(u.mean().item())
    assert abs(stats[1].item() - u.double().std().item()) <= 1e-6 * u.std().item()
    assert stats[2].item() == u.numel()
    z, w = ops.znorm_weights(u.to(d), stats, mode="max", thr=1.0)
    zo = O.znorm(u)
    assert_close_rel(z, zo, 1e-5, atol=1e-5)
    safe = (zo - 1.0).abs() > 1e-4     # away from the threshold the mask must agree exactly
    assert bits_equal(w.cpu()[safe], O.znorm_threshold_mask(zo, 1.0, "max")[safe])
    _, w2 = ops.znorm_weights(u.to(d), stats, mode="multiscale", want_z=False)
    wo = O.multiscale_weights(zo)
    safe = torch.stack([(zo - e).abs() > 1e-4 for e in (-1.0, -2.0, -3.0)]).all(0)
    assert bits_equal(w2.cpu()[safe], wo[safe])
    # given the SAME z tensor the weights are bit exact (normalize=False path), band edges included
    zz = torch.tensor([[-3.5, -3.0, -2.5, -2.0, -1.5, -1.0, -0.5, 2.0]])
    _, w3 = ops.znorm_weights(zz.to(d), None, mode="multiscale", normalize=False, want_z=False)
    assert bits_equal(w3, O.multiscale_weights(zz))
    # rank-sharded statistics combine to the global ones (Chan merge)
    parts = torch.stack([ops.znorm_stats(u[:1].to(d)), ops.znorm_stats(u[1:].to(d))])
    comb = ops.znorm_stats_combine(parts)
    assert_close_rel(comb[:2], stats[:2], 1e-6)


# ------------------------------------------------------------------------------------------- F3
def coeffs_for(ops, t, prev_t, eta=0.0, **kw):
    betas = O.make_betas()
    ac = torch.cumprod(1 - betas, 0)
    c = O.DDIMCoeffs(ac, torch.tensor(1.0), t, prev_t, eta)
    k = ops.make_coeffs(c.sqrt_alpha_t.item(), c.sqrt_beta_t.item(), c.sqrt_alpha_prev.item(), c.dir_coef.item(),
                        sigma=float(c.sigma), add_noise=eta > 0, **kw)
    return c, k


@pytest.mark.parametrize("ptype", ["epsilon", "sample", "v_prediction"])
@pytest.mark.parametrize("eta,clip,ucmo", [(0.0, True, False), (0.5, True, True), (1.0, False, False)])
def test_ddim_step_bit_exact(ops, ptype, eta, clip, ucmo):
    eps, _, sample = synth(4, 3, 16, 1, seed=7)
    noise = torch.randn(eps.shape, generator=torch.Generator().manual_seed(8))
    c, k = coeffs_for(ops, 500, 480, eta, clip_sample=clip, prediction_type=ptype, use_clipped_model_output=ucmo)
    d = dev()
    prev, x0, e2 = ops.ddim_step(eps.to(d), sample.to(d), k, noise=noise.to(d) if eta > 0 else None, want_eps=True)
    wp, wx0, we = O.ddim_step(eps, sample, c, ptype, clip, 1.0, eta, noise, ucmo)
    assert bits_equal(x0, wx0) and bits_equal(e2, we) and bits_equal(prev, wp)


def test_ddim_last_step_and_fp16_scores(ops):
    eps, _, sample = synth(2, 3, 16, 1, seed=9)
    c, k = coeffs_for(ops, 0, -20)
    d = dev()
    prev, x0, _ = ops.ddim_step(eps.to(d), sample.to(d), k)
    wp, wx0, _ = O.ddim_step(eps, sample, c)
    assert bits_equal(prev, wp) and bits_equal(x0, wx0)
    h = eps.half()
    prev, x0, _ = ops.ddim_step(h.to(d), sample.to(d), k)
    assert prev.dtype == torch.float32
    wp, _, _ = O.ddim_step(h.float(), sample, c)
    assert_close_rel(prev, wp, 1e-2, atol=1e-6)


# ------------------------------------------------------------------------------------------- F4/F5/F6 (+F3)
def test_guided_posterior_step(ops):
    eps, scores, sample = synth(4, 3, 16, 5, seed=21)
    d = dev()
    c, k = coeffs_for(ops, 300, 280)
    a_hat = torch.cumprod(1 - O.make_betas(), 0)[30]
    for batch_sum in (True, False):
        for kind in ("higher", "lower"):
            u, mask, eps_g, prev, x0 = O.uncertainty_step_posterior(scores, eps, sample, 0.9, 5, a_hat, c, batch_sum=batch_sum,
                                                                   sum_source=scores[-1], threshold_type=kind)
            ug = u.to(d)
            thr = ops.quantile_threshold(ug, 0.9)
            S = ops.batch_sum(scores[-1].to(d)) if batch_sum else scores[-1].to(d)
            if batch_sum:
                assert_close_rel(S, scores[-1].sum(0), 1e-6, atol=1e-6)
                S = scores[-1].sum(0).to(d)      # same S tensor -> the blend must then be bit exact
            r = ops.guided_step(eps.to(d), sample.to(d), k, guidance="posterior", u=ug, thr=thr, aux=S, aux_broadcast=batch_sum,
                                higher=(kind == "higher"), post_M=5.0, inv_alpha_hat=float(1 / a_hat), want_x0=True, want_mask=True)
            assert bits_equal(r["mask"], mask)
            assert bits_equal(r["eps"], eps_g) and bits_equal(r["x0"], x0) and bits_equal(r["prev"], prev)


def test_guided_posterior_zero_variance_propagates_nan(ops):
    """u == 0 -> 1/u = inf -> NaN in the posterior score even where mask == 0 (0*inf), like the reference."""
    d = dev()
    eps = torch.randn(2, 3, 8, 8)
    u = torch.rand(2, 3, 8, 8)
    u[0, 0, 0, :4] = 0.0
    mask = O.calculate_threshold_map(0.5, None, u, "higher")
    want = O.posterior_blend(eps, u, mask, 5, torch.tensor(0.5), batch_sum=False)
    r = ops.guided_step(eps.to(d), None, None, guidance="posterior", u=u.to(d), mask=mask.to(d), aux=eps.to(d), post_M=5.0,
                        inv_alpha_hat=2.0)
    assert bits_equal(r["eps"], want) and torch.isnan(want).any()


def test_guided_gradient_blends(ops):
    d = dev()
    eps, _, sample = synth(3, 3, 16, 1, seed=31)
    g = torch.randn(eps.shape, generator=torch.Generator().manual_seed(32))
    u = torch.rand(eps.shape, generator=torch.Generator().manual_seed(33))
    mask = O.calculate_threshold_map(0.9, None, u, "higher")
    c, k = coeffs_for(ops, 700, 680)
    thr = ops.quantile_threshold(u.to(d), 0.9)
    want = O.gradient_blend_masked(eps, g, mask, 0.3)
    r = ops.guided_step(eps.to(d), sample.to(d), k, guidance="grad_blend", u=u.to(d), thr=thr, aux=g.to(d), lam=0.3)
    assert bits_equal(r["eps"], want) and bits_equal(r["prev"], O.ddim_step(want, sample, c)[0])
    want = O.gradient_add_masked(eps, g, mask, 0.7)
    r = ops.guided_step(eps.to(d), None, None, guidance="grad_add", mask=mask.to(d), aux=g.to(d), lam=0.7)
    assert bits_equal(r["eps"], want)


def test_guided_weights_restep(ops):
    d = dev()
    eps, scores, sample = synth(4, 3, 16, 5, seed=41)
    c, k = coeffs_for(ops, 400, 380)
    for multiscale in (False, True):
        z, w, prev, x0, eps2 = O.uncertainty_step_znorm(scores, eps, sample, 1.0, "max", c, multiscale=multiscale)
        r = ops.guided_step(eps.to(d), sample.to(d), k, guidance="weights", mask=w.to(d), want_x0=True)
        assert bits_equal(r["eps"], eps2) and bits_equal(r["x0"], x0) and bits_equal(r["prev"], prev)


# ------------------------------------------------------------------------------------------- F7 / F8
def test_perturb_bit_exact(ops):
    d = dev()
    x = torch.randn(4, 3, 16, 16, generator=torch.Generator().manual_seed(1))
    nz = torch.randn(4, 3, 16, 16, generator=torch.Generator().manual_seed(2))
    betas = O.make_betas()
    ac = torch.cumprod(1 - betas, 0)
    t = 180
    got = ops.perturb(x.to(d), nz.to(d), torch.sqrt(1 - betas[t]).item(), torch.sqrt(betas[t]).item())
    assert bits_equal(got, O.perturb_predict_next(x, nz, betas[t]))
    got = ops.perturb(x.to(d), nz.to(d), (ac[t] ** 0.5).item(), ((1 - ac[t]) ** 0.5).item())
    assert bits_equal(got, O.perturb_add_noise(x, nz, ac[t]))


def test_accumulate_slot(ops):
    d = dev()
    maps = [torch.rand(3, 2, 8, 8, generator=torch.Generator().manual_seed(s)) for s in range(4)]
    buf = torch.empty(3, 4, 2, 8, 8, device=d)
    for t, m in enumerate(maps):
        ops.accumulate_slot(m.to(d), buf[:, t])
    assert bits_equal(buf, O.accumulate_maps([maps]))
    bufh = torch.empty(3, 4, 2, 8, 8, device=d, dtype=torch.float16)
    ops.accumulate_slot(maps[1].to(d), bufh[:, 1])
    assert bits_equal(bufh[:, 1], maps[1].half())


@pytest.mark.parametrize("B,n_shape,dtype", [(128, (3, 32, 32), torch.float32), (37, (4, 16, 16), torch.float32),
                                             (5, (3, 8, 8), torch.float32), (19, (1, 15, 15), torch.float32),
                                             (64, (4, 32, 32), torch.float16)])
def test_batch_sum(ops, B, n_shape, dtype):
    """du_batch_sum (cluster/DSMEM path for B >= 16 and vectorisable rows, scalar path otherwise) == x.sum(0) in fp64"""
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, *n_shape, generator=g).to(dtype)
    got = ops.batch_sum(x.to(dev()))
    want = x.double().sum(0).float()
    assert bits_equal(got, want)
    out = torch.empty(n_shape, device=dev())
    assert ops.batch_sum(x.to(dev()), out=out) is out and bits_equal(out, want)
