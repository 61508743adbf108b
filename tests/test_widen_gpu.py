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
