"""CPU restatement of the pipeline classes and threshold-guided loops on the uncertainty path — TEST INFRASTRUCTURE ONLY
(imported by tests/ alone; the product never imports oracle/).  Each function cites the reference lines it follows; paths are
relative to /root/reference/diffusion_uncertainty/ (PU = pipeline_uncertainty).  Pinned: tests/test_oracle_golden.py replays the
fixtures recorded from the reference classes / functions themselves (tests/golden/make_golden.py `pipelines`) through these
functions on the CPU, bit for bit for everything that is not downstream of a model gradient.

The plain DDIM update the guided pipelines are driven with (scripts/generate_images_with_uncertainty_threshold.py:202-203,
diffusers' DDIMScheduler: the F3 block every reference scheduler file copies) is du_oracle.ddim_step.
"""
from math import sqrt
from typing import Callable, Optional, Union

import torch
from torch import Tensor

from . import du_oracle as O


def image_uint8(x: Tensor) -> Tensor:
    """generate_samples.py:203-212 / PU/…posterior_distribution.py:164-175."""
    return ((x / 2 + 0.5).clamp(0, 1) * 255.0).round().to(torch.uint8)


def plain_ddim_step(ac: Tensor, final_alpha: Tensor, n_steps: int, eps: Tensor, t: int, sample: Tensor, clip_sample: bool = True):
    prev_t = t - 1000 // n_steps
    c = O.DDIMCoeffs(ac, final_alpha, t, prev_t, 0.0)
    return O.ddim_step(eps, sample, c, clip_sample=clip_sample)[0]


def perturbed_predictions(predict: Callable, x: Tensor, eps: Tensor, like: Tensor, alpha_hat_t, M: int):
    """PU/…posterior_distribution.py:222-227 (the same in …second_order.py:296-301): x0 = (x - sqrt(1-a) eps)/sqrt(a),
    x_hat = sqrt(a) x0 + sqrt(1-a) randn_like(prev_noisy_sample), M forwards."""
    sa, sb = sqrt(alpha_hat_t), sqrt(1 - alpha_hat_t)
    x0 = (x - sb * eps) / sa
    return [predict(sa * x0 + sb * torch.randn_like(like)) for _ in range(M)]


def posterior_window_step(predict: Callable, x: Tensor, eps: Tensor, alpha_hat_t, M: int, threshold: Union[float, Tensor], i: int,
                          ddim: Callable, threshold_type: str = "higher") -> Tensor:
    """One in-window step of DiffusionClassConditionalGuidedPosteriorDistribution.__call__ (PU/…posterior_distribution.py:141-162
    with estimate_score_update :201-237): F7 -> M forwards -> F1c -> F2a/F2b -> F5 (batch-axis sum of the LAST perturbed
    prediction, :236) -> the scheduler step on the guided score."""
    preds = perturbed_predictions(predict, x, eps, x, alpha_hat_t, M)
    u = torch.var(torch.stack(preds + [eps], dim=0), dim=0, unbiased=True)
    inv_var = 1 / u
    post = (1 / ((M * inv_var) + (1 / alpha_hat_t))) * (inv_var * preds[-1].sum(dim=0))
    mask = O.calculate_threshold_map(threshold, i, u, threshold_type)
    guided = post * mask + eps * (1 - mask)
    return ddim(guided, x)


def posterior_pipeline(model, x_T: Tensor, y: Tensor, threshold, batch_size: int, n_steps: int, start_step: int, num_steps: int, M: int,
                       ac: Tensor):
    """DiffusionClassConditionalGuidedPosteriorDistribution.__call__ with X_T / y given (:117-188).  Returns (uint8 images, x_0 of the
    last batch).  alpha_hat_t = alphas_cumprod[i] is indexed by the STEP number (:151); window test `start + num >= i >= start` (:153)."""
    timesteps = O.leading_timesteps(1000, n_steps).tolist()
    one = torch.tensor(1.0)
    images, last = [], None
    with torch.no_grad():
        for a in range(0, x_T.shape[0], batch_size):
            x, yb = x_T[a:a + batch_size], y[a:a + batch_size]
            for i, t in enumerate(timesteps):
                t_tensor = torch.full((yb.shape[0],), t, dtype=torch.long)
                predict = lambda z: model(z, t_tensor, y=yb)[:, :3]                                   # noqa: E731
                ddim = lambda e, s, t=t: plain_ddim_step(ac, one, n_steps, e, t, s)                   # noqa: E731
                eps = predict(x)
                if (start_step + num_steps) >= i >= start_step:
                    x = posterior_window_step(predict, x, eps, ac[i], M, threshold, i, ddim)
                else:
                    x = ddim(eps, x)
            images.append(image_uint8(x))
            last = x
    return torch.cat(images, dim=0), last


def second_order_pipeline(model, x_T: Tensor, y: Tensor, threshold, batch_size: int, n_steps: int, start_step: int, num_steps: int, M: int,
                          ac: Tensor, threshold_type: str = "higher"):
    """DiffusionClassConditionalGuidedSecondOrder.__call__ (PU/…second_order.py:110-194) with update_with_uncertainty (:196-258)
    and estimate_score_update (:283-306): centred second moment over M re-noised forwards, threshold map,
    eps + u * sign(randn_like(eps)) * mask (:249), scheduler step.  Window test `start + num > i >= start` (:158)."""
    timesteps = O.leading_timesteps(1000, n_steps).tolist()
    one = torch.tensor(1.0)
    images, last = [], None
    with torch.no_grad():
        for a in range(0, x_T.shape[0], batch_size):
            x, yb = x_T[a:a + batch_size], y[a:a + batch_size]
            for i, t in enumerate(timesteps):
                t_tensor = torch.full((yb.shape[0],), t, dtype=torch.long)
                predict = lambda z: model(z, t_tensor, y=yb)[:, :3]                                   # noqa: E731
                eps = predict(x)
                prev = plain_ddim_step(ac, one, n_steps, eps, t, x)
                if (start_step + num_steps) > i >= start_step:
                    preds = perturbed_predictions(predict, x, eps, prev, ac[i], M)
                    u = (torch.stack(preds, dim=0) - eps.unsqueeze(0)).pow(2).mean(dim=0)
                    mask = O.calculate_threshold_map(threshold, i, u, threshold_type)
                    guided = eps + u * torch.sign(torch.randn_like(eps)) * mask
                    prev = plain_ddim_step(ac, one, n_steps, guided, t, x)
                x = prev
            images.append(image_uint8(x))
            last = x
    return torch.cat(images, dim=0), last


def ema_momentum(momentum: Optional[Tensor], u: Tensor, i: int, beta: float = 0.99):
    """PU/…second_order.py:212-218: (momentum', corrected, sqrt)."""
    momentum = u if momentum is None else beta * momentum + (1 - beta) * u
    corrected = momentum / (1 - beta ** i + 1e-5)
    return momentum, corrected, torch.sqrt(corrected)


def legacy_mul_blend(eps: Tensor, mask: Tensor, g: Tensor) -> Tensor:
    """generate_samples.py:953."""
    return eps * (1 - mask) + eps * mask * g
