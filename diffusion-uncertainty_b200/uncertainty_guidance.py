"""Percentile uncertainty guidance of a predicted score — drop-in for diffusion_uncertainty/uncertainty_guidance.py.

`get_uncertainty_guided_score_with_percentile` is what the Stable Diffusion / SD3 / Flux pipelines call inside their
denoising loop (pipeline_uncertainty/pipeline_stable_diffusion_uncertainty_guided.py:764-772).  Per call:
  F7  M re-noised model inputs  x_hat = sqrt(abar) x0 + sqrt(1-abar) n            (reference :85-88)   du_ddim_step + du_perturb
  F1c unbiased variance over the M predictions + the original eps               (:99-104)            du_moments
  F2a per-image percentile threshold, strict `>` mask                            (:109-110)           du_quantile_threshold
  F5  posterior score blend  eps(1-m) + m [1/(M/u + 1/abar)] (1/u) eps.sum(0)   (:112-117)           du_guided_step
or, with `use_posterior = False`, the gradient form  eps + lr * d(uncertainty)/d(eps) * mask (:118-127): the scalar
objective and its backward pass through the score model stay in torch autograd, the mask and
the blend run in the kernels.

The model forwards and every `torch.randn_like` draw stay in torch, in the reference's order.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops

use_posterior = True   # module-level switch, as in the reference (:8)


# ---- model call conventions (reference :134-182) ----------------------------------------------------------------------
def predict_model_stable_diffusion(model, sample, t_tensor, y, guidance_scale: float, extra_diffusion_kwargs=None):
    kw = dict(extra_diffusion_kwargs or {})
    pred_noise = model(sample=sample, timestep=t_tensor, encoder_hidden_states=y, **kw)[0]
    uncond, text = pred_noise.chunk(2)
    return uncond + guidance_scale * (text - uncond)


def predict_model_stable_diffusion_3(model, sample, t_tensor, y, guidance_scale: float, extra_diffusion_kwargs=None):
    kw = dict(extra_diffusion_kwargs or {})
    kw["return_dict"] = False
    pred_noise = model(hidden_states=sample, timestep=t_tensor, encoder_hidden_states=y, **kw)[0]
    uncond, text = pred_noise.chunk(2)
    return uncond + guidance_scale * (text - uncond)


def predict_model_flux(model, sample, t_tensor, extra_diffusion_kwargs=None):
    kw = dict(extra_diffusion_kwargs or {})
    kw["return_dict"] = False
    return model(hidden_states=sample, timestep=t_tensor, **kw)[0]


def predict_model(model, sample, t_tensor, y, extra_diffusion_kwargs=None):
    return model(sample, t_tensor, y=y, **dict(extra_diffusion_kwargs or {}))[:, :3]


def _forward(model, model_type, x_hat, t_tensor, y, guidance_scale, extra):
    if model_type == "unet":
        # the reference assigns this to the wrong name and raises UnboundLocalError (:89-90); the
        # evident intent is implemented
        return predict_model(model, x_hat, t_tensor, y, extra_diffusion_kwargs=extra), t_tensor
    if model_type == "stable-diffusion-3":
        t_tensor = t_tensor.reshape((-1,))
        return predict_model_stable_diffusion_3(model, x_hat, t_tensor, y, guidance_scale, extra_diffusion_kwargs=extra), t_tensor
    if model_type == "flux":
        # the reference divides t by 1000 on EVERY iteration (:94) and passes one argument too many (:96)
        t_tensor = t_tensor.reshape((-1,)) / 1000.
        return predict_model_flux(model, x_hat, t_tensor, extra), t_tensor
    assert guidance_scale is not None
    return predict_model_stable_diffusion(model, x_hat, t_tensor, y, guidance_scale, extra_diffusion_kwargs=extra), t_tensor


def _as_rows(eps: Tensor, like: Tensor) -> Tensor:
    """the centre in the dtype of the predictions (the fused kernel reads scores and centre through one vector type)"""
    return eps if eps.dtype == like.dtype else eps.to(like.dtype)


def _rows_like(t: Tensor, ref: Tensor) -> Tensor:
    """Broadcast along the batch axis without copying (the CFG-doubled latent meets the single guided score)."""
    return t if t.shape == ref.shape else t.expand(ref.shape)


def get_uncertainty_guided_score_with_percentile(pred_epsilon: Tensor, input: Tensor, t_tensor: Tensor, y: Tensor, model,
                                                 alpha_hat_t, percentile: float, model_type: str,
                                                 num_uncertainty_samples: int = 5, guidance_scale: Optional[float] = None,
                                                 lr: float = 1.0, extra_diffusion_kwargs: Optional[dict] = None) -> Tensor:
    """Same arguments and result as the reference function (:61-131).  CUDA tensors only."""
    M = num_uncertainty_samples
    a = torch.as_tensor(alpha_hat_t)
    sa, sb = float(torch.sqrt(a)), float(torch.sqrt(1 - a))      # the reference's `torch.sqrt` on the 0-dim tensor
    c_x0 = ops.make_coeffs(sa, sb, 0.0, 0.0, clip_sample=False)

    if use_posterior:
        with torch.no_grad():
            # pred_x_0 = (input - sqrt(1-abar) eps) / sqrt(abar)
            x0 = ops.ddim_step(_rows_like(pred_epsilon, input), input, c_x0, want_prev=False, want_x0=True)[1]
            preds = []
            for _ in range(M):
                x_hat = ops.perturb_fresh(x0, sa, sb, noise_like=pred_epsilon)     # `torch.randn_like(pred_epsilon)` drawn in the kernel
                out, t_tensor = _forward(model, model_type, x_hat, t_tensor, y, guidance_scale, extra_diffusion_kwargs)
                preds.append(out)
            # F1c -> F2a -> F5 as ONE launch (du_fused_uncertainty_step with skip_ddim; the batch-axis sum `pred_epsilon.sum(dim=0)`,
            # :116, precedes it as its dependent launch when B > 1); shapes the fused kernel does not take run the three-kernel chain
            r = ops.uncertainty_step(preds, _as_rows(pred_epsilon, preds[0]), None, percentile, None, a, moments_mode="var_with_center",
                                     batch_sum=True)
            return r["eps"]

    # ---- gradient form: the objective and its backward stay in autograd
    pred_epsilon.requires_grad = True
    with torch.set_grad_enabled(True):
        pred_x_0 = (input - sb * pred_epsilon) / sa
        preds = []
        for _ in range(M):
            x_hat = sa * pred_x_0 + sb * torch.randn_like(pred_epsilon)
            out, t_tensor = _forward(model, model_type, x_hat, t_tensor, y, guidance_scale, extra_diffusion_kwargs)
            preds.append(out)
        u = ops.moments_autograd(preds, "var")          # torch.var(stack(preds), 0) and its backward: du_moments / du_moments_backward
        u.mean(dim=0).sum().backward()
    with torch.no_grad():
        u = u.detach()
        thr = ops.quantile_threshold(u, percentile)
        assert pred_epsilon.grad is not None
        r = ops.guided_step(pred_epsilon.detach(), None, None, guidance="grad_add", u=u, thr=thr, aux=pred_epsilon.grad,
                            lam=float(lr), want_eps=True)
    return r["eps"]
