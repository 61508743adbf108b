"""Threshold maps and score updates of the class-conditional guided pipelines — drop-in for the module-level functions of
diffusion_uncertainty/pipeline_uncertainty/pipeline_sampler_class_conditional_uncertainty_guided_posterior_distribution.py
(`calculate_threshold_map` :10-30, `estimate_score_update_posterior` :32-68; duplicates in …_second_order.py:10-20).
"""
from __future__ import annotations

from math import sqrt
from typing import Optional, Union

import torch

from .. import ops


def calculate_threshold_map(threshold: Union[torch.Tensor, float], i: Optional[int], pixel_wise_uncertainty: torch.Tensor,
                            threshold_type: str) -> torch.Tensor:
    """float threshold: per-image percentile (F2a: du_quantile_threshold + du_threshold_mask, bit-identical to
    torch.quantile + compare); tensor threshold: `u > threshold[i]` (F2b: du_tensor_threshold_mask).  fp32 0/1 mask."""
    if not isinstance(threshold, (torch.Tensor, float)):
        raise TypeError(f"threshold must be a torch.Tensor or a float, got {type(threshold)}")   # the reference is @beartype'd
    if threshold_type not in ("higher", "lower"):
        raise TypeError(f"threshold_type must be 'higher' or 'lower', got {threshold_type!r}")
    if i is not None and not isinstance(i, int):
        raise TypeError(f"i must be an int or None, got {type(i)}")
    higher = threshold_type == "higher"
    u = pixel_wise_uncertainty
    if isinstance(threshold, float):
        if u.dim() < 2:
            raise IndexError("Dimension out of range: flatten(1) needs at least 2 dimensions")   # as u.flatten(1) would
        thr = ops.quantile_threshold(u, threshold)
        return ops.threshold_mask(u, thr, higher=higher)
    thr_i = threshold[i]
    if thr_i.device != u.device:
        thr_i = thr_i.to(u.device)
    if u.dim() == 4 and thr_i.numel() == u[0].numel():
        return ops.tensor_threshold_mask(u, thr_i, higher=higher)
    # general broadcast shapes (e.g. a [T] vector of scalar thresholds): materialise the row once
    row = thr_i.unsqueeze(0) if u.dim() == 4 else thr_i
    row = torch.broadcast_to(row, u.shape)[0].contiguous() if row.dim() == u.dim() else torch.broadcast_to(row, u.shape[1:]).contiguous()
    return ops.tensor_threshold_mask(u, row, higher=higher)


def estimate_score_update_posterior(M: int, model: torch.nn.Module, scheduler, input: torch.Tensor, y_slice: torch.Tensor,
                                    t_tensor: torch.Tensor, noisy_residual: torch.Tensor, prev_noisy_sample: torch.Tensor,
                                    alpha_hat_t, predict=None):
    """(pixel_wise_uncertainty, new_score): F7 -> M forwards -> F1c -> posterior score with the reference's batch-axis sum
    of the LAST perturbed prediction (…posterior_distribution.py:52-67)."""
    a = float(alpha_hat_t)
    sa, sb = sqrt(a), sqrt(1 - a)            # math.sqrt on the value, as the reference (:52-54)
    c_x0 = ops.make_coeffs(sa, sb, 0.0, 0.0, clip_sample=False)
    if predict is None:
        predict = lambda x, t, y: model(x, t, y=y)[:, :3]   # noqa: E731  (generate_samples.predict_model for ADM)
    with torch.no_grad():
        x0 = ops.ddim_step(noisy_residual, input, c_x0, want_prev=False, want_x0=True)[1]
        preds = []
        for _ in range(M):
            noise = torch.randn_like(prev_noisy_sample)
            preds.append(predict(ops.perturb(x0, noise, sa, sb), t_tensor, y_slice))
        u = ops.moments(preds, center=noisy_residual, mode="var_with_center", out_dtype=preds[0].dtype)
        S = ops.batch_sum(preds[-1])
        # post = [1/(M/u + 1/abar)] (1/u) S everywhere: the posterior blend with an all-ones mask
        ones = torch.ones_like(u, dtype=torch.float32)
        r = ops.guided_step(noisy_residual, None, None, guidance="posterior", u=u, mask=ones, aux=S, aux_broadcast=True,
                            post_M=float(M), inv_alpha_hat=float(1 / torch.as_tensor(alpha_hat_t, dtype=torch.float32)), want_eps=True)
    return u, r["eps"]
