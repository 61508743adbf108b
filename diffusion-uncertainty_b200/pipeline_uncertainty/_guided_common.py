"""Pieces shared by the guided class-conditional pipelines (posterior, second order, gradient): the DDIM scalars of a plain
scheduler, the batch bookkeeping of their `__call__` loops and the uint8 epilogue."""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops


def class_names(obj) -> set:
    return {c.__name__ for c in type(obj).__mro__}


def ddim_coeffs(scheduler, t: int, eta: float = 0.0):
    """DdimCoeffs of `scheduler.step(., t, .)` for the fused step: from the scheduler's own cache when it is one of this
    package's (UncertaintyDDIMCore._step_scalars), else from `alphas_cumprod` / `config` with the reference's fp32 tensor
    expressions (SU/scheduling_ddim_uncertainty_zigzag_centered.py:462-468, 294-302, 507) — a diffusers DDIMScheduler works too.
    Returns None when the scheduler's update is not the plain deterministic DDIM rule (the caller then calls step())."""
    if hasattr(scheduler, "_step_scalars"):
        if getattr(scheduler.config, "thresholding", False):
            return None
        return scheduler._step_scalars(int(t), eta, False)[0]
    cfg = scheduler.config
    if getattr(cfg, "thresholding", False) or not hasattr(scheduler, "alphas_cumprod"):
        return None
    prev_t = int(t) - cfg.num_train_timesteps // scheduler.num_inference_steps
    a_t = scheduler.alphas_cumprod[int(t)]
    a_prev = scheduler.alphas_cumprod[prev_t] if prev_t >= 0 else scheduler.final_alpha_cumprod
    return ops.make_coeffs(float(a_t ** 0.5), float((1 - a_t) ** 0.5), float(a_prev ** 0.5), float((1 - a_prev) ** 0.5),
                           clip_sample=bool(getattr(cfg, "clip_sample", False)),
                           clip_range=float(getattr(cfg, "clip_sample_range", 1.0)),
                           prediction_type=getattr(cfg, "prediction_type", "epsilon"))


def start_batch(X_T: Optional[torch.Tensor], y: Optional[torch.Tensor], num_classes, done: int, batch_size: int, channels: int,
                image_size: int, device, generator: torch.Generator, seed: int):
    """(input, y_slice) of one batch: a slice of the given tensors, or draws re-seeded per batch exactly as the reference
    (`generator.manual_seed(init_seed_rng + i_batch)` for BOTH draws, …guided_posterior_distribution.py:123-133)."""
    if X_T is not None:
        x = X_T[done:done + batch_size].to(device)
    else:
        x = torch.randn(batch_size, channels, image_size, image_size, device=device, dtype=torch.float32,
                        generator=generator.manual_seed(seed))
    if y is not None:
        y_slice = y[done:done + batch_size].to(device)
    else:
        y_slice = torch.randint(0, num_classes, (batch_size,), device=device, generator=generator.manual_seed(seed))
    return x, y_slice


def finish(results_y, results_x, scheduler, images, fid_evaluator):
    out = {'y': torch.cat(results_y, dim=0).cpu(), 'x_t': torch.cat(results_x, dim=0).cpu(), 'timestep': scheduler.timesteps,
           'gen_images': torch.cat(images, dim=0).cpu()}
    if fid_evaluator is not None:
        out['fid'] = fid_evaluator.compute()
    return out
