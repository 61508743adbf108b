"""Module-name alias of the reference's pipeline_uncertainty/pipeline_sampler_class_conditional_uncertainty_guided_second_order.py
for the functions on the uncertainty path: `calculate_threshold_map` (:10-20), the perturbed-forward centred second moment
(:300-306, du_moments CENTERED) and the second-order blend `eps + u * sign(randn) * mask` (:249)."""
import torch

from .. import ops
from .threshold_guidance import calculate_threshold_map, estimate_score_update_posterior  # noqa: F401


def second_order_blend(noisy_residual: torch.Tensor, pixel_wise_uncertainty: torch.Tensor, thresholded_map: torch.Tensor) -> torch.Tensor:
    """:249  eps + u * sign(randn_like(eps)) * mask.  The sign draw stays torch (RNG order is part of parity); the blend is one
    du_guided_step launch (GRAD_ADD with lambda = 1: eps + (1*g)*m, g = u * sign)."""
    g = pixel_wise_uncertainty * torch.sign(torch.randn_like(noisy_residual))
    return ops.guided_step(noisy_residual, None, None, guidance="grad_add", mask=thresholded_map, aux=g, lam=1.0, want_eps=True)["eps"]


def second_order_momentum_update(second_order_momentum, pixel_wise_uncertainty: torch.Tensor, i: int, momentum_beta: float = 0.99):
    """:212-218 — the EMA of the map, its bias correction and square root (computed and only printed by the reference)."""
    if second_order_momentum is None:
        second_order_momentum = pixel_wise_uncertainty
    else:
        second_order_momentum = momentum_beta * second_order_momentum + (1 - momentum_beta) * pixel_wise_uncertainty
    corrected = second_order_momentum / (1 - momentum_beta ** (i) + 1e-5)
    return second_order_momentum, corrected, torch.sqrt(corrected)
