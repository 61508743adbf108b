"""Drop-in for diffusion_uncertainty/pipeline_uncertainty/pipeline_sampler_class_conditional_uncertainty_guided_gradient.py:
`DiffusionClassConditionalGuidedGradient` with the reference's constructor, `__call__` and `estimate_score_update`.

The loop shell is Python, as in the reference (:27-157).  On the uncertainty path:
  * F1a under autograd: `(stack(eps_hat) - eps[None]).pow(2).mean(0)` and its backward (:190-194) run in du_moments /
    du_moments_backward (ops.moments_autograd); the score model's own backward is torch autograd;
  * F2a / F2b: calculate_threshold_map (:100) — du_quantile_threshold + du_threshold_mask / du_tensor_threshold_mask;
  * F6: `eps (1 - m) + (eps + lambda g) m` (:114-118) — one du_guided_step launch (GRAD_BLEND);
  * N4: the uint8 image epilogue (:131-142) — du_image_uint8.
The reference's debug prints (:103-123, each a device->host sync) are dropped.
"""
from __future__ import annotations

from math import sqrt
from typing import Dict, Optional, Union

import torch

from .. import ops
from ..generate_samples import predict_model
from .threshold_guidance import calculate_threshold_map


def guided_gradient_blend(noisy_residual: torch.Tensor, update_scores: torch.Tensor, thresholded_map: torch.Tensor,
                          lambda_update: float) -> torch.Tensor:
    """:114-118  post = eps + lambda * g;  eps' = eps (1 - m) + post m   (du_guided_step, GRAD_BLEND, no DDIM part)."""
    return ops.guided_step(noisy_residual, None, None, guidance="grad_blend", mask=thresholded_map, aux=update_scores,
                           lam=float(lambda_update), want_eps=True)["eps"]


class DiffusionClassConditionalGuidedGradient:
    def __init__(self, model, scheduler, threshold: Union[torch.Tensor, float], image_size: int, device: torch.device, batch_size: int,
                 init_seed_rng: int, fid_evaluator: Optional[object] = None, M: int = 5, gradient_wrt: str = "input",
                 lambda_update: float = 0.1, threshold_type: str = "higher", gradient_direction: str = "descend"):
        self.model = model
        self.scheduler = scheduler
        self.threshold = threshold
        self.image_size = image_size
        self.device = device
        self.fid_evaluator = fid_evaluator
        self.batch_size = batch_size
        self.is_uvit = type(model).__name__ == "UViTAE"
        self.init_seed_rng = init_seed_rng
        self.M = M
        self.lambda_update = lambda_update
        self.gradient_wrt = gradient_wrt
        self.threshold_type = threshold_type
        self.gradient_direction = gradient_direction

    def __call__(self, num_samples: Optional[int] = None, num_classes: Optional[int] = None, X_T: Optional[torch.Tensor] = None,
                 y: Optional[torch.Tensor] = None, start_step: int = 0, num_steps: Optional[int] = None) -> Dict[str, torch.Tensor]:
        assert num_samples is not None or X_T is not None, "Either num_samples or X_T must be provided"
        assert num_classes is not None or y is not None, "Either num_classes or y must be provided"
        num_generated_samples = 0
        samples_x_t, samples_y, samples_gen_images = [], [], []
        if num_steps is None:
            num_steps = self.scheduler.timesteps.shape[0] - start_step
        if num_samples is None:
            num_samples = X_T.shape[0]
        if isinstance(self.threshold, torch.Tensor):
            assert self.threshold.shape[0] == self.scheduler.timesteps.shape[0], f'{self.threshold.shape=} {self.scheduler.timesteps.shape=}'
        else:
            assert isinstance(self.threshold, float), f'{self.threshold=}'
            assert self.threshold >= 0 and self.threshold <= 1, f'{self.threshold=}'
        self.scheduler.config.after_step = start_step
        self.scheduler.config.num_steps_uc = num_steps
        self.scheduler.set_timesteps(len(self.scheduler.timesteps))
        generator = torch.Generator(device=self.device)
        i_batch = 0
        while num_samples > num_generated_samples:
            if X_T is not None:
                input = X_T[num_generated_samples:num_generated_samples + self.batch_size].to(self.device)
            else:
                input = torch.randn(self.batch_size, 3, self.image_size, self.image_size, device=self.device, dtype=torch.float32,
                                    generator=generator.manual_seed(self.init_seed_rng + i_batch))
            samples_x_t.append(input.cpu().clone())
            if y is not None:
                y_slice = y[num_generated_samples:num_generated_samples + self.batch_size].to(self.device)
            else:
                y_slice = torch.randint(0, num_classes, (self.batch_size,), device=self.device,
                                        generator=generator.manual_seed(self.init_seed_rng + i_batch))
            samples_y.append(y_slice)
            self.scheduler.prompt_embeds = y_slice
            with torch.no_grad():
                for i, t in enumerate(self.scheduler.timesteps):
                    t = t.item()
                    t_tensor = torch.full((y_slice.shape[0],), t, device=self.device, dtype=torch.long)
                    noisy_residual = predict_model(self.model, input, t_tensor, y_slice)
                    output = self.scheduler.step(noisy_residual, t, input)
                    prev_noisy_sample = output.prev_sample
                    alpha_hat_t = self.scheduler.alphas_cumprod[i]      # indexed by STEP, as the reference (:95)
                    if (start_step + num_steps) > i >= start_step:
                        u, update_scores = self.estimate_score_update(input, y_slice, i, t_tensor, noisy_residual, prev_noisy_sample,
                                                                      alpha_hat_t)
                        thresholded_map = calculate_threshold_map(self.threshold, None, u, self.threshold_type)
                        noisy_residual = guided_gradient_blend(noisy_residual.detach(), update_scores, thresholded_map, self.lambda_update)
                        output = self.scheduler.step(noisy_residual, t, input.detach())
                        prev_noisy_sample = output.prev_sample
                    input = prev_noisy_sample
                if self.is_uvit:
                    input = self.model.decode(input)
            gen_images = ops.image_uint8(input)          # (x/2 + .5).clamp(0,1) * 255 -> round -> uint8  (:131-142)
            num_generated_samples += gen_images.shape[0]
            if self.fid_evaluator is not None:
                self.fid_evaluator.update(gen_images, real=False)
            samples_gen_images.append(gen_images)
            i_batch += 1
        results = {'y': torch.cat(samples_y, dim=0).cpu(), 'x_t': torch.cat(samples_x_t, dim=0).cpu(),
                   'timestep': self.scheduler.timesteps, 'gen_images': torch.cat(samples_gen_images, dim=0).cpu()}
        if self.fid_evaluator is not None:
            results['fid'] = self.fid_evaluator.compute()
        return results

    def estimate_score_update(self, input, y_slice, i, t_tensor, noisy_residual, prev_noisy_sample, alpha_hat_t):
        """(pixel_wise_uncertainty, update_scores) — :159-210.  The first backward (`pred_epsilon.mean(0).sum()`, :182-183)
        is part of the reference's gradient (it accumulates into the same .grad) and is reproduced."""
        a = float(alpha_hat_t)
        sa, sb = sqrt(a), sqrt(1 - a)
        noisy_residual = noisy_residual.detach().requires_grad_(self.gradient_wrt == 'score')
        input = input.detach().requires_grad_(self.gradient_wrt == 'input')
        with torch.enable_grad():
            pred_epsilon = predict_model(self.model, input, t_tensor, y_slice)
            if pred_epsilon.requires_grad:
                pred_epsilon.mean(dim=0).sum().backward()
            pred_x_0 = (input - sb * noisy_residual) / sa
            preds = []
            for _ in range(self.M):
                x_hat_t = sa * pred_x_0 + sb * torch.randn_like(prev_noisy_sample)
                preds.append(predict_model(self.model, x_hat_t, t_tensor, y_slice))
            u = ops.moments_autograd(preds, "centered", center=noisy_residual)
            u.mean(dim=0).sum().backward()
        update_scores = input.grad if self.gradient_wrt == 'input' else noisy_residual.grad
        return u.detach(), update_scores
