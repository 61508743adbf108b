"""Drop-in for diffusion_uncertainty/pipeline_uncertainty/pipeline_sampler_class_conditional_uncertainty_guided_second_order.py:
`calculate_threshold_map` (:10-20), the pipeline class `DiffusionClassConditionalGuidedSecondOrder` (:71-330) and its pieces.

Per window step (:158-160, 196-258, 283-306): M re-noised forwards (F7), the centred second moment about the original score
(F1a, du_moments CENTERED), the threshold map (F2a / F2b), the second-order momentum of the map (an EMA with bias correction
and square root that the reference computes and only prints: du_ema_update), the update `eps + u * sign(randn_like(eps)) * mask`
(:249 — the draw comes from the Philox kernel, sign / product / mask / add are one du_guided_step launch, SIGN_ADD) and the
scheduler step on the guided score.  The ~30 debug prints per step (each a device->host sync) are dropped.
"""
from __future__ import annotations

from math import sqrt
from typing import Dict, Optional, Union

import torch

from .. import ops
from ..generate_samples import predict_model
from ._guided_common import class_names, ddim_coeffs, finish, start_batch
from .threshold_guidance import calculate_threshold_map, estimate_score_update_posterior  # noqa: F401


def second_order_blend(noisy_residual: torch.Tensor, pixel_wise_uncertainty: torch.Tensor, thresholded_map: torch.Tensor,
                       sample: Optional[torch.Tensor] = None, coeffs=None):
    """:249  eps + u * sign(randn_like(eps)) * mask.  The normal draw is torch's own stream, produced by du_perturb_randn's
    generator path when it can be (ops.randn_like: same bits, same generator offset), else by torch; everything after it is one
    du_guided_step launch.  With `sample` and `coeffs` the DDIM update of :257 runs in the same launch and the result is a dict
    (eps, prev); otherwise the guided score alone is returned."""
    like = noisy_residual
    if ops.randn_fusable(like):
        n = ops.randn_like(like)
    else:
        n = torch.randn_like(like)
    r = ops.guided_step(noisy_residual, sample, coeffs, guidance="sign_add", u=pixel_wise_uncertainty, mask=thresholded_map, aux=n,
                        want_prev=sample is not None, want_eps=True)
    return r if sample is not None else r["eps"]


def second_order_momentum_update(second_order_momentum, pixel_wise_uncertainty: torch.Tensor, i: int, momentum_beta: float = 0.99):
    """:212-218 — (momentum', momentum' / (1 - beta**i + 1e-5), sqrt of that), one du_ema_update launch."""
    return ops.ema_update(second_order_momentum, pixel_wise_uncertainty, momentum_beta, i)


class DiffusionClassConditionalGuidedSecondOrder:

    def __init__(self, model, scheduler, threshold: Union[torch.Tensor, float], image_size: int, device: torch.device, batch_size: int,
                 init_seed_rng: int, fid_evaluator: Optional[object] = None, M: int = 5, threshold_type: str = "higher"):
        assert isinstance(threshold, (torch.Tensor, float)), "Threshold must be a tensor or a float"
        if isinstance(threshold, float):
            assert 0 <= threshold <= 1, "Threshold percentile must be between 0 and 1"
        self.model = model
        self.scheduler = scheduler
        self.threshold = threshold
        self.image_size = image_size
        self.device = torch.device(device)
        self.fid_evaluator = fid_evaluator
        self.batch_size = batch_size
        self.is_uvit = "UViTAE" in class_names(model)
        self.init_seed_rng = init_seed_rng
        self.M = M
        self.lambda_update = 7
        self.threshold_type = threshold_type
        self.track_momentum = True     # the reference computes the momentum and only prints it; set False to skip the launch

    def __call__(self, num_samples: Optional[int] = None, num_classes: Optional[int] = None, X_T: Optional[torch.Tensor] = None,
                 y: Optional[torch.Tensor] = None, start_step: int = 0, num_steps: Optional[int] = None) -> Dict[str, torch.Tensor]:
        assert num_samples is not None or X_T is not None, "Either num_samples or X_T must be provided"
        assert num_classes is not None or y is not None, "Either num_classes or y must be provided"
        if self.device.type != "cuda":
            raise RuntimeError(f"device {self.device}: the uncertainty path has no CPU fallback")
        sched = self.scheduler
        num_generated_samples = 0
        samples_x_t, samples_y, samples_gen_images = [], [], []
        if num_steps is None:
            num_steps = sched.timesteps.shape[0] - start_step
        if num_samples is None:
            num_samples = X_T.shape[0]
        if isinstance(self.threshold, torch.Tensor):
            assert self.threshold.shape[0] == sched.timesteps.shape[0], f'{self.threshold.shape=} {sched.timesteps.shape=}'
        sched.config.after_step = start_step
        sched.config.num_steps_uc = num_steps
        sched.set_timesteps(len(sched.timesteps))
        generator = torch.Generator(device=self.device)
        i_batch = 0
        while num_samples > num_generated_samples:
            input, y_slice = start_batch(X_T, y, num_classes, num_generated_samples, self.batch_size, 4 if self.is_uvit else 3,
                                         self.image_size, self.device, generator, self.init_seed_rng + i_batch)
            samples_x_t.append(input.cpu().clone())
            samples_y.append(y_slice)
            sched.prompt_embeds = y_slice
            momentum_beta = 0.99
            second_order_momentum = torch.zeros_like(input)          # (:146 — never None, so the EMA starts from zero)
            with torch.no_grad():
                for i, t in enumerate(sched.timesteps.tolist()):
                    t_tensor = torch.full((y_slice.shape[0],), t, device=self.device, dtype=torch.long)
                    noisy_residual = predict_model(self.model, input, t_tensor, y_slice)
                    output = sched.step(noisy_residual, t, input)
                    prev_noisy_sample = output.prev_sample
                    alpha_hat_t = sched.alphas_cumprod[i]
                    if (start_step + num_steps) > i >= start_step:
                        prev_noisy_sample, second_order_momentum = self.update_with_uncertainty(
                            input, y_slice, momentum_beta, second_order_momentum, i, t, t_tensor, noisy_residual, prev_noisy_sample, alpha_hat_t)
                    input = prev_noisy_sample
                gen_images = self.model.decode(input) if self.is_uvit else input
                gen_images = ops.image_uint8(gen_images)
            num_generated_samples += gen_images.shape[0]
            if self.fid_evaluator is not None:
                self.fid_evaluator.update(gen_images, real=False)
            samples_gen_images.append(gen_images)
            i_batch += 1
        return finish(samples_y, samples_x_t, sched, samples_gen_images, self.fid_evaluator)

    def update_with_uncertainty(self, input, y_slice, momentum_beta, second_order_momentum, i, t, t_tensor, noisy_residual,
                                prev_noisy_sample, alpha_hat_t):
        """(x_{t-1}, momentum') — :196-258."""
        u = self.estimate_score_update(input, y_slice, i, t_tensor, noisy_residual, prev_noisy_sample, alpha_hat_t)
        thresholded_map = calculate_threshold_map(self.threshold, i, u, self.threshold_type)
        if self.track_momentum:
            second_order_momentum = second_order_momentum_update(second_order_momentum, u, i, momentum_beta)[0]
        coeffs = ddim_coeffs(self.scheduler, t) if type(self.scheduler).__name__ == "DDIMScheduler" else None
        if coeffs is not None:
            return second_order_blend(noisy_residual, u, thresholded_map, sample=input, coeffs=coeffs)["prev"], second_order_momentum
        guided = second_order_blend(noisy_residual, u, thresholded_map)
        return self.scheduler.step(guided, t, input).prev_sample, second_order_momentum

    def calculate_threshold_map(self, i, pixel_wise_uncertainty):
        return calculate_threshold_map(self.threshold, i, pixel_wise_uncertainty, "higher")

    def estimate_score_update(self, input, y_slice, i, t_tensor, noisy_residual, prev_noisy_sample, alpha_hat_t):
        """pixel_wise_uncertainty = mean_m (eps_hat_m - eps)^2 over M re-noised forwards — :283-306."""
        sa, sb = sqrt(alpha_hat_t), sqrt(1 - alpha_hat_t)
        c_x0 = ops.make_coeffs(sa, sb, 0.0, 0.0, clip_sample=False)
        x0 = ops.ddim_step(noisy_residual, input, c_x0, want_prev=False, want_x0=True)[1]
        preds = [predict_model(self.model, ops.perturb_fresh(x0, sa, sb, noise_like=prev_noisy_sample), t_tensor, y_slice)
                 for _ in range(self.M)]
        return ops.moments(preds, center=noisy_residual, mode="centered", out_dtype=torch.float32)
