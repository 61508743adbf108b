"""The threshold-guided sampling loops that inline the uncertainty path — drop-ins for

  generate_samples_model_scheduler_class_conditioned_with_threshold     pipeline_uncertainty/uncertainty_guidance.py:12-125
  generate_samples_uvit_scheduler_class_conditioned_with_threshold      generate_samples.py:721-860   (BASELINE config 4)
  generate_samples_model_scheduler_class_conditioned_with_percentile    generate_samples.py:861-983   (legacy)

One loop body serves the three; what differs is the map that is differentiated (centred second moment / variance / standard
deviation over 5 re-noised forwards), which map is compared with what (the loop's own map or the scheduler's
`output.uncertainty`, a fitted tensor threshold or a per-image percentile) and the blend.  On the uncertainty path:
  F7   x0 = (x - sqrt(1-a) eps) / sqrt(a) and x_hat = sqrt(a) x0 + sqrt(1-a) n stay on the autograd graph (the gradient with
       respect to eps flows through them): du_ddim_step / du_perturb_randn as differentiable ops (ops.x0_autograd, ops.perturb_fresh_autograd);
  F1   the reduction over the M predictions and its backward: du_moments / du_moments_backward (ops.moments_autograd);
  F2b  `u > threshold[i]`: du_tensor_threshold_mask;   F2a  per-image percentile: du_quantile_threshold + du_threshold_mask;
  F6   `eps + mask * (-g)` (and the legacy `eps (1-m) + eps m g`) fused with the scheduler's DDIM update where the scheduler is a
       plain one: du_guided_step;   N4 the uint8 epilogue: du_image_uint8.
The score model's forward and backward stay torch (north_star).  Reference quirks kept: `alpha_hat_t = alphas_cumprod[i]` indexed
by the STEP number, the window tests (`>=` against `start + num` in two of the loops, `>` in the U-ViT one), five perturbed
forwards whatever the scheduler's M, both re-seeded draws of a batch sharing one seed.
"""
from __future__ import annotations

from math import sqrt
from typing import Optional, Union

import torch

from . import ops
from .generate_samples import predict_model
from .schedulers_uncertainty.mixin import SchedulerUncertaintyMixin

N_PERTURBED = 5    # `for _ in range(5)` in all three loops


def _guided_gradient(model, input, y_slice, t_tensor, noisy_residual, like, alpha_hat_t, mode: str):
    """(map, d(map.mean(0).sum()) / d eps): the autograd block of the three loops (e.g. generate_samples.py:800-817)."""
    sa, sb = sqrt(alpha_hat_t), sqrt(1 - alpha_hat_t)
    eps = noisy_residual.detach().requires_grad_(True)
    with torch.enable_grad():
        x0 = ops.x0_autograd(eps, input.detach(), sa, sb)
        preds = []
        for _ in range(N_PERTURBED):
            x_hat = ops.perturb_fresh_autograd(x0, sa, sb, noise_like=like)     # n = torch.randn_like(prev_noisy_sample), drawn in the kernel
            preds.append(predict_model(model, x_hat, t_tensor, y_slice))
        if mode == "centered":
            u = ops.moments_autograd(preds, "centered", center=eps)
        else:
            u = ops.moments_autograd(preds, mode)
        u.mean(dim=0).sum().backward()
    return u.detach(), eps.grad


def _loop(num_samples, batch_size, channels, image_size, model, scheduler, num_classes, device, fid_evaluator, x_T, y, start_step,
          num_steps, seed, skip_seed, *, map_mode: str, window, mask_fn, blend: str, model_call, decode: bool, always_configure: bool):
    device = torch.device(device if device is not None else "cpu")
    if device.type != "cuda":
        raise RuntimeError(f"device {device}: the uncertainty path has no CPU fallback")
    is_unc = isinstance(scheduler, SchedulerUncertaintyMixin)
    if num_steps is None:
        num_steps = scheduler.timesteps.shape[0]
    if always_configure or is_unc:
        scheduler.config.after_step = start_step
        scheduler.config.num_steps_uc = num_steps
        scheduler.set_timesteps(len(scheduler.timesteps))
    generator = torch.Generator(device=device)
    samples_x_t, samples_y, images = [], [], []
    done, i_batch = 0, 0
    while num_samples > done:
        if x_T is not None:
            input = x_T[done:done + batch_size].to(device)
        else:
            input = torch.randn(batch_size, channels, image_size, image_size, device=device, dtype=torch.float32,
                                generator=generator.manual_seed(seed + i_batch * skip_seed))
        samples_x_t.append(input.cpu().clone())
        if y is not None:
            y_slice = y[done:done + batch_size].to(device)
        elif isinstance(num_classes, int):
            y_slice = torch.randint(0, num_classes, (batch_size,), device=device, generator=generator.manual_seed(seed + i_batch * skip_seed))
        else:
            assert num_samples == num_classes.shape[0]
            y_slice = num_classes[done:done + batch_size]
            if y_slice.shape[0] < batch_size:
                input = input[:y_slice.shape[0]]
        samples_y.append(y_slice)
        if is_unc:
            scheduler.prompt_embeds = y_slice
        with torch.no_grad():
            for i, t in enumerate(scheduler.timesteps.tolist()):
                t_tensor = torch.full((y_slice.shape[0],), t, device=device, dtype=torch.long)
                noisy_residual = model_call(model, input, t_tensor, y_slice)
                output = scheduler.step(noisy_residual, t, input)
                prev = output.prev_sample
                if is_unc and window(scheduler, i, t, start_step, num_steps):
                    alpha_hat_t = scheduler.alphas_cumprod[i]
                    u, grad = _guided_gradient(model, input, y_slice, t_tensor, noisy_residual, prev, alpha_hat_t, map_mode)
                    mask = mask_fn(u, output, i)
                    if blend == "add_neg":       # eps + mask * (grad * -1)
                        guided = ops.guided_step(noisy_residual, None, None, guidance="grad_add", mask=mask, aux=grad, lam=-1.0,
                                                 want_eps=True)["eps"]
                    else:                        # eps (1 - m) + eps m g   ==  eps (1-m) + (eps g) m: GRAD_BLEND's form with post = eps*g
                        guided = ops.guided_step(noisy_residual, None, None, guidance="mul_blend", mask=mask, aux=grad, want_eps=True)["eps"]
                    prev = scheduler.step(guided, t, input).prev_sample
                input = prev
        gen = model.decode(input) if decode else input
        gen = ops.image_uint8(gen)
        done += gen.shape[0]
        if fid_evaluator is not None:
            fid_evaluator.update(gen, real=False)
        images.append(gen)
        i_batch += 1
    results = {'y': torch.cat(samples_y, dim=0).cpu(), 'x_t': torch.cat(samples_x_t, dim=0).cpu(), 'timestep': scheduler.timesteps,
               'gen_images': torch.cat(images, dim=0).cpu()}
    if fid_evaluator is not None:
        results['fid'] = fid_evaluator.compute()
    return results


@torch.no_grad()
def generate_samples_model_scheduler_class_conditioned_with_threshold(num_samples, batch_size, image_size, model, scheduler,
                                                                      num_classes: Union[int, torch.Tensor], threshold: torch.Tensor,
                                                                      device=None, fid_evaluator=None, x_T=None, y=None,
                                                                      start_step: int = 0, num_steps: Optional[int] = None,
                                                                      seed: int = 0, is_cifar10: bool = False):
    """pipeline_uncertainty/uncertainty_guidance.py:12-125: ADM (or the CIFAR-10 UNet2DModel) loop; map = centred second moment
    about eps (:97-99); mask = own map > threshold[i] (:106-107); eps' = eps + mask * (-grad) (:112); window `start+num >= i >= start`."""
    assert threshold.shape[0] == scheduler.timesteps.shape[0], f'{threshold.shape=} {scheduler.timesteps.shape=}'

    def call(m, x, tt, yy):
        return m(x, tt).sample if is_cifar10 else m(x, tt, y=yy)[:, :3]

    return _loop(num_samples, batch_size, 3, image_size, model, scheduler, num_classes, device, fid_evaluator, x_T, y, start_step, num_steps,
                 seed, 1, map_mode="centered", window=lambda s, i, t, a, n: (a + n) >= i >= a,
                 mask_fn=lambda u, out, i: ops.tensor_threshold_mask(u, threshold[i].to(u.device), higher=True),
                 blend="add_neg", model_call=call, decode=False, always_configure=True)


@torch.no_grad()
def generate_samples_uvit_scheduler_class_conditioned_with_threshold(num_samples, batch_size, image_size, model, scheduler,
                                                                     num_classes: Union[int, torch.Tensor], threshold: torch.Tensor,
                                                                     device=None, fid_evaluator=None, x_T: Optional[torch.Tensor] = None,
                                                                     y: Optional[torch.Tensor] = None, start_step: int = 0,
                                                                     num_steps: Optional[int] = None, seed: int = 0, skip_seed: int = 1):
    """generate_samples.py:721-860 (BASELINE config 4): U-ViT latent loop; map = torch.var over the 5 predictions (:815); mask =
    the SCHEDULER's map `output.uncertainty > threshold[i]` (:819-820); eps' = eps + mask * (-grad) (:826); window `start+num > i >= start`;
    `model.decode` before the uint8 epilogue."""
    assert threshold.shape[0] == scheduler.timesteps.shape[0], f'{threshold.shape=} {scheduler.timesteps.shape=}'
    return _loop(num_samples, batch_size, 4, image_size, model, scheduler, num_classes, device, fid_evaluator, x_T, y, start_step, num_steps,
                 seed, skip_seed, map_mode="var", window=lambda s, i, t, a, n: (a + n) > i >= a,
                 mask_fn=lambda u, out, i: ops.tensor_threshold_mask(out.uncertainty, threshold[i].to(u.device), higher=True),
                 blend="add_neg", model_call=lambda m, x, tt, yy: m(x, tt, yy), decode=True, always_configure=False)


@torch.no_grad()
def generate_samples_model_scheduler_class_conditioned_with_percentile(num_samples, batch_size, image_size, model, scheduler,
                                                                       num_classes: Union[int, torch.Tensor], percentile: float,
                                                                       device=None, fid_evaluator=None, x_T=None, start_step: int = 0,
                                                                       num_steps: Optional[int] = None, seed: int = 0):
    """generate_samples.py:861-983 (legacy): map = torch.std over the 5 predictions (:941); the per-image percentile mask of that
    map (:945-946) is then used AS THE THRESHOLD of the scheduler's map — `output.uncertainty > mask` with the 0/1 mask promoted
    to float (:948) — and the blend is eps (1-m) + eps m g (:953).  Window: the scheduler's own window AND `start+num >= i >= start`."""

    def mask_fn(u, out, i):
        pm = ops.threshold_mask(u, ops.quantile_threshold(u, percentile), higher=True)
        return ops.mask_greater(out.uncertainty, pm)

    return _loop(num_samples, batch_size, 3, image_size, model, scheduler, num_classes, device, fid_evaluator, x_T, None, start_step, num_steps,
                 seed, 1, map_mode="std",
                 window=lambda s, i, t, a, n: s.timestep_after_step >= t >= s.timestep_end_step and (a + n) >= i >= a,
                 mask_fn=mask_fn, blend="mul", model_call=lambda m, x, tt, yy: m(x, tt, y=yy)[:, :3], decode=False, always_configure=False)
