"""L4 sampling loop with device-side uncertainty-map accumulation — drop-in for
`generate_samples_model_scheduler_class_conditioned_from_tensor` (diffusion_uncertainty/generate_samples.py:125-231).

What changes relative to the reference loop is only where the per-step outputs go:
  reference  `uncertanties.append(output.uncertainty.cpu())` / `scores.append(output.pred_epsilon.cpu())` every in-window
             step (synchronous pageable D2H), `torch.stack(dim=1)` per batch, `torch.cat(dim=0)` over batches (:189-201, 229-231)
  here       the scheduler's moments kernel writes each map straight into slot [:, k] of a device `[B, T_uc, C, H, W]`
             buffer, the score goes there with one du_accumulate_slot launch, and each batch leaves the GPU as ONE async
             copy into its slice of a pinned `[N, T_uc, C, H, W]` host tensor — the final result, no stack / cat.
The model call, the timestep loop and the returned dict (`gen_images` uint8, `uncertainty`, `score`, `intermediates`, `fid`)
are the reference's.
"""
from __future__ import annotations

from typing import Any, Optional

import torch

from . import ops
from .accumulate import UncertaintyMapAccumulator
from .schedulers_uncertainty.mixin import SchedulerUncertaintyMixin


def predict_model(model, x, t_tensor, y):
    """Model dispatch of the reference loops (generate_samples.py:670-676): diffusers UNet2DModel -> `.sample`, UViTAE ->
    positional class label, anything else the ADM convention with the learned-sigma head dropped.  Matched by class name so
    that neither diffusers nor the U-ViT package has to be importable."""
    names = {c.__name__ for c in type(model).__mro__}
    if "UNet2DModel" in names:
        return model(x, t_tensor).sample
    if "UViTAE" in names:
        return model(x, t_tensor, y)
    return model(x, t_tensor, y=y)[:, :3]


@torch.no_grad()
def generate_samples_model_scheduler_class_conditioned_from_tensor(X_T: torch.Tensor, y: torch.Tensor, batch_size: int,
                                                                   device: torch.device, model: torch.nn.Module, scheduler,
                                                                   fid_evaluator: Any = None, save_intermediates: bool = False):
    assert X_T.shape[0] == y.shape[0], f"{X_T.shape=} {y.shape=}"
    return _sampling_loop(X_T, y, batch_size, device, model, scheduler, fid_evaluator, save_intermediates, conditioned=True)


@torch.no_grad()
def generate_samples_model_scheduler_unconditioned_from_tensor(X_T: torch.Tensor, batch_size: int, device: torch.device,
                                                               model: torch.nn.Module, scheduler, fid_evaluator: Any = None,
                                                               save_intermediates: bool = False):
    """Drop-in for the unconditioned (CIFAR-10 DDPM) loop, generate_samples.py:366-463: the model is called as
    `model(x, t).sample[:, :3]` (:414), no class labels, no `scale_model_input`; outputs and accumulation as above."""
    return _sampling_loop(X_T, None, batch_size, device, model, scheduler, fid_evaluator, save_intermediates, conditioned=False)


@torch.no_grad()
def generate_samples_model_scheduler_class_conditioned_uvit_from_tensor(X_T: torch.Tensor, y: torch.Tensor, batch_size: int,
                                                                        uvit_ae: torch.nn.Module, scheduler,
                                                                        device="cpu", fid_evaluator: Any = None):
    """Drop-in for the U-ViT latent loop, generate_samples.py:469-571: `uvit_ae(x, t, y)` on the `[B,4,h,w]` latent (no channel
    slice), `uvit_ae.decode(x)` before the uint8 epilogue, `timestep` in the result.  (The reference's default device 'cpu'
    is kept in the signature; this path has no CPU fallback and raises on it.)"""
    assert X_T.shape[0] == y.shape[0], f"{X_T.shape=} {y.shape=}"
    res = _sampling_loop(X_T, y, batch_size, device, uvit_ae, scheduler, fid_evaluator, False, conditioned=True, uvit=True)
    return {"timestep": scheduler.timesteps, **res}


@torch.no_grad()
def generate_samples_model_scheduler_class_conditioned(num_samples: int, batch_size: int, image_size: int, model: torch.nn.Module,
                                                       scheduler, num_classes, device=None, fid_evaluator: Any = None,
                                                       init_seed_rng: int = 0, is_uvit: bool = False, skip_seed: int = 1,
                                                       is_cifar10: bool = False):
    """Drop-in for the seed-driven loop, generate_samples.py:18-125 (scripts/compute_fid_imagenet.py:139): batch k starts from
    `torch.randn(batch_size, C, S, S, generator=Generator(device).manual_seed(init_seed_rng + k * skip_seed))` and, for an
    integer `num_classes`, labels drawn from the same re-seeded generator; a label TENSOR fixes the labels and truncates the
    last batch.  Whole batches are generated (an integer `num_classes` never truncates), `x_t` keeps the untruncated draws —
    both as in the reference.  Model call: `model(x, t, y=y)[:, :3]`, `model(x, t, y)` (is_uvit, + decode) or
    `model(x, t).sample` (is_cifar10).  Returns y, x_t, timestep, gen_images (+ uncertainty, score, fid)."""
    return _seeded_loop(num_samples, batch_size, 4 if is_uvit else 3, image_size, model, scheduler, num_classes, device, fid_evaluator,
                        init_seed_rng, skip_seed, is_uvit, is_cifar10)


@torch.no_grad()
def generate_samples_model_scheduler_class_conditioned_uvit(num_samples: int, batch_size: int, image_size: int, uvit_ae: torch.nn.Module,
                                                            scheduler, num_classes, device="cpu", fid_evaluator: Any = None,
                                                            init_seed_rng: int = 0):
    """Drop-in for the seed-driven U-ViT loop, generate_samples.py:573-668: the latent shape comes from `uvit_ae.in_chans` /
    `uvit_ae.img_size` (the `image_size` argument is unused there too), batch k is seeded with `init_seed_rng + k`."""
    return _seeded_loop(num_samples, batch_size, uvit_ae.in_chans, uvit_ae.img_size, uvit_ae, scheduler, num_classes, device,
                        fid_evaluator, init_seed_rng, 1, True, False)


def _seeded_loop(num_samples, batch_size, channels, image_size, model, scheduler, num_classes, device, fid_evaluator, init_seed_rng,
                 skip_seed, is_uvit, is_cifar10):
    device = torch.device(device if device is not None else "cpu")
    if device.type != "cuda":
        raise RuntimeError(f"device {device}: the uncertainty path has no CPU fallback")
    generator = torch.Generator(device=device)
    starts, labels, raw = [], [], []
    generated, k = 0, 0
    while num_samples > generated:
        seed = init_seed_rng + k * skip_seed
        x = torch.randn(batch_size, channels, image_size, image_size, device=device, dtype=torch.float32,
                        generator=generator.manual_seed(seed))
        raw.append(x.cpu().clone())
        if isinstance(num_classes, int):
            yb = torch.randint(0, num_classes, (batch_size,), device=device, generator=generator.manual_seed(seed))
        else:
            assert num_samples == num_classes.shape[0]
            yb = num_classes[generated:generated + batch_size]
            if yb.shape[0] < batch_size:
                x = x[:yb.shape[0]]
        starts.append(x)
        labels.append(yb)
        generated += x.shape[0]
        k += 1
    y_all = torch.cat(labels, dim=0)
    res = _sampling_loop(torch.cat(starts, dim=0), y_all, batch_size, device, model, scheduler, fid_evaluator, False,
                         conditioned=True, uvit=is_uvit, sample_attr=is_cifar10)
    return {"y": y_all.cpu(), "x_t": torch.cat(raw, dim=0), "timestep": scheduler.timesteps, **res}


def _sampling_loop(X_T, y, batch_size, device, model, scheduler, fid_evaluator, save_intermediates, conditioned: bool,
                   uvit: bool = False, sample_attr: bool = False):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"device {device}: the uncertainty path has no CPU fallback")
    num_samples = X_T.shape[0]
    with_unc = isinstance(scheduler, SchedulerUncertaintyMixin)
    n_steps = len(scheduler.timesteps)

    host_unc = host_score = None
    copies = []
    images, intermediates_all = [], []
    start = 0
    while start < num_samples:
        stop = min(start + batch_size, num_samples)
        x = X_T[start:stop].to(device)
        y_batch = y[start:stop].to(device) if conditioned else None
        B = stop - start
        scheduler.set_timesteps(n_steps)
        acc_u = acc_s = None
        if with_unc:
            if conditioned:
                scheduler.prompt_embeds = y_batch
            t_uc = len(scheduler.uncertainty_timesteps()) if hasattr(scheduler, "uncertainty_timesteps") else sum(
                1 for t in scheduler.timesteps.tolist() if scheduler.timestep_after_step >= t >= scheduler.timestep_end_step)
            if t_uc > 0:
                acc_u = UncertaintyMapAccumulator(B, t_uc, x.shape[1:], device)
                acc_s = UncertaintyMapAccumulator(B, t_uc, x.shape[1:], device)
                if hasattr(scheduler, "attach_accumulator"):
                    scheduler.attach_accumulator(acc_u)
        inter = []
        host_copies = getattr(scheduler, "host_copies", False)
        if host_copies:
            scheduler.host_copies = False      # MC-dropout's per-step .cpu() of x0 / score (mc_dropout.py:551-554): not in this loop
        try:
            for t in scheduler.timesteps.tolist():
                t_tensor = torch.full((B,), t, device=device, dtype=torch.long)
                if uvit:
                    noisy_residual = model(x, t_tensor, y_batch)
                elif sample_attr:
                    noisy_residual = model(x, t_tensor).sample            # (:71, no channel slice)
                elif conditioned:
                    x = scheduler.scale_model_input(x, t)
                    noisy_residual = predict_model(model, x, t_tensor, y_batch)
                else:
                    noisy_residual = model(x, t_tensor).sample[:, :3]
                output = scheduler.step(noisy_residual, t, x)
                if save_intermediates:
                    inter.append(output.prev_sample)
                if with_unc and scheduler.timestep_after_step >= t >= scheduler.timestep_end_step:
                    if not getattr(scheduler, "map_in_sink", False):     # the step did not write its map into the sink itself
                        acc_u.stash(output.uncertainty)                  # (other dtype / shape than the slots: converted here)
                    acc_s.stash(output.pred_epsilon)
                x = output.prev_sample
        finally:
            if host_copies:
                scheduler.host_copies = True
            if with_unc and hasattr(scheduler, "attach_accumulator"):
                scheduler.attach_accumulator(None)
        if acc_u is not None:
            if host_unc is None:
                shape = (num_samples, acc_u.num_slots) + tuple(x.shape[1:])
                host_unc = torch.empty(shape, dtype=acc_u.buffer.dtype, pin_memory=True)
                host_score = torch.empty(shape, dtype=acc_s.buffer.dtype, pin_memory=True)
            copies.append(acc_u.to_host_async(host_unc[start:stop])[1])
            copies.append(acc_s.to_host_async(host_score[start:stop])[1])
        if uvit:
            x = model.decode(x)               # latent -> pixels: the reference's autoencoder module (:539)
        gen = ops.image_uint8(x)          # (x/2 + .5).clamp(0,1)*255 -> round -> uint8 (:203-212), one launch
        if fid_evaluator is not None:
            fid_evaluator.update(gen, real=False)
        images.append(gen)
        if save_intermediates:
            intermediates_all.append(torch.stack(inter, dim=1))
        start = stop

    results = {"gen_images": torch.cat(images, dim=0).cpu()}
    if save_intermediates:
        results["intermediates"] = torch.cat(intermediates_all, dim=0).cpu()
    if fid_evaluator is not None:
        results["fid"] = fid_evaluator.compute()
    if host_unc is not None:
        for ev in copies:
            ev.synchronize()
        results["uncertainty"] = host_unc
        results["score"] = host_score
    return results


# the threshold-guided loops of the same reference module (generate_samples.py:721-983) live in guided_loops.py
from .guided_loops import (generate_samples_model_scheduler_class_conditioned_with_percentile,  # noqa: E402,F401
                           generate_samples_uvit_scheduler_class_conditioned_with_threshold)
