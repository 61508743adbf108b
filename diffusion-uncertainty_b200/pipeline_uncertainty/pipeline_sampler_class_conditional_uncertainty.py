"""Drop-in for diffusion_uncertainty/pipeline_uncertainty/pipeline_sampler_class_conditional_uncertainty.py:
`DiffusionClassConditionalWithUncertainty` — the sampling pipeline whose scheduler returns an uncertainty map per in-window
step (reference :9-147; constructor :11-21, `generate` :25-63, `sample` :65-75, `__call__` :78-147, batch helpers :149-190,
`predict_score` :192-210).

What changes is where the per-step outputs go: the reference appends `output.uncertainty.cpu()` and
`output.pred_epsilon.cpu()` every in-window step (:49-53, synchronous pageable copies), stacks them per batch (:62-63) and
concatenates over batches (:146-147).  Here the scheduler's moments kernel writes each map into slot `[:, k]` of a device
`[B, T_uc, C, H, W]` buffer (UncertaintyMapAccumulator), the score follows with one du_accumulate_slot launch, and each batch
leaves the GPU as one asynchronous copy into its slice of the pinned result.  The uint8 epilogue (:57-60) is du_image_uint8.
"""
from __future__ import annotations

from functools import singledispatchmethod
from typing import Dict, List, Optional, Tuple

import torch

from .. import ops
from ..accumulate import UncertaintyMapAccumulator


def _class_names(obj) -> set:
    return {c.__name__ for c in type(obj).__mro__}


class DiffusionClassConditionalWithUncertainty:

    def __init__(self, model, scheduler, image_size: int, device: torch.device, batch_size: int, init_seed_rng: int,
                 fid_evaluator: Optional[object] = None, return_intermediates: bool = False):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(f"device {device}: the uncertainty path has no CPU fallback")
        self.model = model.to(device)
        self.device = device
        self.image_size = image_size
        self.fid_evaluator = fid_evaluator
        self.batch_size = batch_size
        names = _class_names(model)
        self.is_uvit = "UViTAE" in names          # matched by class name: neither the U-ViT package nor diffusers is needed
        self.is_cifar10 = "UNet2DModel" in names
        self.init_seed_rng = init_seed_rng
        self.scheduler = scheduler
        self.return_intermediates = return_intermediates

    # ------------------------------------------------------------------------------------------------ one batch
    def generate(self, X_T, y_batch) -> Tuple[torch.Tensor, ...]:
        """(gen_images uint8, uncertainties [B,T_uc,...], scores [B,T_uc,...][, intermediates]) — reference :25-63."""
        X_t: torch.Tensor = X_T.to(self.device)
        sched = self.scheduler
        sched.prompt_embeds = y_batch
        B = X_t.shape[0]
        t_uc = len(sched.uncertainty_timesteps()) if hasattr(sched, "uncertainty_timesteps") else sum(
            1 for t in sched.timesteps.tolist() if sched.timestep_after_step >= t >= sched.timestep_end_step)
        acc_u = UncertaintyMapAccumulator(B, max(t_uc, 1), X_t.shape[1:], self.device)
        acc_s = UncertaintyMapAccumulator(B, max(t_uc, 1), X_t.shape[1:], self.device)
        intermediates: List[torch.Tensor] = []
        if hasattr(sched, "attach_accumulator"):
            sched.attach_accumulator(acc_u)
        host_copies = getattr(sched, "host_copies", False)
        if host_copies:
            sched.host_copies = False      # MC-dropout's per-step .cpu() of x0 / score is not read by this pipeline
        try:
            with torch.no_grad():
                for t in sched.timesteps.tolist():
                    t_tensor = torch.full((B,), t, device=self.device, dtype=torch.long)
                    noisy_residual = self.predict_score(X_t, y_batch, t_tensor)
                    output = sched.step(noisy_residual, t, X_t)
                    if sched.timestep_after_step >= t >= sched.timestep_end_step:
                        if not getattr(sched, "map_in_sink", False):
                            acc_u.stash(output.uncertainty)
                        acc_s.stash(output.pred_epsilon)
                    if self.return_intermediates:
                        intermediates.append(output.intermediate)     # AttributeError for schedulers without it, as the reference
                    X_t = output.prev_sample
                if self.is_uvit:
                    X_t = self.model.decode(X_t)
                gen_images = ops.image_uint8(X_t)
        finally:
            if host_copies:
                sched.host_copies = True
            if hasattr(sched, "attach_accumulator"):
                sched.attach_accumulator(None)
        unc, ev_u = acc_u.to_host_async()
        sc, ev_s = acc_s.to_host_async()
        ev_u.synchronize()
        ev_s.synchronize()
        if self.return_intermediates:
            return gen_images, unc, sc, torch.stack(intermediates, dim=1).cpu()
        return gen_images, unc, sc

    @singledispatchmethod
    def sample(self, samples, classes):
        raise NotImplementedError

    @sample.register
    def _(self, samples: int, classes: int):
        return self(num_samples=samples, num_classes=classes)

    @sample.register
    def _(self, X_T: torch.Tensor, y: torch.Tensor):
        return self(X_T=X_T, y=y)

    # ------------------------------------------------------------------------------------------------ all batches
    def __call__(self, /, num_samples: Optional[int] = None, num_classes: Optional[int] = None, X_T: Optional[torch.Tensor] = None,
                 y: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        assert num_samples is not None or X_T is not None, "Either num_samples or X_T must be provided"
        assert num_classes is not None or y is not None, "Either num_classes or y must be provided"
        num_generated_samples = 0
        samples_unc, samples_scores, samples_X_t, samples_y, samples_gen, samples_inter = [], [], [], [], [], []
        i_batch = 0
        if num_samples is None:
            num_samples = X_T.shape[0]
        generator = torch.Generator(device=self.device)
        while num_samples > num_generated_samples:
            X_T_batch = self.get_X_T_batch(X_T, num_generated_samples, i_batch, generator)
            y_batch = self.get_y_batch(num_classes, y, num_generated_samples, i_batch, generator=generator)
            samples_X_t.append(X_T_batch.cpu().clone())
            y_batch = y_batch.to(self.device)
            samples_y.append(y_batch)
            out = self.generate(X_T=X_T_batch, y_batch=y_batch)
            gen_images, unc, sc = out[0], out[1], out[2]
            if self.return_intermediates:
                samples_inter.append(out[3])
            samples_unc.append(unc.clone())        # the accumulator's pinned staging tensor is reused by the next batch
            samples_scores.append(sc.clone())
            num_generated_samples += gen_images.shape[0]
            if self.fid_evaluator is not None:
                self.fid_evaluator.update(gen_images, real=False)
            samples_gen.append(gen_images)
            i_batch += 1
        results = {'y': torch.cat(samples_y, dim=0).cpu(), 'x_t': torch.cat(samples_X_t, dim=0).cpu(),
                   'timestep': self.scheduler.timesteps, 'gen_images': torch.cat(samples_gen, dim=0).cpu()}
        if self.return_intermediates:
            results['intermediates'] = torch.cat(samples_inter, dim=0).cpu()
        if self.fid_evaluator is not None:
            results['fid'] = self.fid_evaluator.compute()
        results['uncertainty'] = torch.cat(samples_unc, dim=0)
        results['score'] = torch.cat(samples_scores, dim=0)
        return results

    def get_y_batch(self, num_classes, y, num_generated_samples, i_batch, generator):
        if y is not None:
            return y[num_generated_samples:num_generated_samples + self.batch_size]
        return torch.randint(0, num_classes, (self.batch_size,), device=self.device,
                             generator=generator.manual_seed(self.init_seed_rng + i_batch))

    def get_X_T_batch(self, X_T, num_generated_samples, i_batch, generator):
        if X_T is not None:
            return X_T[num_generated_samples:num_generated_samples + self.batch_size]
        channels = 4 if self.is_uvit else 3
        return torch.randn(self.batch_size, channels, self.image_size, self.image_size, device=self.device, dtype=torch.float32,
                           generator=generator.manual_seed(self.init_seed_rng + i_batch))

    def predict_score(self, input: torch.Tensor, y_batch: torch.Tensor, t_tensor: torch.Tensor) -> torch.Tensor:
        if self.is_uvit:
            return self.model(input, t_tensor, y_batch)
        if self.is_cifar10:
            return self.model(input, t_tensor).sample
        return self.model(input, t_tensor, y=y_batch)[:, :3]
