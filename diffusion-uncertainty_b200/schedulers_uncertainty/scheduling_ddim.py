"""`DDIMScheduler` — the plain DDIM scheduler the guided pipelines are driven with
(scripts/generate_images_with_uncertainty_threshold.py:202-203 builds diffusers' `DDIMScheduler.from_config(scheduler.config)`;
its `step()` arithmetic is the F3 block every reference scheduler file copies, SU/scheduling_ddim_uncertainty_zigzag_centered.py:
461-525).  Same constructor, `set_timesteps`, `step`, `add_noise`, `get_velocity` as the uncertainty schedulers, no uncertainty
window, no per-step `best_noise` draw (diffusers draws `variance_noise` only when eta > 0).  x_{t-1} is one du_ddim_step launch.
"""
from ._core import UncertaintyDDIMCore


class DDIMScheduler(UncertaintyDDIMCore):
    draws_best_noise = False
    eta_uses_best_noise = False

    def set_timesteps(self, num_inference_steps: int, device=None):
        # the guided pipelines assign config.after_step / num_steps_uc before calling this (…guided_posterior_distribution.py:
        # 113-115); for a plain DDIM scheduler they are inert, whatever their values
        after, n_uc = self.config.after_step, self.config.num_steps_uc
        self.config.after_step, self.config.num_steps_uc = 0, 1
        try:
            super().set_timesteps(num_inference_steps, device)
        finally:
            self.config.after_step, self.config.num_steps_uc = after, n_uc
        self.timestep_after_step = None
        self.timestep_end_step = None

    def in_window(self, timestep: int) -> bool:
        return False

    def uncertainty_timesteps(self):
        return []
