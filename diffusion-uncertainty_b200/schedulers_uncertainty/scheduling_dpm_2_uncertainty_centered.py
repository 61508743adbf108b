"""Drop-in for diffusion_uncertainty/schedulers_uncertainty/scheduling_dpm_2_uncertainty_centered.py (factory key
`dpm_2_uncertainty_centered`; the "ADM w/ 2-DPM" rows of the paper's table 1): a DPM-Solver++ multistep scheduler
(order 1 / 2, midpoint or heun) whose `step()` also estimates the centred second moment of M re-noised predictions
inside the timestep window (reference :955-974).

Same class names (`KDPM2DiscreteSchedulerUncertainty`, `KDPM2SchedulerUncertaintyImagenet`,
`KDPM2SchedulerUncertaintyImagenetClassConditioned`, `SchedulerUncertaintyOutput`), constructor arguments, `set_timesteps`,
`step(model_output, timestep, sample, generator=None, return_dict=True)` and outputs.  All tensor arithmetic runs in
libdu_b200.so: the x0 conversion (`du_ddim_step`), the solver update (`du_dpm_solver_update`), the perturbation with its
noise drawn in the kernel (`du_perturb_randn`), the moments over M (`du_moments`, written straight into the accumulation
slot).  The host scalars (sigma / alpha / lambda / h of a step) are computed with the reference's own 0-dim fp32 tensor
expressions (:612-616, :671-690) and handed to the kernels as floats.

Reference behaviour kept on purpose:
  * in the window, `model_output` is already the CONVERTED output (the x0 prediction of DPM-Solver++, :931), so the
    "pred_original_sample" the block re-noises, the centre of the second moment and the returned `pred_epsilon` are all
    derived from that x0 prediction (:958-960, :974, :982);
  * outside the window the return value has `prev_sample` only (:984).
What the reference cannot run is not provided: `algorithm_type` "dpmsolver" / "sde-dpmsolver" call an undefined
`deprecate` in the constructor (:215-217) and the SDE types an undefined `randn_tensor` (:939) — `NotImplementedError`
here; `predict_next=True` raises `NotImplementedError` in the reference too (:965).  Karras / Lu sigma spacings, dynamic
thresholding and the third-order update (solver_order 3, never configured by the reference) are not on the path.
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch

from .. import ops
from ..configuration import ConfigurableScheduler, records_config
from ..outputs import DDIMSchedulerUncertaintyOutput
from ._core import betas_for_alpha_bar
from .mixin import SchedulerUncertaintyClassConditionedMixin, SchedulerUncertaintyMixin  # noqa: F401
from .traits import PredictorClassConditionedTrait


class SchedulerUncertaintyOutput(DDIMSchedulerUncertaintyOutput):
    """`prev_sample`, `pred_original_sample`, `uncertainty`, `pred_epsilon` (reference :34-55)."""
    __slots__ = ()


class SchedulerOutput(DDIMSchedulerUncertaintyOutput):
    """diffusers' plain `SchedulerOutput(prev_sample=...)`, returned outside the window (reference :984)."""
    __slots__ = ()


class KDPM2DiscreteSchedulerUncertainty(ConfigurableScheduler):   # structurally a SchedulerUncertaintyMixin (mixin.py)
    order = 1
    _compatibles: List[str] = []
    has_compatibles = True

    @records_config
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas: Optional[Union[np.ndarray, List[float]]] = None,
                 solver_order: int = 2, prediction_type: str = "epsilon", thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, sample_max_value: float = 1.0, algorithm_type: str = "dpmsolver++",
                 solver_type: str = "midpoint", lower_order_final: bool = True, euler_at_final: bool = False,
                 use_karras_sigmas: Optional[bool] = False, use_lu_lambdas: Optional[bool] = False,
                 final_sigmas_type: Optional[str] = "zero", lambda_min_clipped: float = -float("inf"),
                 variance_type: Optional[str] = None, timestep_spacing: str = "linspace", steps_offset: int = 0, M: int = 30,
                 after_step: int = 10, num_steps_uc: int = 10, unet=None, predict_next: bool = False,
                 prompt_embed: Optional[torch.Tensor] = None):
        if algorithm_type in ("dpmsolver", "sde-dpmsolver", "sde-dpmsolver++"):
            raise NotImplementedError(f"algorithm_type {algorithm_type}: the reference scheduler fails on it as well (undefined "
                                      "`deprecate` / `randn_tensor`, scheduling_dpm_2_uncertainty_centered.py:215-217, 939)")
        if algorithm_type == "deis":
            self.config.algorithm_type = algorithm_type = "dpmsolver++"
        elif algorithm_type != "dpmsolver++":
            raise NotImplementedError(f"{algorithm_type} does is not implemented for {self.__class__}")
        if solver_type not in ("midpoint", "heun"):
            if solver_type in ("logrho", "bh1", "bh2"):
                self.config.solver_type = "midpoint"
            else:
                raise NotImplementedError(f"{solver_type} does is not implemented for {self.__class__}")
        if solver_order not in (1, 2):
            raise NotImplementedError("solver_order 3 (third-order multistep update) is not on the uncertainty path")
        if use_karras_sigmas or use_lu_lambdas or thresholding:
            raise NotImplementedError("Karras / Lu sigma spacings and dynamic thresholding are not on the uncertainty path")
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            self.betas = betas_for_alpha_bar(num_train_timesteps)
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alpha_t = torch.sqrt(self.alphas_cumprod)
        self.sigma_t = torch.sqrt(1 - self.alphas_cumprod)
        self.lambda_t = torch.log(self.alpha_t) - torch.log(self.sigma_t)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=np.float32)[::-1].copy())
        self.model_outputs = [None] * solver_order
        self.lower_order_nums = 0
        self._step_index = None
        self.M = M
        self.after_step = after_step
        self.num_steps_uc = num_steps_uc
        self.unet = unet
        self.predict_next = predict_next
        self.prompt_embed = prompt_embed
        self.prompt_embeds = prompt_embed
        self.timestep_after_step = None
        self.timestep_end_step = None
        self.map_sink = None        # optional UncertaintyMapAccumulator: maps go straight into its slots (F8)
        self.map_in_sink = False    # whether the last in-window step wrote its map there (else the caller stashes it)

    # ------------------------------------------------------------------------------------------------ schedule
    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps: int = None, device: Union[str, torch.device, None] = None):
        """reference :285-365 (plain sigma spacing)."""
        clipped_idx = torch.searchsorted(torch.flip(self.lambda_t, [0]), self.config.lambda_min_clipped)
        last_timestep = ((self.config.num_train_timesteps - clipped_idx).numpy()).item()
        spacing = self.config.timestep_spacing
        if spacing == "linspace":
            timesteps = np.linspace(0, last_timestep - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif spacing == "leading":
            step_ratio = last_timestep // (num_inference_steps + 1)
            timesteps = (np.arange(0, num_inference_steps + 1) * step_ratio).round()[::-1][:-1].copy().astype(np.int64)
            timesteps += self.config.steps_offset
        elif spacing == "trailing":
            step_ratio = self.config.num_train_timesteps / num_inference_steps
            timesteps = np.arange(last_timestep, 0, -step_ratio).round().copy().astype(np.int64)
            timesteps -= 1
        else:
            raise ValueError(f"{spacing} is not supported. Please make sure to choose one of 'linspace', 'leading' or 'trailing'.")
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy().copy()    # `np.array(tensor)` of the reference (:322)
        sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        if self.config.final_sigmas_type == "sigma_min":
            sigma_last = ((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5
        elif self.config.final_sigmas_type == "zero":
            sigma_last = 0
        else:
            raise ValueError(f"`final_sigmas_type` must be one of 'zero', or 'sigma_min', but got {self.config.final_sigmas_type}")
        sigmas = np.concatenate([sigmas, [sigma_last]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)          # CPU, like the reference (:349)
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self._host_timesteps = [int(t) for t in timesteps]
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self._step_index = None
        self.timestep_after_step = self._host_timesteps[self.config.after_step]
        self.timestep_end_step = self._host_timesteps[self.config.after_step + self.config.num_steps_uc - 1]

    def uncertainty_timesteps(self) -> List[int]:
        return [t for t in self._host_timesteps if self.timestep_end_step <= t <= self.timestep_after_step]

    def attach_accumulator(self, acc) -> None:
        self.map_sink = acc

    @staticmethod
    def _sigma_to_alpha_sigma_t(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def _init_step_index(self, timestep):
        t = int(timestep)
        hits = [i for i, v in enumerate(self._host_timesteps) if v == t]
        if len(hits) == 0:
            self._step_index = len(self._host_timesteps) - 1
        elif len(hits) > 1:
            self._step_index = hits[1]
        else:
            self._step_index = hits[0]

    # ------------------------------------------------------------------------------------------------ pieces of step()
    def convert_model_output(self, model_output: torch.Tensor, *args, sample: torch.Tensor = None, **kwargs) -> torch.Tensor:
        """DPM-Solver++ integrates the data prediction (reference :523-541): x0 = (sample - sigma_t eps) / alpha_t."""
        if sample is None:
            if len(args) > 1:
                sample = args[1]
            else:
                raise ValueError("missing `sample` as a required keyward argument")
        pt = self.config.prediction_type
        if pt == "sample":
            return model_output
        if pt not in ("epsilon", "v_prediction"):
            raise ValueError(f"prediction_type given as {pt} must be one of `epsilon`, `sample`, or `v_prediction` for the "
                             "DPMSolverMultistepScheduler.")
        if pt == "epsilon" and self.config.variance_type in ("learned", "learned_range"):
            model_output = model_output[:, :3]
        alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(self.sigmas[self.step_index])
        c = ops.make_coeffs(float(alpha_t), float(sigma_t), 0.0, 0.0, clip_sample=False, prediction_type=pt)
        return ops.ddim_step(model_output, sample, c, want_prev=False, want_x0=True)[1]

    def _lambdas(self, *idx):
        out = []
        for i in idx:
            a, s = self._sigma_to_alpha_sigma_t(self.sigmas[i])
            out.append((a, s, torch.log(a) - torch.log(s)))
        return out

    def dpm_solver_first_order_update(self, model_output: torch.Tensor, *args, sample: torch.Tensor = None, noise=None, **kw):
        """reference :612-620: x_t = (sigma_t / sigma_s) sample - (alpha_t (exp(-h) - 1)) x0."""
        (alpha_t, sigma_t, lambda_t), (_, sigma_s, lambda_s) = self._lambdas(self.step_index + 1, self.step_index)
        h = lambda_t - lambda_s
        a = sigma_t / sigma_s
        b = alpha_t * (torch.exp(-h) - 1.0)
        return ops.dpm_solver_update(sample, model_output, None, float(a), -float(b))

    def multistep_dpm_solver_second_order_update(self, model_output_list: List[torch.Tensor], *args, sample: torch.Tensor = None,
                                                 noise=None, **kw):
        """reference :671-700: D0 = m0, D1 = (1 / r0)(m0 - m1);
        midpoint  x_t = (sigma_t / sigma_s0) sample - (alpha_t (exp(-h) - 1)) D0 - 0.5 (alpha_t (exp(-h) - 1)) D1
        heun      x_t = (sigma_t / sigma_s0) sample - (alpha_t (exp(-h) - 1)) D0 + (alpha_t ((exp(-h) - 1) / h + 1)) D1"""
        (alpha_t, sigma_t, lambda_t), (_, sigma_s0, lambda_s0), (_, _, lambda_s1) = self._lambdas(
            self.step_index + 1, self.step_index, self.step_index - 1)
        m0, m1 = model_output_list[-1], model_output_list[-2]
        h, h_0 = lambda_t - lambda_s0, lambda_s0 - lambda_s1
        r0 = h_0 / h
        k = 1.0 / r0
        a = sigma_t / sigma_s0
        b = alpha_t * (torch.exp(-h) - 1.0)
        if self.config.solver_type == "midpoint":
            c = -float(0.5 * (alpha_t * (torch.exp(-h) - 1.0)))
        else:
            c = float(alpha_t * ((torch.exp(-h) - 1.0) / h + 1.0))
        return ops.dpm_solver_update(sample, m0, m1, float(a), -float(b), c, float(k))

    def predict_model(self, x, t):
        """diffusers UNet2DConditionModel call convention, as the DDIM family's base class."""
        return self.unet(x, t, encoder_hidden_states=self.prompt_embeds, cross_attention_kwargs=None, return_dict=False)[0]

    def _uncertainty_block(self, model_output: torch.Tensor, sample: torch.Tensor, timestep) -> tuple:
        """reference :955-974.  `model_output` is the converted output; the block treats it as a noise prediction."""
        alpha_prod_t = self.alphas_cumprod[int(timestep)]
        beta_prod_t = 1 - alpha_prod_t
        c = ops.make_coeffs(float(alpha_prod_t ** 0.5), float(beta_prod_t ** 0.5), 0.0, 0.0, clip_sample=False)
        pred_original_sample = ops.ddim_step(model_output, sample, c, want_prev=False, want_x0=True)[1]
        if self.predict_next:
            raise NotImplementedError("predict_next not implemented yet")
        sa, sb = float(torch.sqrt(alpha_prod_t)), float(torch.sqrt(1 - alpha_prod_t))
        scores = []
        for _ in range(self.M):
            x_t_hat = ops.perturb_fresh(pred_original_sample, sa, sb)     # `noise = torch.randn_like(...)` drawn in the kernel
            x_t_hat = self.scale_model_input(x_t_hat, timestep)
            scores.append(self.predict_model(x_t_hat, timestep))
        self.map_in_sink = self.map_sink is not None and self.map_sink.accepts(scores[0].shape, torch.float32)
        out = self.map_sink.next_slot(scores[0].shape, torch.float32) if self.map_in_sink else None
        uncertainty = ops.moments(scores, center=model_output, mode="centered", out=out, out_dtype=torch.float32)
        return uncertainty, pred_original_sample

    # ------------------------------------------------------------------------------------------------ step
    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor, generator=None, return_dict: bool = True):
        """reference :876-984."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self.step_index is None:
            self._init_step_index(timestep)
        n = len(self._host_timesteps)
        lower_order_final = (self.step_index == n - 1) and (
            self.config.euler_at_final or (self.config.lower_order_final and n < 15) or self.config.final_sigmas_type == "zero")
        lower_order_second = (self.step_index == n - 2) and self.config.lower_order_final and n < 15

        model_output = self.convert_model_output(model_output, sample=sample)
        for i in range(self.config.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = model_output

        if self.config.solver_order == 1 or self.lower_order_nums < 1 or lower_order_final:
            prev_sample = self.dpm_solver_first_order_update(model_output, sample=sample)
        else:   # solver_order == 2 (the `lower_order_second` clause of the reference only matters for order 3)
            prev_sample = self.multistep_dpm_solver_second_order_update(self.model_outputs, sample=sample)
        del lower_order_second
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1

        t = int(timestep)
        window = self.timestep_end_step <= t <= self.timestep_after_step
        if window:
            uncertainty, pred_original_sample = self._uncertainty_block(model_output, sample, timestep)
        self._step_index += 1
        if not return_dict:
            return (prev_sample,)
        if window:
            return SchedulerUncertaintyOutput(prev_sample=prev_sample, uncertainty=uncertainty,
                                              pred_original_sample=pred_original_sample, pred_epsilon=model_output)
        return SchedulerOutput(prev_sample=prev_sample)

    # ------------------------------------------------------------------------------------------------ misc API
    def scale_model_input(self, sample: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        return sample

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        """reference :994-1027: alpha_t x + sigma_t n with the (alpha, sigma) of each sample's schedule position."""
        step_indices = []
        for timestep in timesteps.reshape(-1).tolist():
            hits = [i for i, v in enumerate(self._host_timesteps) if v == int(timestep)]
            step_indices.append(len(self._host_timesteps) - 1 if not hits else (hits[1] if len(hits) > 1 else hits[0]))
        if len(set(step_indices)) == 1:
            alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(self.sigmas[step_indices[0]].to(original_samples.dtype))
            return ops.perturb(original_samples, noise, float(alpha_t), float(sigma_t))
        sigma = self.sigmas.to(device=original_samples.device, dtype=original_samples.dtype)[step_indices].flatten()
        while len(sigma.shape) < len(original_samples.shape):
            sigma = sigma.unsqueeze(-1)
        alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(sigma)
        return alpha_t * original_samples + sigma_t * noise      # per-sample timesteps: a training-time call, not on the path

    def __len__(self):
        return self.config.num_train_timesteps


class KDPM2SchedulerUncertaintyImagenet(KDPM2DiscreteSchedulerUncertainty):
    def predict_model(self, x, t):
        return self.unet(x, t)[:, :3]


class KDPM2SchedulerUncertaintyImagenetClassConditioned(PredictorClassConditionedTrait, KDPM2SchedulerUncertaintyImagenet):
    class_conditioned: bool = True      # structurally a SchedulerUncertaintyClassConditionedMixin


__all__ = ["SchedulerUncertaintyOutput", "KDPM2DiscreteSchedulerUncertainty", "KDPM2SchedulerUncertaintyImagenet",
           "KDPM2SchedulerUncertaintyImagenetClassConditioned"]
