"""Shared machinery of the drop-in uncertainty schedulers.

The reference ships 18 scheduler modules that are each a full copy of diffusers' DDIM scheduler with one edited block
inside `step()`.  Here the DDIM bookkeeping exists once (`UncertaintyDDIMCore`), every variant is a
small subclass overriding `_uncertainty_block`, and ALL tensor arithmetic of `step()` runs in libdu_b200.so through
`ops` (moments over M, perturbation builders, DDIM update, masked re-step) — there is no torch arithmetic on the data
path and no CPU fallback.  What stays in torch, on purpose:

  * the score model call (`predict_model`): ADM / U-ViT / diffusers UNets remain the reference modules (north_star);
  * noise generation (`torch.randn_like`, `torch.randn(generator=...)`): the RNG stream consumed per step is part of
    parity on identical inputs (SURVEY.md §7 "RNG parity"), so draws happen in the reference's order, including the
    `best_noise` drawn every step even for eta == 0 (…zigzag_centered.py:500) and the never-used `variance_noise`
    (…:519-523);
  * the per-step host scalars: computed with the reference's own 0-dim fp32 CPU tensor expressions
    (…zigzag_centered.py:462-468, 294-302, 497-498, 507) so that they carry the same roundings, then handed to the
    kernels as plain floats.

Reference being mirrored (API, argument meaning, error behaviour):
schedulers_uncertainty/scheduling_ddim_uncertainty_zigzag_centered.py:138-651 (ctor :194-281, set_timesteps :338-387,
step :400-559, add_noise :593-626, get_velocity :629-646).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from .. import ops
from ..configuration import ConfigurableScheduler, records_config
from ..outputs import DDIMSchedulerUncertaintyOutput


def betas_for_alpha_bar(num_diffusion_timesteps: int, max_beta: float = 0.999, alpha_transform_type: str = "cosine") -> torch.Tensor:
    """Glide cosine schedule, `squaredcos_cap_v2` (…zigzag_centered.py:58-99)."""
    if alpha_transform_type == "cosine":
        def alpha_bar_fn(t):
            return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    elif alpha_transform_type == "exp":
        def alpha_bar_fn(t):
            return math.exp(t * -12.0)
    else:
        raise ValueError(f"Unsupported alpha_tranform_type: {alpha_transform_type}")
    betas = []
    for i in range(num_diffusion_timesteps):
        t1, t2 = i / num_diffusion_timesteps, (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar_fn(t2) / alpha_bar_fn(t1), max_beta))
    return torch.tensor(betas, dtype=torch.float32)


def rescale_zero_terminal_snr(betas: torch.Tensor) -> torch.Tensor:
    """Zero-terminal-SNR rescale (…zigzag_centered.py:102-135; arXiv 2305.08891 Alg. 1)."""
    alphas_bar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
    first, last = alphas_bar_sqrt[0].clone(), alphas_bar_sqrt[-1].clone()
    alphas_bar_sqrt = (alphas_bar_sqrt - last) * (first / (first - last))
    alphas_bar = alphas_bar_sqrt ** 2
    alphas = torch.cat([alphas_bar[0:1], alphas_bar[1:] / alphas_bar[:-1]])
    return 1 - alphas


class StepState:
    """Everything one `step()` call knows, handed to the variant's `_uncertainty_block`."""
    __slots__ = ("model_output", "sample", "t", "prev_t", "eta", "use_clipped", "coeffs", "host", "prev", "x0", "eps")

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


class UncertaintyDDIMCore(ConfigurableScheduler):
    """DDIM scheduler whose `step()` also estimates a per-pixel uncertainty map inside a timestep window."""

    order = 1
    _compatibles: List[str] = []
    has_compatibles = True
    # variant switches
    draws_best_noise = True     # `best_noise = torch.randn_like(x0)` every step (all variants except MC-dropout)
    eta_uses_best_noise = True  # eta > 0 adds std * best_noise, ignoring variance_noise (…zigzag_centered.py:523)
    predict_next_fixed: Optional[bool] = None  # zig-zag variants hard-wire predict_next = True (:272)
    host_copies = False         # MC-dropout returns x0 / score / pred_epsilon as CPU tensors (scheduling_ddim_mc_dropout.py:551-554)

    @records_config
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas: Optional[Union[np.ndarray, List[float]]] = None,
                 clip_sample: bool = True, set_alpha_to_one: bool = True, steps_offset: int = 0,
                 prediction_type: str = "epsilon", thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995,
                 clip_sample_range: float = 1.0, sample_max_value: float = 1.0, timestep_spacing: str = "leading",
                 rescale_betas_zero_snr: bool = False, M: int = 30, after_step: int = 10, num_steps_uc: int = 10,
                 unet=None, prompt_embed: Optional[torch.Tensor] = None, prompt_embeds: Optional[torch.Tensor] = None,
                 y: Optional[torch.Tensor] = None, predict_next: bool = False, debug: bool = False):
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
        if rescale_betas_zero_snr:
            self.betas = rescale_zero_terminal_snr(self.betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)          # CPU fp32 [T], like the reference
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        for name in ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "trained_betas", "clip_sample",
                     "set_alpha_to_one", "steps_offset", "prediction_type", "thresholding", "dynamic_thresholding_ratio",
                     "clip_sample_range", "sample_max_value", "timestep_spacing", "rescale_betas_zero_snr"):
            setattr(self, name, locals()[name])
        self.M = M
        self.after_step = after_step
        self.num_steps_uc = num_steps_uc
        self.unet = unet
        self.predict_next = predict_next if self.predict_next_fixed is None else self.predict_next_fixed
        self.prompt_embeds = next((v for v in (prompt_embeds, prompt_embed, y) if v is not None), None)
        self.debug = debug
        self.timestep_after_step = None
        self.timestep_end_step = None
        self.map_sink = None            # optional UncertaintyMapAccumulator (F8): maps are written straight into its slots
        self.map_in_sink = False        # whether the LAST in-window step wrote its map into the sink (else the caller stashes it)
        self._scalar_cache: Dict[Tuple, Tuple] = {}

    # ------------------------------------------------------------------------------------------ plain DDIM API
    def scale_model_input(self, sample: torch.Tensor, timestep: Optional[int] = None) -> torch.Tensor:
        return sample

    def _get_variance(self, timestep, prev_timestep):
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        return (beta_prod_t_prev / beta_prod_t) * (1 - alpha_prod_t / alpha_prod_t_prev)

    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device, None] = None):
        """…zigzag_centered.py:338-387 (without the two debug prints)."""
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {self.config.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                f" maximal {self.config.num_train_timesteps} timesteps.")
        self.num_inference_steps = num_inference_steps
        spacing = self.config.timestep_spacing
        if spacing == "linspace":
            timesteps = np.linspace(0, self.config.num_train_timesteps - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif spacing == "leading":
            step_ratio = self.config.num_train_timesteps // self.num_inference_steps
            timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
            timesteps += self.config.steps_offset
        elif spacing == "trailing":
            step_ratio = self.config.num_train_timesteps / self.num_inference_steps
            timesteps = np.round(np.arange(self.config.num_train_timesteps, 0, -step_ratio)).astype(np.int64)
            timesteps -= 1
        else:
            raise ValueError(f"{spacing} is not supported. Please make sure to choose one of 'leading' or 'trailing'.")
        self.timesteps = torch.from_numpy(timesteps).to(device)
        self._host_timesteps = [int(t) for t in timesteps]
        self.timestep_after_step = self._host_timesteps[self.config.after_step]
        self.timestep_end_step = self._host_timesteps[self.config.after_step + self.config.num_steps_uc - 1]
        self._scalar_cache.clear()

    def uncertainty_timesteps(self) -> List[int]:
        """Timesteps of the current schedule that fall inside the uncertainty window (T_uc of the accumulation buffer)."""
        return [t for t in self._host_timesteps if self.timestep_end_step <= t <= self.timestep_after_step]

    def in_window(self, timestep: int) -> bool:
        return self.timestep_end_step <= timestep <= self.timestep_after_step

    # ------------------------------------------------------------------------------------------ host scalars
    def _step_scalars(self, t: int, eta: float, use_clipped: bool):
        """(DdimCoeffs for the kernels, dict of the reference's 0-dim tensors).  Cached per (t, eta, flags)."""
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        key = (t, prev_t, float(eta), bool(use_clipped), self.config.prediction_type, bool(self.config.clip_sample),
               float(self.config.clip_sample_range), bool(self.config.thresholding))
        hit = self._scalar_cache.get(key)
        if hit is not None:
            return hit
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        variance = self._get_variance(t, prev_t)
        std_dev_t = eta * variance ** (0.5)
        host = dict(alpha_prod_t=alpha_prod_t, alpha_prod_t_prev=alpha_prod_t_prev, beta_prod_t=beta_prod_t,
                    std_dev_t=std_dev_t, prev_t=prev_t,
                    sqrt_alpha_t=float(alpha_prod_t ** (0.5)), sqrt_beta_t=float(beta_prod_t ** (0.5)),
                    sqrt_alpha_prev=float(alpha_prod_t_prev ** (0.5)),
                    dir_coef=float((1 - alpha_prod_t_prev - std_dev_t ** 2) ** (0.5)), sigma=float(std_dev_t))
        coeffs = ops.make_coeffs(host["sqrt_alpha_t"], host["sqrt_beta_t"], host["sqrt_alpha_prev"], host["dir_coef"],
                                 sigma=host["sigma"], clip_sample=bool(self.config.clip_sample) and not self.config.thresholding,
                                 clip_range=float(self.config.clip_sample_range), prediction_type=self.config.prediction_type,
                                 use_clipped_model_output=use_clipped, add_noise=eta > 0)
        self._scalar_cache[key] = (coeffs, host)
        return coeffs, host

    # ------------------------------------------------------------------------------------------ the step
    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise: Optional[torch.Tensor] = None,
             return_dict: bool = True):
        """x_{t-1} by the DDIM rule and, inside the window, the uncertainty map of this step.
        Same signature, return type and exceptions as …zigzag_centered.py:400-559."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self.config.prediction_type not in ("epsilon", "sample", "v_prediction"):
            raise ValueError(f"prediction_type given as {self.config.prediction_type} must be one of `epsilon`, `sample`, or"
                             " `v_prediction`")
        self._check_prediction_type()
        t = int(timestep)
        coeffs, host = self._step_scalars(t, eta, use_clipped_model_output)
        window = self.in_window(t)
        st = StepState(model_output=model_output, sample=sample, t=t, prev_t=host["prev_t"], eta=eta,
                       use_clipped=use_clipped_model_output, coeffs=coeffs, host=host, prev=None, x0=None, eps=None)

        self.map_in_sink = False
        pre = self._before_update(st) if window else None   # MC-dropout samples before the update (RNG order)

        best_noise = torch.randn_like(sample if sample.dtype.is_floating_point else model_output) if self.draws_best_noise else None
        if eta > 0:
            if variance_noise is not None and generator is not None:
                raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either `generator` or"
                                 " `variance_noise` stays `None`.")
            if variance_noise is None:  # drawn even where the reference then ignores it: keeps the RNG stream identical
                variance_noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                             dtype=model_output.dtype)
        noise = (best_noise if self.eta_uses_best_noise else variance_noise) if eta > 0 else None
        self._ddim_update(st, noise)

        uncertainty = None
        if window:
            uncertainty = pre if pre is not None else self._uncertainty_block(st)
        if not return_dict:
            return (st.prev,)
        out = DDIMSchedulerUncertaintyOutput(prev_sample=st.prev, pred_original_sample=st.x0)
        if window:
            out.uncertainty = uncertainty
            out.pred_epsilon = st.eps
        self._finish_output(out, st, window)
        return out

    def _check_prediction_type(self):
        pass

    def _before_update(self, st: StepState):
        return None

    def _finish_output(self, out, st: StepState, window: bool):
        pass

    def _ddim_update(self, st: StepState, noise: Optional[torch.Tensor]):
        """F3 in one launch (du_ddim_step); dynamic thresholding (never enabled by the reference's configs) splits it."""
        need_eps = self.config.prediction_type != "epsilon" or st.use_clipped
        if not self.config.thresholding:
            st.prev, st.x0, eps = ops.ddim_step(st.model_output, st.sample, st.coeffs, noise=noise, want_eps=need_eps)
            st.eps = eps if need_eps else st.model_output
            return
        # thresholding=True: x0 unclipped -> per-image dynamic threshold -> eps / x_{t-1} from the thresholded x0
        c0 = ops.make_coeffs(st.host["sqrt_alpha_t"], st.host["sqrt_beta_t"], 0.0, 0.0, clip_sample=False,
                             prediction_type=self.config.prediction_type)
        _, x0, eps = ops.ddim_step(st.model_output, st.sample, c0, want_prev=False, want_eps=True)
        x0 = self._threshold_sample(x0)
        if st.use_clipped:
            # eps = (sample - sqrt(abar) x0) / sqrt(1-abar): the reference's expression, one rounding per op
            eps = ops.perturb(st.sample, x0, 1.0, -st.host["sqrt_alpha_t"])
            eps = ops.perturb(eps, eps, 1.0 / st.host["sqrt_beta_t"], 0.0)
        prev = ops.perturb(x0, eps, st.host["sqrt_alpha_prev"], st.host["dir_coef"])
        if noise is not None:
            prev = ops.perturb(prev, noise, 1.0, st.host["sigma"])
        st.prev, st.x0, st.eps = prev, x0, eps

    def _restep_general(self, st: StepState, x0_source: torch.Tensor, eps_guided: torch.Tensor):
        """The in-scheduler re-step in its general form (…uncertainty_threshold.py:554-574, …uncertainty_grad.py:551-570): x0 from
        `x0_source` (the UNGUIDED model output) with clamp or dynamic thresholding, the guided score re-derived from the clipped x0
        when use_clipped_model_output, x_(t-1) without the eta noise.  Separate launches; the one-launch du_guided_step covers
        the configurations the reference's own configs use (no thresholding, score not re-derived)."""
        h = st.host
        c0 = ops.make_coeffs(h["sqrt_alpha_t"], h["sqrt_beta_t"], 0.0, 0.0, prediction_type="epsilon",
                             clip_sample=bool(self.config.clip_sample) and not self.config.thresholding,
                             clip_range=float(self.config.clip_sample_range))
        x0 = ops.ddim_step(x0_source, st.sample, c0, want_prev=False, want_x0=True)[1]
        if self.config.thresholding:
            x0 = self._threshold_sample(x0)
        eps = eps_guided
        if st.use_clipped:
            eps = ops.perturb(st.sample, x0, 1.0, -h["sqrt_alpha_t"])
            eps = ops.scale(eps, 1.0 / h["sqrt_beta_t"])
        st.prev, st.x0, st.eps = ops.perturb(x0, eps, h["sqrt_alpha_prev"], h["dir_coef"]), x0, eps

    def _threshold_sample(self, sample: torch.Tensor) -> torch.Tensor:
        """Imagen dynamic thresholding (…zigzag_centered.py:305-336): per-image quantile of |x0| by the radix-select
        kernel, then clamp / rescale."""
        dtype = sample.dtype
        x = sample.float() if dtype not in (torch.float32, torch.float64) else sample
        flat = x.reshape(x.shape[0], -1)
        s = ops.quantile_threshold(flat.abs(), self.config.dynamic_thresholding_ratio, lerp_fma=True)
        s = torch.clamp(s, min=1, max=self.config.sample_max_value).unsqueeze(1)
        flat = torch.clamp(flat, -s, s) / s
        return flat.reshape(sample.shape).to(dtype)

    # ------------------------------------------------------------------------------------------ uncertainty helpers
    def _uncertainty_block(self, st: StepState) -> torch.Tensor:
        raise NotImplementedError

    def _map_out(self, like: torch.Tensor, dtype: torch.dtype = torch.float32) -> Optional[torch.Tensor]:
        """Destination of this step's map: the next slot of the attached accumulation buffer (F8 fused) or None."""
        self.map_in_sink = False
        if self.map_sink is None or not self.map_sink.accepts(like.shape, dtype):
            # no sink, or a map the sink's slots cannot hold as they are (fp16 / bf16 maps of the `var` modes under autocast,
            # the [B,1,H,W] map of flip_threshold): the step returns a fresh tensor with the reference's dtype and shape and
            # the sampling loop stashes it (du_accumulate_slot converts)
            return None
        self.map_in_sink = True
        return self.map_sink.next_slot(like.shape, dtype)

    def _perturbed_input(self, st: StepState, base_x0: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """F7: the model input of one perturbed forward.  predict_next: sqrt(1-beta_t) x_{t-1} + sqrt(beta_t) n
        (…zigzag_centered.py:538); else add_noise(x0, n, t) (:535, 593-626).  One launch: with noise=None the draw
        `n = torch.randn_like(pred_x_0)` (:529) happens inside it (du_perturb_randn, torch's Philox stream bit for bit)."""
        t = st.t
        if self.predict_next:
            a, b = float(torch.sqrt(1 - self.betas[t])), float(torch.sqrt(self.betas[t]))
            base = st.prev
        else:
            a, b = float(self.alphas_cumprod[t] ** 0.5), float((1 - self.alphas_cumprod[t]) ** 0.5)
            base = base_x0
        x_hat = ops.perturb_fresh(base, a, b, noise_like=st.x0) if noise is None else ops.perturb(base, noise, a, b)
        return self.scale_model_input(x_hat, t)

    def _perturbed_scores(self, st: StepState) -> List[torch.Tensor]:
        """M forwards on independently re-noised inputs (…uncertainty_centered.py:522-538)."""
        return [self.predict_model(self._perturbed_input(st, st.x0), st.t) for _ in range(self.M)]

    def _reduce(self, scores: List[torch.Tensor], mode: str, center: Optional[torch.Tensor] = None,
                out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """F1 (+F8): one launch over the M score tensors, no stacking; strided `[:, :3]` views are read in place."""
        like = scores[0]
        if mode == "centered":
            dt = torch.float32   # pow() is an fp32 op under autocast (SURVEY.md §7 "Dtypes under autocast")
        else:
            dt = out_dtype or like.dtype   # torch.var keeps the input dtype
        return ops.moments(scores, center=center, mode=mode, out=self._map_out(like, dt), out_dtype=dt)

    # ------------------------------------------------------------------------------------------ model dispatch
    def predict_model(self, x, t):
        """diffusers UNet2DConditionModel call convention (…zigzag_centered.py:561-569)."""
        return self.unet(x, t, encoder_hidden_states=self.prompt_embeds, cross_attention_kwargs=None, return_dict=False)[0]

    # ------------------------------------------------------------------------------------------ misc DDIM API
    def _get_epsilon(self, sample: torch.Tensor, model_output: torch.Tensor, timestep: int) -> torch.Tensor:
        """…zigzag_centered.py:572-590."""
        if self.config.prediction_type == "epsilon":
            return model_output
        if self.config.prediction_type not in ("sample", "v_prediction"):
            raise ValueError(f"prediction_type given as {self.config.prediction_type} must be one of `epsilon`, `sample`, or"
                             " `v_prediction`")
        a = self.alphas_cumprod[int(timestep)]
        c = ops.make_coeffs(float(a ** 0.5), float((1 - a) ** 0.5), 0.0, 0.0, clip_sample=False,
                            prediction_type=self.config.prediction_type)
        return ops.ddim_step(model_output, sample, c, want_prev=False, want_x0=False, want_eps=True)[2]

    def _per_sample_scalars(self, timesteps, ref: torch.Tensor):
        ac = self.alphas_cumprod.to(device=ref.device, dtype=ref.dtype)
        if isinstance(timesteps, torch.Tensor):
            timesteps = timesteps.to(ref.device)
        sa = (ac[timesteps] ** 0.5).flatten()
        sb = ((1 - ac[timesteps]) ** 0.5).flatten()
        return sa, sb

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps) -> torch.Tensor:
        """sqrt(abar_t) x0 + sqrt(1-abar_t) n (…zigzag_centered.py:593-626).  One timestep: du_perturb (as a differentiable op when
        the inputs are on an autograd graph — the gradient schedulers re-noise a traced x0); a vector of per-sample timesteps:
        du_perturb_rows with the scalars read from device vectors."""
        sa, sb = self._per_sample_scalars(timesteps, original_samples)
        if not original_samples.is_cuda:
            raise RuntimeError("add_noise: the uncertainty path has no CPU fallback, move the tensors to a CUDA device")
        traced = torch.is_grad_enabled() and (original_samples.requires_grad or noise.requires_grad)
        if sa.numel() == 1:
            if traced:
                return ops.perturb_autograd(original_samples, noise, float(sa), float(sb))
            return ops.perturb(original_samples, noise, float(sa), float(sb))
        if traced:
            raise RuntimeError("add_noise: per-sample timesteps are not differentiable here (no reference caller needs it)")
        return ops.perturb_rows(original_samples, noise, sa, sb)

    def get_velocity(self, sample: torch.Tensor, noise: torch.Tensor, timesteps) -> torch.Tensor:
        """sqrt(abar_t) n - sqrt(1-abar_t) x (…zigzag_centered.py:629-646)."""
        sa, sb = self._per_sample_scalars(timesteps, sample)
        if not sample.is_cuda:
            raise RuntimeError("get_velocity: the uncertainty path has no CPU fallback, move the tensors to a CUDA device")
        if sa.numel() == 1:
            return ops.perturb(noise, sample, float(sa), -float(sb))
        return ops.perturb_rows(noise, sample, sa, -sb)

    def __len__(self):
        return self.config.num_train_timesteps

    # ------------------------------------------------------------------------------------------ F8
    def attach_accumulator(self, sink) -> None:
        """Write every in-window map straight into `sink` (an UncertaintyMapAccumulator) instead of a fresh tensor."""
        self.map_sink = sink
