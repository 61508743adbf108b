"""CPU oracle for the per-step uncertainty hot path of Michedev/diffusion-uncertainty.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this file.  The product
(`diffusion-uncertainty_b200/`) never imports it and has no CPU fallback.

Every function restates ONE row of SURVEY.md §8(a) in plain torch-CPU fp32 arithmetic,
one rounded operation per reference operation, in the reference's operation order, and cites the
reference lines it follows (paths relative to /root/reference/diffusion_uncertainty/,
SU = schedulers_uncertainty, PU = pipeline_uncertainty).

Pinning (see tests/golden/make_golden.py + tests/test_oracle_golden.py): the reference ships no
test or golden vector for this path, so the oracle is pinned against outputs of the UNMODIFIED
reference modules imported in the build container (with the `diffusers` base-class stand-in under
tests/golden/_standin) and committed as fixtures under tests/golden/*.npz, plus known-answer facts of
`torch.quantile` / `torch.var`.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# F1 — reductions over the M axis
# --------------------------------------------------------------------------------------------
def centered_second_moment(scores: Sequence[Tensor], eps: Tensor) -> Tensor:
    """F1a.  u = (1/M) sum_m (eps_hat_m - eps)^2   (NOT a variance: divides by M, centred on eps).
    SU/scheduling_ddim_uncertainty_zigzag_centered.py:549 (same expression in
    scheduling_ddim_uncertainty_centered.py:539, ..._centered_d.py:541, ..._uncertainty.py:540,
    scheduling_dpm_2_uncertainty_centered.py:968, PU/...guided_gradient.py:192)."""
    stacked = torch.stack([s.float() for s in scores], dim=0)
    return (stacked - eps.float().unsqueeze(0)).pow(2).mean(dim=0)


def variance_unbiased(scores: Sequence[Tensor]) -> Tensor:
    """F1b.  Unbiased variance (divide by M-1) over the M axis; M == 1 gives NaN.
    SU/scheduling_ddim_mc_dropout.py:506, SU/scheduling_ddim_uncertainty_threshold.py:537,
    SU/scheduling_ddim_infer_noise_multiscale_threshold.py:533, generate_samples.py:815."""
    return torch.var(torch.stack([s.float() for s in scores], dim=0), dim=0)


def variance_with_center(scores: Sequence[Tensor], eps: Tensor) -> Tensor:
    """F1c.  Unbiased variance over the M perturbed predictions PLUS the original eps (appended last).
    uncertainty_guidance.py:101-106; PU/...posterior_distribution.py:58-61, 232-235."""
    return variance_unbiased(list(scores) + [eps])


def mean_over_m(scores: Sequence[Tensor]) -> Tensor:
    """Mean over the M axis (north_star op (1); the reference never returns it but the M-shard merge
    needs it).  Accumulated in float64 so it is the correctly rounded reference value."""
    return torch.stack([s.double() for s in scores], dim=0).mean(dim=0).float()


def raw_second_moment(scores: Sequence[Tensor]) -> Tensor:
    """F1d.  mean_m(eps_hat_m^2).  uncertainty_guidance.py:48."""
    return torch.stack([s.float() for s in scores], dim=0).pow(2).mean(dim=0)


def std_over_m(scores: Sequence[Tensor]) -> Tensor:
    """F1d.  Unbiased std over M.  generate_samples.py:941."""
    return torch.stack([s.float() for s in scores], dim=0).std(dim=0)


# --------------------------------------------------------------------------------------------
# F2 — thresholds and masks
# --------------------------------------------------------------------------------------------
def quantile_rank(n: int, q: float) -> Tuple[int, int, np.float32]:
    """The (lo, hi, weight) triple torch.quantile's 'linear' interpolation uses: rank = fp32(q) *
    fp32(n-1) rounded to fp32, lo = floor, hi = ceil, w = rank - lo (aten/native/Sorting.cpp,
    quantile_compute; called at PU/...posterior_distribution.py:15)."""
    rank = np.float32(q) * np.float32(n - 1)
    lo = int(np.floor(rank))
    hi = int(np.ceil(rank))
    return lo, hi, np.float32(rank - np.float32(lo))


def fma_f32(a, b, c) -> np.float32:
    """a * b + c with ONE rounding to fp32 (round to nearest even) — what a hardware FMA returns —, computed exactly with rationals so
    that the result does not depend on the host's own contraction."""
    a, b, c = float(np.float32(a)), float(np.float32(b)), float(np.float32(c))
    if not (math.isfinite(a) and math.isfinite(b) and math.isfinite(c)):
        return np.float32(a * b + c)
    from fractions import Fraction
    exact = Fraction(a) * Fraction(b) + Fraction(c)
    if exact == 0:
        return np.float32(a * b + c)          # (signed zero as IEEE gives it)
    mag = abs(exact)
    e = mag.numerator.bit_length() - mag.denominator.bit_length()
    if Fraction(2) ** e > mag:
        e -= 1                                  # 2^e <= mag < 2^(e+1)
    quantum = Fraction(2) ** (max(e, -126) - 23)
    n, rem = divmod(mag, quantum)
    n = int(n)
    if rem * 2 > quantum or (rem * 2 == quantum and (n & 1)):
        n += 1
    val = float(n * quantum)                    # exact in double; 2^128 and above overflow to inf in the cast
    with np.errstate(over="ignore"):
        return np.float32(-val if exact < 0 else val)


def lerp_torch(a: np.float32, b: np.float32, w: np.float32, fma: bool = False) -> np.float32:
    """torch.lerp's two-branch formula (aten/native/Lerp.h).  fma=False: one fp32 rounding per operation.  fma=True: the multiply-add
    of the selected branch fused — what torch's CUDA kernel does (nvcc contracts it) and what torch's CPU kernels do where they are
    dispatched to an FMA-capable build (`vec::fmadd(coeff, end - start, base)`); which of the two a given host's torch.quantile
    returns is a property of that host (tests/test_oracle_golden.py checks that it is one of them, row by row)."""
    a, b, w = np.float32(a), np.float32(b), np.float32(w)
    diff = np.float32(b - a)
    if abs(w) < np.float32(0.5):
        return fma_f32(w, diff, a) if fma else np.float32(a + np.float32(w * diff))
    omw = np.float32(np.float32(1.0) - w)
    return fma_f32(-diff, omw, b) if fma else np.float32(b - np.float32(diff * omw))


def quantile_linear_rows(u2d: Tensor, q: float, lerp_fma: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """Independent restatement of torch.quantile(u2d, q, dim=1) (sort -> two order statistics ->
    lerp).  Returns (threshold[B], rank_lo_hi[B,2] int64, values_lo_hi[B,2]).  A row that contains
    a NaN gives a NaN threshold (torch masks such rows explicitly).
    PU/...posterior_distribution.py:15, uncertainty_guidance.py:112."""
    x = u2d.detach().float().cpu().numpy()
    B, n = x.shape
    lo, hi, w = quantile_rank(n, q)
    thr = np.empty(B, dtype=np.float32)
    vals = np.empty((B, 2), dtype=np.float32)
    for b in range(B):
        row = x[b]
        if np.isnan(row).any():
            thr[b] = np.nan
            vals[b] = np.nan
            continue
        # only the lo-th and hi-th order statistics are needed
        part = np.partition(row, (lo, hi))
        vals[b, 0], vals[b, 1] = part[lo], part[hi]
        thr[b] = lerp_torch(part[lo], part[hi], w, lerp_fma)
    ranks = torch.tensor([[lo, hi]] * B, dtype=torch.int64)
    return torch.from_numpy(thr), ranks, torch.from_numpy(vals)


def calculate_threshold_map(threshold, i: Optional[int], u: Tensor, threshold_type: str = "higher") -> Tensor:
    """F2a (float threshold = per-image percentile) and F2b (tensor threshold[i]).
    PU/...posterior_distribution.py:10-30; inline uncertainty_guidance.py:112-113."""
    if isinstance(threshold, float):
        shape = u.shape
        thr = torch.quantile(u.flatten(1).to(torch.float32), threshold, dim=1, keepdim=True)
        thr = thr.view(shape[0], *([1] * (len(shape) - 1)))
        m = (u > thr) if threshold_type == "higher" else (u < thr)
    else:
        thr_i = threshold[i]
        if u.dim() == 4:
            thr_i = thr_i.unsqueeze(0)
        m = (u > thr_i) if threshold_type == "higher" else (u < thr_i)
    return m.float()


def znorm(u: Tensor) -> Tensor:
    """F2c.  Whole-batch z-normalisation, unbiased std.  SU/scheduling_ddim_uncertainty_threshold.py:539-540."""
    return (u - u.mean()) / u.std()


def znorm_threshold_mask(z: Tensor, thr: float, mode: str = "max") -> Tensor:
    """F2c.  mode 'max' keeps z < thr, anything else keeps z > thr.
    SU/scheduling_ddim_uncertainty_threshold.py:549-554."""
    return ((z < thr) if mode == "max" else (z > thr)).float()


def multiscale_weights(z: Tensor) -> Tensor:
    """F2c multiscale bands: 0.8 for -3<z<-2, 0.9 for -2<z<-1, 1.0 for z>=-1, 0 elsewhere.
    SU/scheduling_ddim_infer_noise_multiscale_threshold.py:538-548."""
    m2 = ((z < -2.0) & (z > -3.0)).float()
    m1 = ((z < -1.0) & (z > -2.0)).float()
    m0 = (z >= -1.0).float()
    return m2 * 0.8 + m1 * 0.9 + m0


# --------------------------------------------------------------------------------------------
# F3 / F4 — DDIM update
# --------------------------------------------------------------------------------------------
def make_betas(beta_schedule: str = "linear", beta_start: float = 1e-4, beta_end: float = 0.02,
               num_train_timesteps: int = 1000, trained_betas=None) -> Tensor:
    """SU/scheduling_ddim_uncertainty_zigzag_centered.py:218-231, 58-99."""
    if trained_betas is not None:
        return torch.tensor(trained_betas, dtype=torch.float32)
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if beta_schedule == "squaredcos_cap_v2":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        T = num_train_timesteps
        return torch.tensor([min(1 - f((k + 1) / T) / f(k / T), 0.999) for k in range(T)], dtype=torch.float32)
    raise NotImplementedError(beta_schedule)


def leading_timesteps(num_train: int, num_inference: int, steps_offset: int = 0) -> np.ndarray:
    """'leading' spacing.  SU/scheduling_ddim_uncertainty_zigzag_centered.py:364-369."""
    ratio = num_train // num_inference
    return (np.arange(0, num_inference) * ratio).round()[::-1].copy().astype(np.int64) + steps_offset


class DDIMCoeffs:
    """Host scalars of one DDIM step, computed with the reference's own 0-dim fp32 tensor
    expressions.  SU/scheduling_ddim_uncertainty_zigzag_centered.py:462-468, 294-302, 497-498, 507."""

    def __init__(self, alphas_cumprod: Tensor, final_alpha_cumprod: Tensor, t: int, prev_t: int, eta: float):
        a_t = alphas_cumprod[t]
        a_prev = alphas_cumprod[prev_t] if prev_t >= 0 else final_alpha_cumprod
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        variance = (b_prev / b_t) * (1 - a_t / a_prev)
        std = eta * variance ** 0.5
        self.alpha_prod_t = a_t
        self.alpha_prod_t_prev = a_prev
        self.sqrt_alpha_t = a_t ** 0.5
        self.sqrt_beta_t = b_t ** 0.5
        self.sqrt_alpha_prev = a_prev ** 0.5
        self.dir_coef = (1 - a_prev - std ** 2) ** 0.5
        self.sigma = std


def ddim_step(model_output: Tensor, sample: Tensor, c: DDIMCoeffs, prediction_type: str = "epsilon",
              clip_sample: bool = True, clip_range: float = 1.0, eta: float = 0.0,
              noise: Optional[Tensor] = None, use_clipped_model_output: bool = False):
    """F3.  Returns (prev_sample, pred_original_sample, pred_epsilon).
    SU/scheduling_ddim_uncertainty_zigzag_centered.py:472-525."""
    if prediction_type == "epsilon":
        x0 = (sample - c.sqrt_beta_t * model_output) / c.sqrt_alpha_t
        eps = model_output
    elif prediction_type == "sample":
        x0 = model_output
        eps = (sample - c.sqrt_alpha_t * x0) / c.sqrt_beta_t
    elif prediction_type == "v_prediction":
        x0 = c.sqrt_alpha_t * sample - c.sqrt_beta_t * model_output
        eps = c.sqrt_alpha_t * model_output + c.sqrt_beta_t * sample
    else:
        raise ValueError(prediction_type)
    if clip_sample:
        x0 = x0.clamp(-clip_range, clip_range)
    if use_clipped_model_output:
        eps = (sample - c.sqrt_alpha_t * x0) / c.sqrt_beta_t
    prev = c.sqrt_alpha_prev * x0 + c.dir_coef * eps
    if eta > 0:
        prev = prev + c.sigma * noise
    return prev, x0, eps


def masked_restep(model_output: Tensor, sample: Tensor, weights: Tensor, c: DDIMCoeffs,
                  clip_sample: bool = True, clip_range: float = 1.0, use_clipped_model_output: bool = False):
    """F4.  eps' = eps * weights; x0 recomputed from the UNMASKED model_output; the eta noise is
    not re-added.  SU/scheduling_ddim_uncertainty_threshold.py:554-574."""
    eps = model_output * weights
    x0 = (sample - c.sqrt_beta_t * model_output) / c.sqrt_alpha_t
    if clip_sample:
        x0 = x0.clamp(-clip_range, clip_range)
    if use_clipped_model_output:
        eps = (sample - c.sqrt_alpha_t * x0) / c.sqrt_beta_t
    prev = c.sqrt_alpha_prev * x0 + c.dir_coef * eps
    return prev, x0, eps


# --------------------------------------------------------------------------------------------
# F5 / F6 — guided score blends
# --------------------------------------------------------------------------------------------
def posterior_blend(eps: Tensor, u: Tensor, mask: Tensor, M: int, alpha_hat_t, sum_source: Optional[Tensor] = None,
                    batch_sum: bool = True) -> Tensor:
    """F5.  post = [1/(M/u + 1/abar)] * (1/u) * S,  eps' = eps(1-mask) + mask*post.
    S = sum_source.sum(dim=0) (the reference's batch-axis sum, broadcast back; identity at B=1) when
    batch_sum, else sum_source itself.  uncertainty_guidance.py:115-120 (sum_source = eps);
    PU/...posterior_distribution.py:63-68,160 (sum_source = last perturbed prediction)."""
    src = eps if sum_source is None else sum_source
    S = src.sum(dim=0) if batch_sum else src
    inv_var = 1 / u
    post_var_trace = (M * inv_var) + (1 / alpha_hat_t)
    post_precision = 1 / post_var_trace
    post = post_precision * (inv_var * S)
    return (eps * (1 - mask)) + (mask * post)


def gradient_blend_masked(eps: Tensor, g: Tensor, mask: Tensor, lam: float) -> Tensor:
    """F6.  eps(1-m) + (eps + lam*g) m.  PU/...guided_gradient.py:117-118."""
    post = eps + lam * g
    return eps * (1 - mask) + post * mask


def gradient_add_masked(eps: Tensor, g: Tensor, mask: Tensor, lr: float) -> Tensor:
    """F6.  eps + lr*g*mask.  uncertainty_guidance.py:129."""
    return eps + (lr * g * mask)


# --------------------------------------------------------------------------------------------
# F7 — perturbation builders
# --------------------------------------------------------------------------------------------
def perturb_predict_next(prev_sample: Tensor, noise: Tensor, beta_t: Tensor) -> Tensor:
    """F7.  sqrt(1-beta_t) x_{t-1} + sqrt(beta_t) n.  SU/...zigzag_centered.py:538."""
    return torch.sqrt(1 - beta_t) * prev_sample + torch.sqrt(beta_t) * noise


def perturb_add_noise(x0: Tensor, noise: Tensor, alpha_prod_t: Tensor) -> Tensor:
    """F7.  sqrt(abar_t) x0 + sqrt(1-abar_t) n.  SU/...zigzag_centered.py:593-626 (add_noise);
    uncertainty_guidance.py:86-88."""
    return (alpha_prod_t ** 0.5) * x0 + ((1 - alpha_prod_t) ** 0.5) * noise


# --------------------------------------------------------------------------------------------
# F8 — accumulation
# --------------------------------------------------------------------------------------------
def accumulate_maps(per_batch_maps: List[List[Tensor]]) -> Tensor:
    """F8.  per batch: stack(dim=1) of the per-step maps -> [B,T_uc,...]; cat(dim=0) over batches.
    generate_samples.py:192-201, 229-231."""
    return torch.cat([torch.stack([m.cpu() for m in maps], dim=1) for maps in per_batch_maps], dim=0)


# --------------------------------------------------------------------------------------------
# Whole per-step chains (what bench.py's cpu_baseline / reference arm time, and what the fused CUDA
# step is checked against)
# --------------------------------------------------------------------------------------------
def uncertainty_step_posterior(scores: Sequence[Tensor], eps: Tensor, sample: Tensor, q: float, M: int,
                               alpha_hat_t, c: DDIMCoeffs, clip_sample: bool = True, batch_sum: bool = True,
                               sum_source: Optional[Tensor] = None, threshold_type: str = "higher"):
    """F1c -> F2a -> F5 -> F3: the percentile-guided posterior step of
    PU/...posterior_distribution.py:153-162 / uncertainty_guidance.py:101-120 followed by the DDIM
    update of SU/...zigzag_centered.py:472-510.  Returns (u, thr_mask, eps_guided, prev_sample, x0)."""
    u = variance_with_center(scores, eps)
    mask = calculate_threshold_map(float(q), None, u, threshold_type)
    eps_g = posterior_blend(eps, u, mask, M, alpha_hat_t, sum_source=sum_source, batch_sum=batch_sum)
    prev, x0, _ = ddim_step(eps_g, sample, c, clip_sample=clip_sample)
    return u, mask, eps_g, prev, x0


def uncertainty_step_znorm(scores: Sequence[Tensor], eps: Tensor, sample: Tensor, thr: float, mode: str,
                           c: DDIMCoeffs, clip_sample: bool = True, multiscale: bool = False, normalize: bool = True):
    """F1b -> F2c -> F4: the in-scheduler threshold step.
    SU/scheduling_ddim_uncertainty_threshold.py:537-574 / ..._multiscale_threshold.py:533-569.
    Returns (z, weights, prev_sample, x0, eps')."""
    u = variance_unbiased(scores)
    z = znorm(u) if normalize else u
    w = multiscale_weights(z) if multiscale else znorm_threshold_mask(z, thr, mode)
    prev, x0, eps2 = masked_restep(eps, sample, w, c, clip_sample=clip_sample)
    return z, w, prev, x0, eps2


# --------------------------------------------------------------------------------------------
# Whole scheduler step (L3), restated per variant from the F-rows above
# --------------------------------------------------------------------------------------------
def flip_uncertainty(eps: Tensor, flipped_back_output: Tensor, channel_amax: bool = False) -> Tensor:
    """SU/scheduling_ddim_flip.py:493 `(pred_epsilon - flipped_output).pow(2)`; flip_threshold.py:506 adds
    `.amax(dim=1, keepdim=True)`.  `flipped_back_output` is already flipped back (dims=[2])."""
    u = (eps - flipped_back_output).pow(2)
    return u.amax(dim=1, keepdim=True) if channel_amax else u


def gradient_score_update(predict, input: Tensor, noisy_residual: Tensor, noise_like: Tensor, M: int, alpha_hat_t,
                          gradient_wrt: str = "input"):
    """PU/pipeline_sampler_class_conditional_uncertainty_guided_gradient.py:159-210 `estimate_score_update`:
    returns (pixel_wise_uncertainty, update_scores).  `predict(x)` is predict_model(model, x, t_tensor, y_slice)."""
    from math import sqrt
    noisy_residual = noisy_residual.detach().clone().requires_grad_(gradient_wrt == "score")
    input = input.detach().clone().requires_grad_(gradient_wrt == "input")
    with torch.enable_grad():
        pred_epsilon = predict(input)
        pred_epsilon.mean(dim=0).sum().backward()                                                   # :182-183
        pred_x_0 = (input - sqrt(1 - alpha_hat_t) * noisy_residual) / sqrt(alpha_hat_t)             # :184
        preds = []
        for _ in range(M):
            x_hat_t = sqrt(alpha_hat_t) * pred_x_0 + sqrt(1 - alpha_hat_t) * torch.randn_like(noise_like)   # :187
            preds.append(predict(x_hat_t))
        u = (torch.stack(preds, dim=0) - noisy_residual.unsqueeze(0)).pow(2).mean(dim=0)            # :192
        u.mean(dim=0).sum().backward()                                                              # :193-194
    return u.detach(), (input.grad if gradient_wrt == "input" else noisy_residual.grad)


def column_kth(x: Tensor, k: int) -> Tensor:
    """scripts/compute_threshold_pixel_wise.py:90-100: value at argsort(dim=0)[k] per position."""
    idx = x.argsort(dim=0)[k].unsqueeze(0)
    return x.gather(dim=0, index=idx).squeeze(0)


def fit_pixel_thresholds(uncertainties: Tensor, perc: float) -> Tensor:
    """scripts/compute_threshold_pixel_wise.py:86-100 for uncertainties [N, T_uc, C, H, W]."""
    n = uncertainties.shape[0]
    return torch.stack([column_kth(uncertainties[:, i], int(n * perc)) for i in range(uncertainties.shape[1])], dim=0)


class OracleOutput:
    def __init__(self, prev_sample, pred_original_sample, uncertainty=None, pred_epsilon=None):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample
        self.uncertainty = uncertainty
        self.pred_epsilon = pred_epsilon


class OracleScheduler:
    """CPU restatement of the `step()` of the reference's uncertainty schedulers.

    variant                reference file (SU/)                                     step block
    'zigzag_centered'      scheduling_ddim_uncertainty_zigzag_centered.py            :461-559
    'zigzag'               scheduling_ddim_uncertainty_zigzag.py                     :461-559
    'centered'             scheduling_ddim_uncertainty_centered.py                   :455-548
    'infer_noise'          scheduling_ddim_infer_noise.py                            :455-542
    'mc_dropout'           scheduling_ddim_mc_dropout.py                             :455-556
    'threshold'            scheduling_ddim_uncertainty_threshold.py                  :455-583
    'multiscale'           scheduling_ddim_infer_noise_multiscale_threshold.py       :455-578
    'flip'                 scheduling_ddim_flip.py                                   :440-530
    'flip_threshold'       scheduling_ddim_flip_threshold.py                         :441-583
    'uncertainty_grad'     scheduling_ddim_uncertainty_grad.py                       :455-585 (predict_next=False: the only live mode)
    'mc_dropout_gradient'  scheduling_ddim_mc_dropout_gradient.py                    :440-549
    `predict(x, t)` is the model call (`predict_model`, SU/traits.py:8-18)."""

    def __init__(self, variant: str, predict, M: int, after_step: int, num_steps_uc: int, num_zigzag: int = 4,
                 predict_next: bool = False, beta_schedule: str = "linear", beta_start: float = 1e-4,
                 beta_end: float = 0.02, num_train_timesteps: int = 1000, clip_sample: bool = True,
                 clip_sample_range: float = 1.0, set_alpha_to_one: bool = True, steps_offset: int = 0,
                 prediction_type: str = "epsilon", uncertainty_threshold: float = 1.0,
                 uncertainty_threshold_mode: str = "max", uncertainty_normalize: bool = True, unet=None):
        self.variant, self.predict, self.M = variant, predict, M
        self.after_step, self.num_steps_uc, self.num_zigzag = after_step, num_steps_uc, num_zigzag
        self.predict_next = True if variant in ("zigzag_centered", "zigzag") else predict_next
        self.betas = make_betas(beta_schedule, beta_start, beta_end, num_train_timesteps)
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps, self.steps_offset = num_train_timesteps, steps_offset
        self.clip_sample, self.clip_range, self.prediction_type = clip_sample, clip_sample_range, prediction_type
        self.thr, self.thr_mode, self.normalize = uncertainty_threshold, uncertainty_threshold_mode, uncertainty_normalize
        self.unet = unet
        self.num_inference_steps = None
        self.prompt_embeds = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        self.timesteps = torch.from_numpy(leading_timesteps(self.num_train_timesteps, n, self.steps_offset))
        self.timestep_after_step = self.timesteps[self.after_step].item()
        self.timestep_end_step = self.timesteps[self.after_step + self.num_steps_uc - 1].item()

    def _perturb(self, x0, prev, noise, t):
        if not self.predict_next:
            return perturb_add_noise(x0, noise, self.alphas_cumprod[t])
        return perturb_predict_next(prev, noise, self.betas[t])

    def step(self, model_output, timestep: int, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             variance_noise=None):
        t = timestep
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        c = DDIMCoeffs(self.alphas_cumprod, self.final_alpha_cumprod, t, prev_t, eta)
        in_window = self.timestep_end_step <= t <= self.timestep_after_step
        v = self.variant
        if v == "mc_dropout_gradient":
            # M dropout forwards of the SAME sample under autograd, gradient of the map w.r.t. the sample (:490-515);
            # inside the window x0 is recomputed from the model output WITHOUT clipping (:515)
            u = g = None
            if in_window:
                self.unet.train()
                with torch.enable_grad():
                    sg = sample.detach().clone().requires_grad_(True)
                    u = torch.var(torch.stack([self.predict(sg, t) for _ in range(self.M)], dim=0), dim=0)
                    u.mean(dim=0).sum().backward()
                g = sg.grad
                self.unet.eval()
            noise = None
            if eta > 0:
                noise = variance_noise if variance_noise is not None else torch.randn(model_output.shape)
            prev, x0, eps = ddim_step(model_output, sample, c, self.prediction_type, self.clip_sample, self.clip_range,
                                      eta, noise, use_clipped_model_output)
            if in_window:
                eps = 0.9 * model_output + 0.1 * g
                x0 = (sample - c.sqrt_beta_t * model_output) / c.sqrt_alpha_t
                if use_clipped_model_output:
                    eps = (sample - c.sqrt_alpha_t * x0) / c.sqrt_beta_t
                prev = c.sqrt_alpha_prev * x0 + c.dir_coef * eps
                if eta > 0:
                    prev = prev + c.sigma * noise
            return OracleOutput(prev, x0, u.detach() if u is not None else None, eps if in_window else None)

        if v == "mc_dropout":
            # no best_noise draw; M dropout forwards on the SAME sample happen before the update
            u = None
            if in_window:
                self.unet.train()
                u = variance_unbiased([self.predict(sample, t) for _ in range(self.M)])
                self.unet.eval()
            noise = None
            if eta > 0:
                noise = variance_noise if variance_noise is not None else torch.randn(model_output.shape)
            prev, x0, eps = ddim_step(model_output, sample, c, self.prediction_type, self.clip_sample, self.clip_range,
                                      eta, noise, use_clipped_model_output)
            return OracleOutput(prev, x0, u, eps if in_window else None)

        best_noise = torch.randn_like(sample)  # drawn every step, even for eta == 0
        prev, x0, eps = ddim_step(model_output, sample, c, self.prediction_type, self.clip_sample, self.clip_range,
                                  eta, best_noise, use_clipped_model_output)
        if not in_window:
            return OracleOutput(prev, x0)

        if v in ("flip", "flip_threshold"):
            # one forward on the H-flipped x0, flipped back (flip.py:486-493)
            f = torch.flip(self.predict(torch.flip(x0, dims=[2]), t), dims=[2])
            u = flip_uncertainty(eps, f, channel_amax=(v == "flip_threshold"))
            if v == "flip":
                return OracleOutput(prev, x0, u, eps)
            z = znorm(u) if self.normalize else u                                   # flip_threshold.py:521-522
            w = znorm_threshold_mask(z, self.thr, self.thr_mode)                    # :536-541 ([B,1,H,W], broadcast over C)
            prev2, x02, eps2 = masked_restep(model_output, sample, w, c, self.clip_sample, self.clip_range,
                                             use_clipped_model_output)              # :541-561
            return OracleOutput(prev2, x02, z, eps2)

        if v == "uncertainty_grad":
            # scheduling_ddim_uncertainty_grad.py:518-570: the map is differentiated through the score model w.r.t. eps
            with torch.enable_grad():
                e = eps.detach().clone().requires_grad_(True)
                x0g = (sample - c.sqrt_beta_t * e) / c.sqrt_alpha_t
                sc = []
                for _ in range(self.M):
                    noise = torch.randn_like(x0g)
                    sc.append(self.predict(perturb_add_noise(x0g, noise, self.alphas_cumprod[t]), t))
                u = torch.var(torch.stack(sc, dim=0), dim=0)
                u.mean(dim=0).sum().backward()
            g = e.grad
            eps2 = eps + g * c.alpha_prod_t
            x02 = (sample - c.sqrt_beta_t * model_output) / c.sqrt_alpha_t
            if self.clip_sample:
                x02 = x02.clamp(-self.clip_range, self.clip_range)
            if use_clipped_model_output:
                eps2 = (sample - c.sqrt_alpha_t * x02) / c.sqrt_beta_t
            prev2 = c.sqrt_alpha_prev * x02 + c.dir_coef * eps2
            return OracleOutput(prev2, x02, u.detach(), eps2)

        scores = []
        if v in ("zigzag_centered", "zigzag"):
            for _ in range(self.M):
                x_t1 = x0.clone()
                for j in range(self.num_zigzag):
                    noise = torch.randn_like(x0)
                    base = prev if v == "zigzag_centered" else x_t1
                    x_hat = perturb_predict_next(base, noise, self.betas[t])
                    out = self.predict(x_hat, t)
                    if j != self.num_zigzag - 1:
                        x_t1 = (x_hat - c.sqrt_beta_t * out) / c.sqrt_alpha_t
                scores.append(out)
        else:
            for _ in range(self.M):
                noise = torch.randn_like(x0)
                scores.append(self.predict(self._perturb(x0, prev, noise, t), t))

        if v in ("zigzag_centered", "centered"):
            return OracleOutput(prev, x0, centered_second_moment(scores, eps), eps)
        if v in ("zigzag", "infer_noise"):
            return OracleOutput(prev, x0, variance_unbiased(scores), eps)
        # 'threshold' / 'multiscale': F1b -> F2c -> F4
        z, w, prev2, x02, eps2 = uncertainty_step_znorm(
            scores, model_output, sample, self.thr, self.thr_mode, c, self.clip_sample,
            multiscale=(v == "multiscale"), normalize=self.normalize)
        return OracleOutput(prev2, x02, z, eps2)
