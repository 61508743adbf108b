"""CPU oracle of the `dpm_2_uncertainty_centered` scheduler — TEST INFRASTRUCTURE ONLY (imported by tests/ alone; the product
never imports oracle/).

Restates diffusion_uncertainty/schedulers_uncertainty/scheduling_dpm_2_uncertainty_centered.py in plain torch-CPU fp32, one
tensor operation per reference operation (citations are line numbers of that file), for the configurations the reference can
actually run: algorithm_type "dpmsolver++", solver_order 1 / 2, midpoint / heun, plain sigma spacing, no thresholding.
PINNED: bit-exact against tests/golden/sched_dpm2*.npz, recorded from the unmodified reference (tests/golden/make_golden.py dpm).
"""
import numpy as np
import torch

from .du_oracle import make_betas


class OracleDPMOutput:
    def __init__(self, prev_sample, pred_original_sample=None, uncertainty=None, pred_epsilon=None):
        self.prev_sample, self.pred_original_sample = prev_sample, pred_original_sample
        self.uncertainty, self.pred_epsilon = uncertainty, pred_epsilon


def _alpha_sigma(sigma):
    alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)                                   # :436-440
    return alpha_t, sigma * alpha_t


class OracleDPM2Scheduler:
    def __init__(self, predict, M, after_step, num_steps_uc, solver_order=2, solver_type="midpoint", final_sigmas_type="zero",
                 lower_order_final=True, euler_at_final=False, timestep_spacing="linspace", steps_offset=0,
                 beta_schedule="linear", prediction_type="epsilon", variance_type=None, num_train_timesteps=1000, **ignored):
        self.predict, self.M, self.after_step, self.num_steps_uc = predict, M, after_step, num_steps_uc
        self.solver_order, self.solver_type, self.final_sigmas_type = solver_order, solver_type, final_sigmas_type
        self.lower_order_final, self.euler_at_final = lower_order_final, euler_at_final
        self.spacing, self.steps_offset, self.prediction_type, self.variance_type = timestep_spacing, steps_offset, prediction_type, variance_type
        self.T = num_train_timesteps
        self.betas = make_betas(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.prompt_embeds = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, n):                                                 # :285-365
        last = self.T                                                            # lambda_min_clipped = -inf clips nothing
        if self.spacing == "linspace":
            ts = np.linspace(0, last - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif self.spacing == "leading":
            ts = (np.arange(0, n + 1) * (last // (n + 1))).round()[::-1][:-1].copy().astype(np.int64) + self.steps_offset
        else:
            ts = np.arange(last, 0, -(self.T / n)).round().copy().astype(np.int64) - 1
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy().copy()       # np.array(tensor), :322
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        last_sigma = ((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5 if self.final_sigmas_type == "sigma_min" else 0
        self.sigmas = torch.from_numpy(np.concatenate([sig, [last_sigma]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0
        self.step_index = None
        self.timestep_after_step = int(ts[self.after_step])
        self.timestep_end_step = int(ts[self.after_step + self.num_steps_uc - 1])

    def step(self, model_output, timestep, sample):
        n = len(self.timesteps)
        if self.step_index is None:                                             # :857-871
            hits = (self.timesteps == int(timestep)).nonzero()
            self.step_index = n - 1 if len(hits) == 0 else int(hits[1 if len(hits) > 1 else 0])
        i = self.step_index
        first_at_end = (i == n - 1) and (self.euler_at_final or (self.lower_order_final and n < 15) or self.final_sigmas_type == "zero")
        # convert_model_output :523-541 (data prediction)
        if self.prediction_type == "epsilon":
            if self.variance_type in ("learned", "learned_range"):
                model_output = model_output[:, :3]
            a_i, s_i = _alpha_sigma(self.sigmas[i])
            model_output = (sample - s_i * model_output) / a_i
        elif self.prediction_type == "v_prediction":
            a_i, s_i = _alpha_sigma(self.sigmas[i])
            model_output = a_i * sample - s_i * model_output
        for j in range(self.solver_order - 1):
            self.model_outputs[j] = self.model_outputs[j + 1]
        self.model_outputs[-1] = model_output

        alpha_t, sigma_t = _alpha_sigma(self.sigmas[i + 1])
        alpha_s0, sigma_s0 = _alpha_sigma(self.sigmas[i])
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        if self.solver_order == 1 or self.lower_order_nums < 1 or first_at_end:   # :612-620
            prev = (sigma_t / sigma_s0) * sample - (alpha_t * (torch.exp(-h) - 1.0)) * model_output
        else:                                                                    # :671-700
            alpha_s1, sigma_s1 = _alpha_sigma(self.sigmas[i - 1])
            lambda_s1 = torch.log(alpha_s1) - torch.log(sigma_s1)
            m0, m1 = self.model_outputs[-1], self.model_outputs[-2]
            h_0 = lambda_s0 - lambda_s1
            r0 = h_0 / h
            D0, D1 = m0, (1.0 / r0) * (m0 - m1)
            if self.solver_type == "midpoint":
                prev = ((sigma_t / sigma_s0) * sample - (alpha_t * (torch.exp(-h) - 1.0)) * D0
                        - 0.5 * (alpha_t * (torch.exp(-h) - 1.0)) * D1)
            else:
                prev = ((sigma_t / sigma_s0) * sample - (alpha_t * (torch.exp(-h) - 1.0)) * D0
                        + (alpha_t * ((torch.exp(-h) - 1.0) / h + 1.0)) * D1)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1

        out = OracleDPMOutput(prev)
        if self.timestep_end_step <= int(timestep) <= self.timestep_after_step:   # :955-974
            alpha_prod_t = self.alphas_cumprod[int(timestep)]
            beta_prod_t = 1 - alpha_prod_t
            x0 = (sample - beta_prod_t ** (0.5) * model_output) / alpha_prod_t ** (0.5)
            scores = []
            for _ in range(self.M):
                noise = torch.randn_like(x0)
                x_hat = torch.sqrt(alpha_prod_t) * x0 + torch.sqrt(1 - alpha_prod_t) * noise
                scores.append(self.predict(x_hat, timestep))
            unc = (torch.stack(scores, dim=0) - model_output.unsqueeze(0)).pow(2).mean(dim=0)
            out = OracleDPMOutput(prev, x0, unc, model_output)
        self.step_index += 1
        return out
