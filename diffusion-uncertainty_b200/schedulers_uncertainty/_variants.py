"""The uncertainty-block of every scheduler variant, one small class each on top of
`UncertaintyDDIMCore`.  Each class cites the reference block it reproduces; paths are relative to
/root/reference/diffusion_uncertainty/schedulers_uncertainty/.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import ops
from ..configuration import records_config
from ._core import StepState, UncertaintyDDIMCore


# --------------------------------------------------------------------------------------------------------------------
class ZigZagCentered(UncertaintyDDIMCore):
    """`uncertainty_zigzag_centered` — the BASELINE scheduler (configs 2 and 3).
    scheduling_ddim_uncertainty_zigzag_centered.py:527-549: M x num_zigzag forwards on sqrt(1-beta_t) x_{t-1} +
    sqrt(beta_t) n, the LAST zig's prediction of each chain is kept, u = mean_m (eps_hat_m - eps)^2 (F1a)."""

    predict_next_fixed = True
    centered = True          # False: the `zigzag` sibling (unbiased variance, chain continues from x_t1)

    @records_config
    def __init__(self, *args, num_zigzag: int = 4, **kw):
        super().__init__(*args, **kw)
        self.num_zigzag = num_zigzag

    def _chain_base(self, st: StepState, x_t1: torch.Tensor) -> torch.Tensor:
        return st.prev

    def _uncertainty_block(self, st: StepState) -> torch.Tensor:
        h = st.host
        # x_t1 = (x_hat - sqrt(1-abar) eps_hat) / sqrt(abar): the DDIM x0 expression without clipping
        c_x0 = ops.make_coeffs(h["sqrt_alpha_t"], h["sqrt_beta_t"], 0.0, 0.0, clip_sample=False)
        chain_used = (not self.predict_next) or (not self.centered)
        scores: List[torch.Tensor] = []
        for _ in range(self.M):
            x_t1 = st.x0   # the reference clones; nothing here writes in place
            out = None
            for j in range(self.num_zigzag):
                if self.predict_next:
                    a, b = float(torch.sqrt(1 - self.betas[st.t])), float(torch.sqrt(self.betas[st.t]))
                    x_hat = ops.perturb_fresh(self._chain_base(st, x_t1), a, b, noise_like=st.x0)   # noise drawn in the kernel
                else:
                    x_hat = self.add_noise(x_t1, torch.randn_like(st.x0), st.t)
                x_hat = self.scale_model_input(x_hat, st.t)
                out = self.predict_model(x_hat, st.t)
                if j != self.num_zigzag - 1 and chain_used:
                    # (in the centred variant with predict_next the reference computes x_t1 and never reads it: skipped)
                    x_t1 = ops.ddim_step(out, x_hat, c_x0, want_prev=False, want_x0=True)[1]
            scores.append(out)
        if self.centered:
            return self._reduce(scores, "centered", center=st.eps)
        return self._reduce(scores, "var")


class ZigZag(ZigZagCentered):
    """scheduling_ddim_uncertainty_zigzag.py:527-549: chain runs from x_t1 (not x_{t-1}); u = torch.var over M (F1b)."""

    centered = False

    def _chain_base(self, st: StepState, x_t1: torch.Tensor) -> torch.Tensor:
        return x_t1


# --------------------------------------------------------------------------------------------------------------------
class Centered(UncertaintyDDIMCore):
    """`uncertainty_centered` (all config/generation/*.yaml).  scheduling_ddim_uncertainty_centered.py:522-539: M
    re-noised forwards, u = mean_m (eps_hat_m - eps)^2 (F1a)."""

    def _uncertainty_block(self, st: StepState) -> torch.Tensor:
        return self._reduce(self._perturbed_scores(st), "centered", center=st.eps)


class InferNoise(UncertaintyDDIMCore):
    """scheduling_ddim_infer_noise.py:515-533: as Centered, u = torch.var over M (F1b)."""

    def _uncertainty_block(self, st: StepState) -> torch.Tensor:
        return self._reduce(self._perturbed_scores(st), "var")


class UncertaintyImage(UncertaintyDDIMCore):
    """`uncertainty_image`.  scheduling_ddim_uncertainty_image.py:515-533, 545-554: each perturbed prediction is mapped to
    an x_{t-1} candidate (DDIM update of the perturbed input, no clipping, no noise); u = torch.var of the candidates."""

    def _uncertainty_block(self, st: StepState) -> torch.Tensor:
        h = st.host
        c = ops.make_coeffs(h["sqrt_alpha_t"], h["sqrt_beta_t"], h["sqrt_alpha_prev"], h["dir_coef"], clip_sample=False)
        cands = []
        for _ in range(self.M):
            x_hat = self._perturbed_input(st, st.x0)
            out = self.predict_model(x_hat, st.t)
            cands.append(ops.ddim_step(out, x_hat, c, want_prev=True, want_x0=False)[0])
        return self._reduce(cands, "var")


class CenteredD(UncertaintyDDIMCore):
    """`uncertainty_centered_d`.  scheduling_ddim_uncertainty_centered_d.py:523-541: re-noise over a d-step jump
    (true_alpha = abar_t / abar_{t+d}), model called at `ending_step`, F1a.  The reference indexes alphas_cumprod with a
    STEP index (:532) — reproduced."""

    @records_config
    def __init__(self, *args, uncertainty_distance: int = 20, **kw):
        super().__init__(*args, **kw)
        self.uncertainty_distance = uncertainty_distance

    def set_timesteps(self, num_inference_steps: int, device=None):
        super().set_timesteps(num_inference_steps, device)
        ordered = sorted(self._host_timesteps, reverse=True)
        self.timestep_index = {t: i for i, t in enumerate(ordered)}                  # :392
        self.index_timestep = {i: t for t, i in self.timestep_index.items()}         # :394

    def _uncertainty_block(self, st: StepState) -> torch.Tensor:
        idx = self.timestep_index[st.t]
        dist = min(self.uncertainty_distance, len(self.timesteps) - idx - 1)
        ending_step = idx + dist - 1
        end_alpha = 1 if self.index_timestep[idx + dist] == 0 else self.alphas_cumprod[idx + dist]
        true_alpha = st.host["alpha_prod_t"] / end_alpha
        sa, sb = float(true_alpha ** 0.5), float((1 - true_alpha) ** 0.5)
        # x_t_next = (sample - sqrt(1-a) eps) / sqrt(a);  sample_hat = x_t_next sqrt(a) + sqrt(1-a) n
        c = ops.make_coeffs(sa, sb, 0.0, 0.0, clip_sample=False)
        x_next = ops.ddim_step(st.eps, st.sample, c, want_prev=False, want_x0=True)[1]
        scores = []
        for _ in range(self.M):
            scores.append(self.predict_model(ops.perturb_fresh(x_next, sa, sb, noise_like=st.x0), ending_step))
        return self._reduce(scores, "centered", center=st.eps)


class ActivationNoise(UncertaintyDDIMCore):
    """`uncertainty` / `uncertainty_original`.  scheduling_ddim_uncertainty.py:519-542: the SAME sample is forwarded M
    times while forward hooks add N(0, 0.01^2) to four named ADM blocks; F1a about eps."""

    HOOKED = ("input_blocks.8.0", "output_blocks.12.0", "output_blocks.1.0", "output_blocks.4.0")

    @staticmethod
    def _add_gaussian_noise(module, inputs, output):   # :37-40
        return output + torch.randn_like(output) * 0.01

    def _uncertainty_block(self, st: StepState) -> torch.Tensor:
        hooks = [m.register_forward_hook(self._add_gaussian_noise) for name, m in self.unet.named_modules() if name in self.HOOKED]
        try:
            scores = [self.predict_model(st.sample, st.t) for _ in range(self.M)]
        finally:
            for hk in hooks:
                hk.remove()
        return self._reduce(scores, "centered", center=st.eps)


# --------------------------------------------------------------------------------------------------------------------
class MCDropout(UncertaintyDDIMCore):
    """Default scheduler (`mc_dropout`).  scheduling_ddim_mc_dropout.py:498-556: `unet.train()`, M forwards of the SAME
    sample with dropout active, `unet.eval()`, u = torch.var over M (F1b).  No `best_noise`; eta > 0 uses `variance_noise`;
    `pred_original_sample`, `score` and `pred_epsilon` are returned as CPU tensors (:551-554)."""

    draws_best_noise = False
    eta_uses_best_noise = False
    host_copies = True
    class_conditioned = False

    @records_config
    def __init__(self, *args, uncertainty_scale: float = 0.9, dropout: Optional[float] = None, **kw):
        super().__init__(*args, **kw)
        self.uncertainty_scale = uncertainty_scale
        self.first_step = True
        if dropout is not None and self.unet is not None:        # :281-285
            for module in self.unet.modules():
                if isinstance(module, torch.nn.Dropout):
                    module.p = dropout

    def _before_update(self, st: StepState) -> torch.Tensor:
        self.unet.train()
        try:
            scores = [self.predict_model(st.sample, st.t) for _ in range(self.M)]
            u = self._reduce(scores, "var")
            if self.first_step:
                self.first_step = False
                if not any(isinstance(m, torch.nn.Dropout) for m in self.unet.modules()):
                    raise ValueError("The model should have a dropout layer for MC Dropout")
                for m in self.unet.modules():
                    if isinstance(m, torch.nn.Dropout):
                        assert m.training is True
                        assert m.p > 0.0, f"Expected dropout rate > 0, got {m.p}"
        finally:
            self.unet.eval()
        return u

    def _finish_output(self, out, st: StepState, window: bool):
        if self.host_copies:
            out.pred_original_sample = st.x0.cpu()
            out.score = st.eps.cpu()
            if window:
                out.pred_epsilon = st.eps.cpu()
        else:
            out.score = st.eps


# --------------------------------------------------------------------------------------------------------------------
class ZNormThreshold(UncertaintyDDIMCore):
    """fid:`uncertainty_threshold`.  scheduling_ddim_uncertainty_threshold.py:524-574: F1b -> whole-batch z-norm (F2c) ->
    scalar threshold mask -> eps * mask -> x0 from the UNMASKED prediction -> x_{t-1} again (F4).  `output.uncertainty`
    is the z-normalised map, as in the reference (the variable is reassigned at :540)."""

    multiscale = False

    @records_config
    def __init__(self, *args, uncertainty_threshold: float = 1.0, uncertainty_threshold_mode: str = "max",
                 uncertainty_normalize: bool = True, **kw):
        super().__init__(*args, **kw)
        self.uncertainty_threshold = uncertainty_threshold
        self.uncertainty_threshold_mode = uncertainty_threshold_mode
        self.uncertainty_normalize = uncertainty_normalize

    def _check_prediction_type(self):
        assert self.config.prediction_type == "epsilon", \
            f"Actually implemented only for prediction type epsilon - actual {self.config.prediction_type}"

    def _stats(self, u: torch.Tensor) -> torch.Tensor:
        """[mean, unbiased std, count, M2] over the WHOLE batch; hook for the multi-GPU merge (distributed.py)."""
        return ops.znorm_stats(u)

    def _raw_map(self, st: StepState) -> torch.Tensor:
        """the un-normalised map: torch.var over the M perturbed predictions (F1b)"""
        return ops.moments(self._perturbed_scores(st), mode="var")

    def _uncertainty_block(self, st: StepState) -> torch.Tensor:
        u = self._raw_map(st)
        stats = self._stats(u) if self.uncertainty_normalize else None
        mode = "multiscale" if self.multiscale else ("max" if self.uncertainty_threshold_mode == "max" else "min")
        z, w = ops.znorm_weights(u, stats, mode=mode, thr=float(self.uncertainty_threshold), normalize=self.uncertainty_normalize)
        sink = self._map_out(z)
        if sink is not None:
            ops.accumulate_slot(z, sink)
            z = sink
        # F4: the masked re-step; eta noise is NOT re-added (reference :566-571)
        if self.config.thresholding:       # dynamic thresholding replaces the clamp of x0 (:557-558): the general re-step
            guided = ops.guided_step(st.model_output, None, None, guidance="weights", mask=w, want_eps=True)["eps"]
            self._restep_general(st, st.model_output, guided)
            return z
        c = ops.make_coeffs(st.host["sqrt_alpha_t"], st.host["sqrt_beta_t"], st.host["sqrt_alpha_prev"], st.host["dir_coef"],
                            clip_sample=bool(self.config.clip_sample), clip_range=float(self.config.clip_sample_range),
                            use_clipped_model_output=st.use_clipped)
        r = ops.guided_step(st.model_output, st.sample, c, guidance="weights", mask=w, want_prev=True, want_x0=True, want_eps=True)
        st.prev, st.x0, st.eps = r["prev"], r["x0"], r["eps"]
        return z


class MultiscaleThreshold(ZNormThreshold):
    """scheduling_ddim_infer_noise_multiscale_threshold.py:520-569: weights 0.8 / 0.9 / 1.0 on the z bands (-3,-2), (-2,-1),
    [-1, inf); 0 elsewhere."""

    multiscale = True

    @records_config
    def __init__(self, *args, uncertainty_normalize: bool = True, **kw):
        super().__init__(*args, uncertainty_normalize=uncertainty_normalize, **kw)


# --------------------------------------------------------------------------------------------------------------------
class Flip(UncertaintyDDIMCore):
    """`flip`.  scheduling_ddim_flip.py:486-493: ONE extra forward on the H-flipped x0 (`torch.flip(x0, dims=[2])`), the
    prediction flipped back, u = (eps - flipped)^2 — no M axis.  Both flips and the squared difference run in
    du_flip_h / du_flip_sqdiff (the flip back is folded into the difference kernel's addressing)."""

    channel_amax = False
    host_copies = True     # the reference's flip scheduler returns x0, score (and, in the window, the map and pred_epsilon) as CPU
                           # tensors (scheduling_ddim_flip.py:523-526); the sampling loops of this package switch the copies off

    def _finish_output(self, out, st: StepState, window: bool):
        if self.host_copies:
            out.pred_original_sample = st.x0.cpu()
            out.score = st.eps.cpu()
            if window:
                out.uncertainty = out.uncertainty.cpu()
                out.pred_epsilon = st.eps.cpu()
        else:
            out.score = st.eps

    def _flip_map(self, st: StepState, out=None) -> torch.Tensor:
        flipped_output = self.predict_model(ops.flip_h(st.x0), st.t)
        eps = st.model_output if self.config.prediction_type == "epsilon" else st.eps
        return ops.flip_sqdiff(eps, flipped_output, channel_amax=self.channel_amax, out=out)

    def _uncertainty_block(self, st: StepState) -> torch.Tensor:
        return self._flip_map(st, out=self._map_out(st.x0, torch.float32))


class FlipThreshold(ZNormThreshold):
    """fid:`flip_threshold`.  scheduling_ddim_flip_threshold.py:497-561: flip map -> amax over channels ([B,1,H,W]) ->
    whole-batch z-norm -> scalar threshold mask -> eps * mask (broadcast over channels) -> x0 from the UNMASKED
    prediction -> x_{t-1} again (F4)."""

    @records_config
    def __init__(self, *args, uncertainty_scale: float = 0.9, **kw):
        super().__init__(*args, **kw)
        self.uncertainty_scale = uncertainty_scale

    def _raw_map(self, st: StepState) -> torch.Tensor:
        flipped_output = self.predict_model(ops.flip_h(st.x0), st.t)
        return ops.flip_sqdiff(st.model_output, flipped_output, channel_amax=True)


# --------------------------------------------------------------------------------------------------------------------
class UncertaintyGrad(UncertaintyDDIMCore):
    """fid:`uncertainty_grad`.  scheduling_ddim_uncertainty_grad.py:518-570: the map (torch.var over M re-noised forwards) is
    differentiated through the score model with respect to eps; eps' = eps + grad * abar_t; x0 from the UNGUIDED
    prediction; x_{t-1} recomputed (eta noise not re-added).  The reduction and its backward are du_moments /
    du_moments_backward (ops.moments_autograd), the blend + DDIM update one du_guided_step launch; the model's own
    backward is torch autograd.  Only predict_next=False is live in the reference (with predict_next the gradient is None
    and the reference's assert fires): the same AssertionError is raised here."""

    def _uncertainty_block(self, st: StepState) -> torch.Tensor:
        assert not self.predict_next, "grad_pred_epsilon is None: the reference asserts when predict_next=True"
        h = st.host
        sb, sa = h["sqrt_beta_t"], h["sqrt_alpha_t"]
        with torch.enable_grad():
            e = st.eps.detach().clone().requires_grad_(True)
            x0 = (st.sample - sb * e) / sa                      # :522, traced for autograd (model-input side)
            scores = []
            for _ in range(self.M):
                noise = torch.randn_like(x0)
                x_hat = self.scale_model_input(self.add_noise(x0, noise, st.t), st.t)
                scores.append(self.predict_model(x_hat, st.t))
            u = ops.moments_autograd(scores, "var")
            u.mean(dim=0).sum().backward()
        g = e.grad
        assert g is not None
        if self.config.thresholding or st.eps is not st.model_output:
            # x0 comes from model_output (:552) while the gradient is added to pred_epsilon — which differs from model_output once
            # use_clipped_model_output re-derived it — and dynamic thresholding replaces the clamp: the general re-step
            guided = ops.guided_step(st.eps, None, None, guidance="grad_add", aux=g, lam=float(h["alpha_prod_t"]), want_eps=True)["eps"]
            self._restep_general(st, st.model_output, guided)
        else:
            c = ops.make_coeffs(h["sqrt_alpha_t"], h["sqrt_beta_t"], h["sqrt_alpha_prev"], h["dir_coef"],
                                clip_sample=bool(self.config.clip_sample), clip_range=float(self.config.clip_sample_range),
                                use_clipped_model_output=st.use_clipped)
            r = ops.guided_step(st.eps, st.sample, c, guidance="grad_add", aux=g, lam=float(h["alpha_prod_t"]), x0_unguided=True,
                                want_prev=True, want_x0=True, want_eps=True)
            st.prev, st.x0, st.eps = r["prev"], r["x0"], r["eps"]
        u = u.detach()
        sink = self._map_out(u)
        if sink is not None:
            ops.accumulate_slot(u, sink)
            u = sink
        return u


class MCDropoutGradient(MCDropout):
    """fid:`mc_dropout_gradient`.  scheduling_ddim_mc_dropout_gradient.py:490-549: M dropout forwards of the SAME sample under
    autograd, gradient of the map with respect to the sample, eps' = 0.9 eps + 0.1 grad, x0 recomputed from the model
    output WITHOUT clipping, then the ordinary x_{t-1} (eta noise from `variance_noise`)."""

    def _before_update(self, st: StepState) -> torch.Tensor:
        self.unet.train()
        try:
            with torch.enable_grad():
                sg = st.sample.detach().clone().requires_grad_(True)
                scores = [self.predict_model(sg, st.t) for _ in range(self.M)]
                u = ops.moments_autograd(scores, "var")
                u.mean(dim=0).sum().backward()
            self._grad_sample = sg.grad
            assert self._grad_sample is not None
        finally:
            self.unet.eval()
        u = u.detach()
        sink = self._map_out(u)
        if sink is not None:
            ops.accumulate_slot(u, sink)
            u = sink
        return u

    def _ddim_update(self, st: StepState, noise):
        if not self.in_window(st.t):
            return super()._ddim_update(st, noise)
        h = st.host
        c = ops.make_coeffs(h["sqrt_alpha_t"], h["sqrt_beta_t"], h["sqrt_alpha_prev"], h["dir_coef"], clip_sample=False,
                            use_clipped_model_output=st.use_clipped)
        r = ops.guided_step(st.model_output, st.sample, c, guidance="lincomb", aux=self._grad_sample, post_M=0.9, lam=0.1,
                            x0_unguided=True, want_prev=True, want_x0=True, want_eps=True)
        prev = r["prev"]
        if noise is not None:
            prev = ops.perturb(prev, noise, 1.0, h["sigma"])
        st.prev, st.x0, st.eps = prev, r["x0"], r["eps"]
