"""Tensor-level wrappers over the C ABI (include/du_b200.h).  PyTorch is plumbing only here: it owns the
device memory and the stream; every arithmetic step of the path runs in libdu_b200.so.

All functions require CUDA tensors — there is no CPU fallback (BASELINE.json north_star).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional, Sequence, Tuple, Union

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.float16: L.F16, torch.bfloat16: L.BF16}
_MODES = {"var": L.MOM_VAR_UNBIASED, "centered": L.MOM_CENTERED, "var_with_center": L.MOM_VAR_WITH_CENTER,
          "raw": L.MOM_RAW, "std": L.MOM_STD_UNBIASED, "partial": L.MOM_PARTIAL_M2}
_PRED = {"epsilon": L.PRED_EPSILON, "sample": L.PRED_SAMPLE, "v_prediction": L.PRED_V}
_GUIDE = {"none": L.GUIDE_NONE, "posterior": L.GUIDE_POSTERIOR, "grad_blend": L.GUIDE_GRAD_BLEND,
          "grad_add": L.GUIDE_GRAD_ADD, "weights": L.GUIDE_WEIGHTS, "lincomb": L.GUIDE_LINCOMB, "sign_add": L.GUIDE_SIGN_ADD,
          "mul_blend": L.GUIDE_MUL_BLEND}
_ZN = {"max": L.ZN_BELOW, "min": L.ZN_ABOVE, "below": L.ZN_BELOW, "above": L.ZN_ABOVE, "multiscale": L.ZN_MULTISCALE}

# launches issued through this module since import (bench.py reports it as gpu_launches)
launch_count = 0
last_step_path = ""        # "fused" / "unfused": which implementation the last uncertainty_step() call took
_tls = threading.local()   # device index last handed to du_set_device by THIS thread (the library keeps it thread-local too)


def _count(n=1):
    global launch_count
    launch_count += n


def _require_cuda(t: torch.Tensor, name: str):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: the uncertainty path has no CPU fallback, move it to a CUDA device")
    if t.dtype not in _DT:
        raise RuntimeError(f"{name}: dtype {t.dtype} is not supported (float32, float16, bfloat16)")


def _stream(t: torch.Tensor):
    """Stream of t's device for the launch that follows, and tell the library which device that launch belongs to.  The
    library switches to it only for the duration of the call (csrc/du_common.cuh DeviceGuard), so the caller's
    torch.cuda.current_device() never changes and a later torch.cuda.set_device / another thread cannot make this stale:
    the cached index mirrors the library's own thread-local value, which only this function writes."""
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if getattr(_tls, "device", None) != idx:
        L.check(L.load().du_set_device(idx))
        _tls.device = idx
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class Rows:
    """A tensor seen as B rows of n contiguous elements (the C ABI's 'rows view')."""
    __slots__ = ("t", "ptr", "stride", "dt", "B", "n")

    def __init__(self, t: torch.Tensor, name: str = "tensor", batch: bool = True):
        _require_cuda(t, name)
        if t.dim() == 0:
            t = t.reshape(1, 1)
        if not batch:
            t = t.reshape(1, -1) if t.is_contiguous() else t.contiguous().reshape(1, -1)
        B = t.shape[0]
        n = 1
        for d_ in t.shape[1:]:
            n *= int(d_)
        # rows must be internally contiguous; the batch stride may be anything (channel-slice views)
        if B > 0 and n > 0 and not t[0].is_contiguous():
            t = t.contiguous()
        self.t = t
        self.ptr = C.c_void_p(t.data_ptr())
        self.stride = t.stride(0) if (B > 1 and t.dim() > 0) else n
        self.dt = _DT[t.dtype]
        self.B, self.n = B, n


def _same_rows(a: Rows, b: Rows, what: str):
    if a.B != b.B or a.n != b.n:
        raise ValueError(f"{what}: shape mismatch ({a.B}x{a.n} vs {b.B}x{b.n})")


NULL = C.c_void_p(0)


# ------------------------------------------------------------------------------------------------ F1
def moments(scores: Union[Sequence[torch.Tensor], torch.Tensor], center: Optional[torch.Tensor] = None,
            mode: str = "var", out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
            return_mean: bool = False):
    """Reduce M score tensors [B,...] over the M axis without stacking them (du_moments).
    mode: 'var' | 'centered' | 'var_with_center' | 'raw' | 'std' | 'partial'.
    `out` may be a (strided) slot view of the accumulation buffer."""
    if isinstance(scores, torch.Tensor):
        scores = list(scores.unbind(0))
    if len(scores) == 0:
        raise ValueError("moments: need at least one score tensor")
    if len(scores) > L.DU_MAX_M:
        raise ValueError(f"moments: M={len(scores)} exceeds DU_MAX_M={L.DU_MAX_M}")
    rows = [Rows(s, f"scores[{k}]") for k, s in enumerate(scores)]
    r0 = rows[0]
    for r in rows[1:]:
        _same_rows(r0, r, "moments")
        if r.dt != r0.dt:
            raise RuntimeError("moments: all score tensors must share one dtype")
    strides = {r.stride for r in rows}
    if len(strides) != 1:  # mixed layouts: normalise
        rows = [Rows(r.t.contiguous(), "scores") for r in rows]
        r0 = rows[0]
    shape = scores[0].shape
    crow = None
    if center is not None:
        crow = Rows(center, "center")
        _same_rows(r0, crow, "moments(center)")
    if out is None:
        out = torch.empty(shape, device=scores[0].device, dtype=out_dtype or torch.float32)
    orow = Rows(out, "out")
    if orow.t is not out:
        raise ValueError("moments: `out` rows must be contiguous")
    _same_rows(r0, orow, "moments(out)")
    mean = torch.empty(shape, device=scores[0].device, dtype=torch.float32) if return_mean else None
    if out.numel() == 0:
        return (out, mean) if return_mean else out
    ptrs = (C.c_void_p * len(rows))(*[r.ptr for r in rows])
    rc = L.load().du_moments(ptrs, len(rows), r0.stride, r0.dt,
                             crow.ptr if crow else NULL, crow.stride if crow else 0, crow.dt if crow else 0,
                             _MODES[mode], r0.B, r0.n, orow.ptr, orow.stride, orow.dt,
                             C.c_void_p(mean.data_ptr()) if mean is not None else NULL, r0.n, _stream(out))
    L.check(rc)
    _count()
    return (out, mean) if return_mean else out


def moments_backward(scores: Sequence[torch.Tensor], grad_u: torch.Tensor, mode: str = "var", center: Optional[torch.Tensor] = None,
                     need: Optional[Sequence[bool]] = None, need_center: bool = False):
    """Backward of `moments` (du_moments_backward): returns ([grad_scores[m] or None], grad_center or None), fp32.
    `need[m]` = False skips the gradient of sample m."""
    rows = [Rows(s_.detach(), f"scores[{k}]") for k, s_ in enumerate(scores)]
    r0 = rows[0]
    for r in rows[1:]:
        _same_rows(r0, r, "moments_backward")
    if len({(r.stride, r.dt) for r in rows}) != 1:
        rows = [Rows(r.t.float().contiguous(), "scores") for r in rows]
        r0 = rows[0]
    g = Rows(grad_u.detach(), "grad_u"); _same_rows(r0, g, "moments_backward(grad_u)")
    crow = None
    if center is not None:
        crow = Rows(center.detach(), "center"); _same_rows(r0, crow, "moments_backward(center)")
    shape, dev = scores[0].shape, scores[0].device
    need = [True] * len(rows) if need is None else list(need)
    grads = [torch.empty(shape, device=dev, dtype=torch.float32) if nd else None for nd in need]
    gc = torch.empty(shape, device=dev, dtype=torch.float32) if (need_center and crow is not None) else None
    if r0.B * r0.n == 0:
        return grads, gc
    sp = (C.c_void_p * len(rows))(*[r.ptr for r in rows])
    gp = (C.c_void_p * len(rows))(*[C.c_void_p(t.data_ptr()) if t is not None else NULL for t in grads])
    L.check(L.load().du_moments_backward(sp, len(rows), r0.stride, r0.dt, crow.ptr if crow else NULL, crow.stride if crow else 0,
                                         crow.dt if crow else 0, _MODES[mode], g.ptr, g.stride, g.dt, r0.B, r0.n, gp, r0.n,
                                         L.F32, C.c_void_p(gc.data_ptr()) if gc is not None else NULL, r0.n, _stream(grad_u)))
    _count()
    return grads, gc


class _MomentsFn(torch.autograd.Function):
    """`moments` as a differentiable op: forward = du_moments, backward = du_moments_backward.  The schedulers that
    differentiate the map through the score model (SU/scheduling_ddim_uncertainty_grad.py:536-538 and relatives) call it
    through `moments_autograd`; the model's own backward stays torch autograd."""

    @staticmethod
    def forward(ctx, mode, center, *scores):
        ctx.mode = mode
        ctx.has_center = center is not None
        ctx.save_for_backward(*([center] if center is not None else []), *scores)
        return moments([s_.detach() for s_ in scores], center=center.detach() if center is not None else None, mode=mode,
                       out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, grad_u):
        saved = ctx.saved_tensors
        center = saved[0] if ctx.has_center else None
        scores = saved[1:] if ctx.has_center else saved
        need = list(ctx.needs_input_grad[2:])
        need_c = ctx.has_center and ctx.needs_input_grad[1]
        grads, gc = moments_backward(scores, grad_u.contiguous(), ctx.mode, center, need=need, need_center=need_c)
        grads = [g_.to(s_.dtype) if g_ is not None else None for g_, s_ in zip(grads, scores)]
        return (None, gc.to(center.dtype) if gc is not None else None, *grads)


def moments_autograd(scores: Sequence[torch.Tensor], mode: str = "var", center: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Differentiable F1 reduction ('var' | 'std' | 'centered' | 'var_with_center'), fp32 map."""
    return _MomentsFn.apply(mode, center, *scores)


class _X0Fn(torch.autograd.Function):
    """x0 = (x - sb * eps) / sa, differentiable in eps: forward du_ddim_step (the reference's three rounded operations),
    backward -(sb / sa) * g in one launch.  The threshold-guided loops keep this on the autograd graph because the gradient with
    respect to eps flows through the re-noised model inputs (generate_samples.py:806, PU/uncertainty_guidance.py:89)."""

    @staticmethod
    def forward(ctx, eps, x, sa, sb):
        ctx.k = -float(sb) / float(sa)
        c = make_coeffs(float(sa), float(sb), 0.0, 0.0, clip_sample=False)
        return ddim_step(eps.detach(), x.detach(), c, want_prev=False, want_x0=True)[1]

    @staticmethod
    def backward(ctx, g):
        return (scale(g.contiguous(), ctx.k) if ctx.needs_input_grad[0] else None), None, None, None


def x0_autograd(eps: torch.Tensor, x: torch.Tensor, sa: float, sb: float) -> torch.Tensor:
    return _X0Fn.apply(eps, x, sa, sb)


def mask_greater(u: torch.Tensor, thr: torch.Tensor, higher: bool = True) -> torch.Tensor:
    """(u > thr).float() with one threshold PER ELEMENT (thr has u's shape) — du_tensor_threshold_mask over the flattened tensors."""
    if tuple(u.shape) != tuple(thr.shape):
        raise ValueError("mask_greater: u and thr must have the same shape")
    u = u if u.is_contiguous() else u.contiguous()
    thr = thr if thr.is_contiguous() else thr.contiguous()
    return tensor_threshold_mask(u.reshape(1, -1), thr.reshape(-1), higher=higher).reshape(u.shape)


# ------------------------------------------------------------------------------------------------ N4 flips
def _chw(x: torch.Tensor):
    if x.dim() != 4:
        raise ValueError("flip ops expect [B, C, H, W] tensors")
    return int(x.shape[1]), int(x.shape[2]), int(x.shape[3])


def flip_h(x: torch.Tensor) -> torch.Tensor:
    """torch.flip(x, dims=[2]) (du_flip_h) — the model input of the flipped forward, SU/scheduling_ddim_flip.py:487."""
    r = Rows(x, "x")
    Cc, H, W = _chw(x)
    out = torch.empty(x.shape, device=x.device, dtype=x.dtype)
    if out.numel() == 0:
        return out
    L.check(L.load().du_flip_h(r.ptr, r.stride, r.dt, r.B, Cc, H, W, C.c_void_p(out.data_ptr()), r.n, _DT[out.dtype], _stream(x)))
    _count()
    return out


def flip_sqdiff(eps: torch.Tensor, flipped_output: torch.Tensor, channel_amax: bool = False,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(eps - flip_h(flipped_output))^2, optionally followed by amax over channels (keepdim) — du_flip_sqdiff,
    SU/scheduling_ddim_flip.py:488-493, SU/scheduling_ddim_flip_threshold.py:504-506.  fp32 result."""
    e, f = Rows(eps, "eps"), Rows(flipped_output, "flipped_output")
    _same_rows(e, f, "flip_sqdiff")
    Cc, H, W = _chw(eps)
    shape = (eps.shape[0], 1, H, W) if channel_amax else tuple(eps.shape)
    if out is None:
        out = torch.empty(shape, device=eps.device, dtype=torch.float32)
    orow = Rows(out, "out")
    if orow.t is not out or out.dtype != torch.float32 or tuple(out.shape) != tuple(shape):
        raise ValueError("flip_sqdiff: `out` must be a float32 tensor of the result shape with contiguous rows")
    if out.numel() == 0:
        return out
    L.check(L.load().du_flip_sqdiff(e.ptr, e.stride, e.dt, f.ptr, f.stride, f.dt, e.B, Cc, H, W, int(channel_amax), orow.ptr, orow.stride,
                                    _stream(eps)))
    _count()
    return out


# ------------------------------------------------------------------------------------------------ N2 / N3
def column_kth(x: torch.Tensor, k: int) -> torch.Tensor:
    """k-th smallest (0-based) over dim 0 of x [N, ...] per remaining position, NaN last (du_column_kth):
    `x.gather(0, x.argsort(dim=0)[k].unsqueeze(0)).squeeze(0)` of scripts/compute_threshold_pixel_wise.py:90-100.
    x may be a [:, i] slice of the [N, T_uc, C, H, W] accumulated maps (row stride = T_uc * C*H*W)."""
    r = Rows(x, "x")
    if r.B < 1 or not (0 <= int(k) < r.B):
        raise IndexError(f"column_kth: k={k} out of range for {r.B} samples")
    out = torch.empty(x.shape[1:], device=x.device, dtype=x.dtype)
    L.check(L.load().du_column_kth(r.ptr, r.dt, r.B, r.n, r.stride, int(k), C.c_void_p(out.data_ptr()), _stream(x)))
    _count()
    return out


def fit_pixel_thresholds(uncertainties: torch.Tensor, perc: float) -> torch.Tensor:
    """The per-timestep / per-pixel thresholds of scripts/compute_threshold_pixel_wise.py:86-100: for uncertainties
    [N, T_uc, C, H, W] returns [T_uc, C, H, W], the int(N * perc)-th smallest value over the samples at every timestep."""
    N, T_ = int(uncertainties.shape[0]), int(uncertainties.shape[1])
    k = int(N * perc)
    return torch.stack([column_kth(uncertainties[:, i], k) for i in range(T_)], dim=0)


def row_sum(x: torch.Tensor) -> torch.Tensor:
    """x.sum(dim=(1, 2, ...)) per sample as fp32 (du_row_sum) — scripts/uncertainty_benchmark_imagenet.py:314."""
    _require_cuda(x, "x")
    if x.shape[0] == 0 or x.numel() == 0:
        return torch.zeros(x.shape[0], device=x.device, dtype=torch.float32)
    flat = x.reshape(x.shape[0], -1) if x.is_contiguous() else x.contiguous().reshape(x.shape[0], -1)
    r = Rows(flat, "x")
    out = torch.zeros(r.B, device=x.device, dtype=torch.float32)
    if r.B == 0 or r.n == 0:
        return out
    L.check(L.load().du_row_sum(r.ptr, r.stride, r.dt, r.B, r.n, C.c_void_p(out.data_ptr()), _stream(x)))
    _count()
    return out


def slot_sum(x: torch.Tensor) -> torch.Tensor:
    """x.sum(dim=1) of the accumulated maps [B, T_uc, ...] as fp32 (du_slot_sum) — scripts/compute_ause.py:128."""
    _require_cuda(x, "x")
    if x.dim() < 3:
        raise ValueError("slot_sum expects [B, T, ...]")
    x = x if x.is_contiguous() else x.contiguous()
    B, T_ = int(x.shape[0]), int(x.shape[1])
    n = 1
    for d_ in x.shape[2:]:
        n *= int(d_)
    out = torch.zeros((B,) + tuple(x.shape[2:]), device=x.device, dtype=torch.float32)
    if out.numel() == 0 or T_ == 0:
        return out
    L.check(L.load().du_slot_sum(C.c_void_p(x.data_ptr()), T_ * n, n, _DT[x.dtype], B, T_, n, C.c_void_p(out.data_ptr()), n, _stream(x)))
    _count()
    return out


def moments_merge(means: Sequence[torch.Tensor], m2s: Sequence[torch.Tensor], counts: Sequence[int], mode: str = "var",
                  return_mean: bool = False):
    """Chan-merge R per-rank partial moments (du_moments_merge)."""
    R = len(m2s)
    m2s = [m.contiguous() for m in m2s]
    for m in m2s:
        _require_cuda(m, "m2")
    N = m2s[0].numel()
    out = torch.empty_like(m2s[0], dtype=torch.float32)
    mean_out = torch.empty_like(out) if return_mean else None
    mp = None
    if mode != "centered":
        means = [m.contiguous() for m in means]
        mp = (C.c_void_p * R)(*[C.c_void_p(m.data_ptr()) for m in means])
    qp = (C.c_void_p * R)(*[C.c_void_p(m.data_ptr()) for m in m2s])
    cp = (C.c_int * R)(*[int(c) for c in counts])
    rc = L.load().du_moments_merge(mp, qp, cp, R, _MODES[mode], N, C.c_void_p(out.data_ptr()),
                                   C.c_void_p(mean_out.data_ptr()) if return_mean else NULL, _stream(out))
    L.check(rc)
    _count()
    return (out, mean_out) if return_mean else out


# ------------------------------------------------------------------------------------------------ F2
def quantile_threshold(u: torch.Tensor, q: float, lerp_fma: bool = False, return_details: bool = False):
    """Per-image threshold == torch.quantile(u.flatten(1).float(), q, dim=1), bit for bit (du_quantile_threshold)."""
    _require_cuda(u, "u")
    if u.dtype != torch.float32:
        u = u.to(torch.float32)  # the reference casts too (posterior_distribution.py:15)
    r = Rows(u, "u")
    thr = torch.empty(r.B, device=u.device, dtype=torch.float32)
    ranks = torch.empty((r.B, 2), device=u.device, dtype=torch.int32) if return_details else None
    vals = torch.empty((r.B, 2), device=u.device, dtype=torch.float32) if return_details else None
    lib = L.load()
    nbytes = lib.du_quantile_scratch_bytes(r.B, r.n)
    scratch = torch.empty(max(int(nbytes), 16), device=u.device, dtype=torch.uint8)
    rc = lib.du_quantile_threshold(r.ptr, r.B, r.n, r.stride, float(q), int(bool(lerp_fma)), C.c_void_p(thr.data_ptr()),
                                   C.c_void_p(ranks.data_ptr()) if return_details else NULL,
                                   C.c_void_p(vals.data_ptr()) if return_details else NULL,
                                   C.c_void_p(scratch.data_ptr()), scratch.numel(), _stream(u))
    if rc == -4 or (rc == -1 and "quantile()" in lib.du_last_error().decode()):
        raise RuntimeError(lib.du_last_error().decode())  # torch.quantile raises RuntimeError for these
    L.check(rc)
    _count()
    return (thr, ranks, vals) if return_details else thr


def threshold_mask(u: torch.Tensor, thr: torch.Tensor, higher: bool = True) -> torch.Tensor:
    r = Rows(u, "u")
    _require_cuda(thr, "thr")
    thr = thr.reshape(-1).to(torch.float32).contiguous()
    if thr.numel() != r.B:
        raise ValueError("threshold_mask: one threshold per image expected")
    mask = torch.empty(u.shape, device=u.device, dtype=torch.float32)
    rc = L.load().du_threshold_mask(r.ptr, r.stride, r.dt, C.c_void_p(thr.data_ptr()), int(higher), r.B, r.n,
                                    C.c_void_p(mask.data_ptr()), r.n, _stream(u))
    L.check(rc)
    _count()
    return mask


def tensor_threshold_mask(u: torch.Tensor, thr_map: torch.Tensor, higher: bool = True) -> torch.Tensor:
    r = Rows(u, "u")
    t = Rows(thr_map, "threshold", batch=False)
    if t.n != r.n:
        raise ValueError(f"tensor threshold has {t.n} elements per image, map has {r.n}")
    mask = torch.empty(u.shape, device=u.device, dtype=torch.float32)
    rc = L.load().du_tensor_threshold_mask(r.ptr, r.stride, r.dt, t.ptr, t.dt, int(higher), r.B, r.n,
                                           C.c_void_p(mask.data_ptr()), r.n, _stream(u))
    L.check(rc)
    _count()
    return mask


def znorm_stats(u: torch.Tensor) -> torch.Tensor:
    """Device tensor [mean, unbiased std, count, M2] over ALL elements of u (du_znorm_stats)."""
    r = Rows(u, "u")
    lib = L.load()
    stats = torch.empty(4, device=u.device, dtype=torch.float32)
    nbytes = int(lib.du_znorm_scratch_bytes(r.B, r.n))
    scratch = torch.empty(max(nbytes // 8, 2), device=u.device, dtype=torch.float64)
    rc = lib.du_znorm_stats(r.ptr, r.stride, r.dt, r.B, r.n, C.c_void_p(stats.data_ptr()), C.c_void_p(scratch.data_ptr()),
                            scratch.numel() * 8, _stream(u))
    L.check(rc)
    _count(2)
    return stats


def znorm_stats_combine(stats_blocks: torch.Tensor) -> torch.Tensor:
    _require_cuda(stats_blocks, "stats")
    s = stats_blocks.reshape(-1, 4).to(torch.float32).contiguous()
    out = torch.empty(4, device=s.device, dtype=torch.float32)
    L.check(L.load().du_znorm_stats_combine(C.c_void_p(s.data_ptr()), s.shape[0], C.c_void_p(out.data_ptr()), _stream(s)))
    _count()
    return out


def znorm_weights(u: torch.Tensor, stats: Optional[torch.Tensor], mode: str = "max", thr: float = 1.0, normalize: bool = True,
                  want_z: bool = True, want_w: bool = True):
    r = Rows(u, "u")
    z = torch.empty(u.shape, device=u.device, dtype=torch.float32) if want_z else None
    w = torch.empty(u.shape, device=u.device, dtype=torch.float32) if want_w else None
    rc = L.load().du_znorm_weights(r.ptr, r.stride, r.dt, C.c_void_p(stats.data_ptr()) if stats is not None else NULL,
                                   int(normalize), _ZN[mode], float(thr), r.B, r.n,
                                   C.c_void_p(z.data_ptr()) if want_z else NULL, r.n,
                                   C.c_void_p(w.data_ptr()) if want_w else NULL, r.n, _stream(u))
    L.check(rc)
    _count()
    return z, w


# ------------------------------------------------------------------------------------------------ F3
def make_coeffs(sqrt_alpha_t, sqrt_beta_t, sqrt_alpha_prev, dir_coef, sigma=0.0, clip_sample=True, clip_range=1.0,
                prediction_type="epsilon", use_clipped_model_output=False, add_noise=False) -> L.DdimCoeffs:
    if prediction_type not in _PRED:
        raise ValueError(f"prediction_type given as {prediction_type} must be one of `epsilon`, `sample`, or `v_prediction`")
    return L.DdimCoeffs(float(sqrt_alpha_t), float(sqrt_beta_t), float(sqrt_alpha_prev), float(dir_coef), float(sigma),
                        float(clip_range), _PRED[prediction_type], int(bool(clip_sample)),
                        int(bool(use_clipped_model_output)), int(bool(add_noise)))


def ddim_step(model_output: torch.Tensor, sample: torch.Tensor, coeffs: L.DdimCoeffs, noise: Optional[torch.Tensor] = None,
              want_prev: bool = True, want_x0: bool = True, want_eps: bool = False):
    """x_{t-1}, x0 (and eps) in one pass (du_ddim_step).  Outputs take sample's dtype promoted with fp32 scalars,
    i.e. sample.dtype, like the reference's eager expressions."""
    mo, s = Rows(model_output, "model_output"), Rows(sample, "sample")
    _same_rows(mo, s, "ddim_step")
    out_dtype = torch.promote_types(model_output.dtype, sample.dtype)
    nz = None
    if coeffs.add_noise:
        if noise is None:
            raise ValueError("ddim_step: eta > 0 needs a noise tensor")
        nz = Rows(noise, "noise")
        _same_rows(mo, nz, "ddim_step(noise)")
    mk = lambda want: torch.empty(sample.shape, device=sample.device, dtype=out_dtype) if want else None  # noqa: E731
    prev, x0, eps = mk(want_prev), mk(want_x0), mk(want_eps)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else NULL  # noqa: E731
    od = _DT[out_dtype]
    rc = L.load().du_ddim_step(mo.ptr, mo.stride, mo.dt, s.ptr, s.stride, s.dt, nz.ptr if nz else NULL, nz.stride if nz else 0,
                               nz.dt if nz else 0, C.byref(coeffs), mo.B, mo.n, p(prev), mo.n, od, p(x0), mo.n, od, p(eps), mo.n, od,
                               _stream(sample))
    L.check(rc)
    _count()
    return prev, x0, eps


def guided_step(eps: torch.Tensor, sample: Optional[torch.Tensor], coeffs: Optional[L.DdimCoeffs], guidance: str = "none",
                u: Optional[torch.Tensor] = None, thr: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                aux: Optional[torch.Tensor] = None, aux_broadcast: bool = False, higher: bool = True, lam: float = 1.0,
                post_M: float = 0.0, inv_alpha_hat: float = 0.0, want_prev: bool = True, want_x0: bool = False,
                want_eps: bool = True, want_mask: bool = False, x0_unguided: bool = False):
    """mask + guided score + DDIM update in one elementwise pass (du_guided_step).
    Returns dict(prev, x0, eps, mask) with None for outputs not requested.  `mask` may be [B,1,H,W] against [B,C,H,W]
    rows (broadcast over channels); x0_unguided: x0 from the unguided eps, direction from the guided one."""
    e = Rows(eps, "eps")
    P = L.GuidedParams()
    P.eps, P.eps_stride, P.eps_dtype, P.guidance = e.ptr, e.stride, e.dt, _GUIDE[guidance]
    keep = [e]
    skip = sample is None
    out_dtype = eps.dtype if skip else torch.promote_types(eps.dtype, sample.dtype)
    if u is not None or mask is not None or aux is not None:
        out_dtype = torch.promote_types(out_dtype, torch.float32)
    if not skip:
        s = Rows(sample, "sample"); _same_rows(e, s, "guided_step(sample)"); keep.append(s)
        P.sample, P.sample_stride, P.sample_dtype = s.ptr, s.stride, s.dt
        P.ddim = coeffs
    P.skip_ddim = int(skip)
    P.higher = int(higher)
    P.x0_unguided = int(bool(x0_unguided))
    if u is not None:
        if u.dtype != torch.float32:
            u = u.float()
        ur = Rows(u, "u"); _same_rows(e, ur, "guided_step(u)"); keep.append(ur)
        P.u, P.u_stride = ur.ptr, ur.stride
    if thr is not None:
        _require_cuda(thr, "thr")
        thr = thr.reshape(-1).to(torch.float32).contiguous(); keep.append(thr)
        if thr.numel() != e.B:
            raise ValueError("guided_step: one threshold per image expected")
        P.thr = C.c_void_p(thr.data_ptr())
    if mask is not None:
        if mask.dtype != torch.float32:
            mask = mask.float()
        mr = Rows(mask, "mask"); keep.append(mr)
        if mr.n != e.n:
            if mr.B != e.B or mr.n == 0 or e.n % mr.n != 0:
                raise ValueError("guided_step: a broadcast mask must hold a whole divisor of the elements of an image")
            P.mask_period = mr.n
        else:
            _same_rows(e, mr, "guided_step(mask)")
        P.mask, P.mask_stride = mr.ptr, mr.stride
    if aux is not None:
        ar = Rows(aux, "aux", batch=not aux_broadcast); keep.append(ar)
        if ar.n != e.n:
            raise ValueError("guided_step: aux has the wrong number of elements per image")
        P.aux, P.aux_stride, P.aux_dtype, P.aux_broadcast = ar.ptr, ar.stride, ar.dt, int(aux_broadcast)
    P.lam, P.post_M, P.inv_alpha_hat = float(lam), float(post_M), float(inv_alpha_hat)
    P.B, P.n = e.B, e.n
    shape, dev = eps.shape, eps.device
    res = {"prev": None, "x0": None, "eps": None, "mask": None}
    od = _DT[out_dtype]
    if want_prev and not skip:
        res["prev"] = torch.empty(shape, device=dev, dtype=out_dtype)
        P.prev_out, P.prev_stride, P.prev_dtype = C.c_void_p(res["prev"].data_ptr()), e.n, od
    if want_x0 and not skip:
        res["x0"] = torch.empty(shape, device=dev, dtype=out_dtype)
        P.x0_out, P.x0_stride, P.x0_dtype = C.c_void_p(res["x0"].data_ptr()), e.n, od
    if want_eps:
        res["eps"] = torch.empty(shape, device=dev, dtype=out_dtype)
        P.eps_out, P.eps_out_stride, P.eps_out_dtype = C.c_void_p(res["eps"].data_ptr()), e.n, od
    if want_mask:
        res["mask"] = torch.empty(shape, device=dev, dtype=torch.float32)
        P.mask_out, P.mask_out_stride = C.c_void_p(res["mask"].data_ptr()), e.n
    L.check(L.load().du_guided_step(C.byref(P), _stream(eps)))
    _count()
    del keep
    return res


def batch_sum(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x.sum(dim=0) as fp32 (du_batch_sum) — the reference's batch-axis sum of the posterior score."""
    r = Rows(x, "x")
    if out is None:
        out = torch.empty(x.shape[1:], device=x.device, dtype=torch.float32)
    elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != r.n or not out.is_cuda:
        raise ValueError("batch_sum: `out` must be a contiguous float32 CUDA tensor of one image's size")
    L.check(L.load().du_batch_sum(r.ptr, r.stride, r.dt, r.B, r.n, C.c_void_p(out.data_ptr()), _stream(x)))
    _count()
    return out


# ------------------------------------------------------------------------------------------------ F7 / F8
def perturb(x: torch.Tensor, noise: Optional[torch.Tensor], a: float, b: float = 0.0,
            out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """a*x + b*noise (du_perturb); noise None: a*x."""
    xr = Rows(x, "x")
    nr = None
    if noise is not None:
        nr = Rows(noise, "noise")
        _same_rows(xr, nr, "perturb")
    out = torch.empty(x.shape, device=x.device,
                      dtype=out_dtype or (x.dtype if noise is None else torch.promote_types(x.dtype, noise.dtype)))
    if out.numel() == 0:
        return out
    rc = L.load().du_perturb(xr.ptr, xr.stride, xr.dt, nr.ptr if nr else NULL, nr.stride if nr else 0, nr.dt if nr else 0,
                             float(a), float(b), xr.B, xr.n, C.c_void_p(out.data_ptr()), xr.n, _DT[out.dtype], _stream(x))
    L.check(rc)
    _count()
    return out


def scale(x: torch.Tensor, a: float) -> torch.Tensor:
    """a*x in one launch (du_perturb without a noise tensor)."""
    return perturb(x, None, a)


def perturb_rows(x: torch.Tensor, noise: torch.Tensor, a_rows: torch.Tensor, b_rows: torch.Tensor) -> torch.Tensor:
    """a[b]*x[b] + b[b]*noise[b] with one scalar pair per row, read from device vectors (du_perturb_rows): add_noise /
    get_velocity with a vector of per-sample timesteps."""
    xr, nr = Rows(x, "x"), Rows(noise, "noise")
    _same_rows(xr, nr, "perturb_rows")
    _require_cuda(a_rows, "a_rows"); _require_cuda(b_rows, "b_rows")
    a_rows = a_rows.reshape(-1).to(torch.float32).contiguous()
    b_rows = b_rows.reshape(-1).to(torch.float32).contiguous()
    if a_rows.numel() != xr.B or b_rows.numel() != xr.B:
        raise ValueError("perturb_rows: one (a, b) pair per row expected")
    out = torch.empty(x.shape, device=x.device, dtype=torch.promote_types(x.dtype, noise.dtype))
    if out.numel() == 0:
        return out
    L.check(L.load().du_perturb_rows(xr.ptr, xr.stride, xr.dt, nr.ptr, nr.stride, nr.dt, C.c_void_p(a_rows.data_ptr()),
                                     C.c_void_p(b_rows.data_ptr()), xr.B, xr.n, C.c_void_p(out.data_ptr()), xr.n, _DT[out.dtype],
                                     _stream(x)))
    _count()
    return out


class _PerturbFn(torch.autograd.Function):
    """a*x + b*noise as a differentiable op (forward du_perturb, backward a*g / b*g by du_perturb's scale form): the re-noising
    of the gradient schedulers, whose model input stays on the autograd graph (SU/scheduling_ddim_uncertainty_grad.py:522-531)."""

    @staticmethod
    def forward(ctx, x, noise, a, b):
        ctx.a, ctx.b = float(a), float(b)
        return perturb(x.detach(), noise.detach(), a, b)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        gx = scale(g, ctx.a) if ctx.needs_input_grad[0] else None
        gn = scale(g, ctx.b) if ctx.needs_input_grad[1] else None
        return gx, gn, None, None


def perturb_autograd(x: torch.Tensor, noise: torch.Tensor, a: float, b: float) -> torch.Tensor:
    return _PerturbFn.apply(x, noise, a, b)


class _PerturbFreshFn(torch.autograd.Function):
    """a*x + b*randn_like(noise_like), differentiable in x, the draw made inside the kernel when it can be (perturb_fresh)."""

    @staticmethod
    def forward(ctx, x, a, b, noise_like):
        ctx.a = float(a)
        return perturb_fresh(x.detach(), a, b, noise_like=noise_like)

    @staticmethod
    def backward(ctx, g):
        return (scale(g.contiguous(), ctx.a) if ctx.needs_input_grad[0] else None), None, None, None


def perturb_fresh_autograd(x: torch.Tensor, a: float, b: float, noise_like: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _PerturbFreshFn.apply(x, a, b, noise_like)


def ema_update(momentum: Optional[torch.Tensor], u: torch.Tensor, beta: float, step: int):
    """(momentum', momentum' / (1 - beta**step + 1e-5), sqrt of that) in one launch (du_ema_update) —
    PU/pipeline_sampler_class_conditional_uncertainty_guided_second_order.py:212-218.  momentum None: first step."""
    _require_cuda(u, "u")
    u = u if u.is_contiguous() else u.contiguous()
    if momentum is not None:
        _require_cuda(momentum, "momentum")
        momentum = momentum.to(torch.float32).contiguous()
        if momentum.numel() != u.numel():
            raise ValueError("ema_update: momentum and map differ in size")
    new = torch.empty(u.shape, device=u.device, dtype=torch.float32)
    corrected, root = torch.empty_like(new), torch.empty_like(new)
    if u.numel() == 0:
        return new, corrected, root
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else NULL  # noqa: E731
    L.check(L.load().du_ema_update(p(momentum), p(u), _DT[u.dtype], float(beta), float(1 - beta), float(1 - beta ** step + 1e-5),
                                   u.numel(), p(new), p(corrected), p(root), _stream(u)))
    _count()
    return new, corrected, root


def dpm_solver_update(sample: torch.Tensor, m0: torch.Tensor, m1: Optional[torch.Tensor], a: float, b: float, c: float = 0.0,
                      k: float = 0.0) -> torch.Tensor:
    """(a*sample + b*m0) + c*(k*(m0 - m1)) in one launch (du_dpm_solver_update); m1 None: a*sample + b*m0."""
    sr, r0 = Rows(sample, "sample"), Rows(m0, "m0")
    _same_rows(sr, r0, "dpm_solver_update")
    r1 = None
    if m1 is not None:
        r1 = Rows(m1, "m1")
        _same_rows(sr, r1, "dpm_solver_update")
    dt = torch.promote_types(sample.dtype, m0.dtype)
    out = torch.empty(sample.shape, device=sample.device, dtype=dt)
    rc = L.load().du_dpm_solver_update(sr.ptr, sr.stride, sr.dt, r0.ptr, r0.stride, r0.dt, r1.ptr if r1 else None,
                                       r1.stride if r1 else 0, r1.dt if r1 else L.F32, float(a), float(b), float(c), float(k),
                                       sr.B, sr.n, C.c_void_p(out.data_ptr()), sr.n, _DT[dt], _stream(sample))
    L.check(rc)
    _count()
    return out


_TORCH_RANDN_LIKE = torch.randn_like     # the genuine function: a caller that replaced torch.randn_like keeps its replacement


def randn_fusable(x: torch.Tensor, generator: Optional[torch.Generator] = None) -> bool:
    """True when `torch.randn_like(x)` may be drawn inside the perturbation kernel with the identical result: a dense CUDA
    tensor (torch's generator kernel then indexes elements linearly, the mapping du_perturb_randn reproduces), fewer than
    2^31 elements, torch.randn_like not replaced by the caller, not inside a CUDA-graph capture (torch's generator keeps its
    offset on the device there; use DeviceRng for captured draws)."""
    return (x.is_cuda and x.dtype in _DT and x.is_contiguous() and 0 < x.numel() < 2 ** 31
            and torch.randn_like is _TORCH_RANDN_LIKE and not torch.cuda.is_current_stream_capturing()
            and (generator is None or generator.device.type == "cuda"))


def randn_offset_increment(numel: int) -> int:
    """Philox offset one `numel`-element normal draw consumes on the current device (du_randn_offset_increment)."""
    inc = C.c_uint64(0)
    L.check(L.load().du_randn_offset_increment(int(numel), C.byref(inc)))
    return int(inc.value)


def _claim_philox(x: torch.Tensor, generator: Optional[torch.Generator]):
    """Take (seed, offset) for one draw of x.numel() normals from the torch generator of x's device and advance it exactly
    as torch's own kernel launch would (CUDAGeneratorImpl::philox_cuda_state)."""
    if generator is not None:
        gen = generator
    else:
        torch.cuda.init()          # default_generators is empty until CUDA is initialised
        gen = torch.cuda.default_generators[x.device.index if x.device.index is not None else torch.cuda.current_device()]
    seed, offset = gen.initial_seed(), gen.get_offset()
    gen.set_offset(offset + randn_offset_increment(x.numel()))
    return seed & 0xFFFFFFFFFFFFFFFF, offset


def perturb_randn(x: torch.Tensor, a: float, b: float, generator: Optional[torch.Generator] = None, want_noise: bool = False,
                  device_state: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None, extra_offset: int = 0):
    """a*x + b*n, n = what `torch.randn_like(x)` would return now (bit-identical, generator advanced the same way), drawn in
    the kernel (du_perturb_randn): x read once, out written once, no noise tensor in HBM unless want_noise.
    device_state: int64[2] CUDA tensor {seed, offset} for draws replayed from a CUDA graph (see DeviceRng)."""
    if not (x.is_cuda and x.dtype in _DT and x.is_contiguous()):
        raise RuntimeError("perturb_randn needs a dense CUDA tensor (no CPU fallback; use torch.randn_like + perturb for views)")
    _stream_ptr = _stream(x)
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    noise = torch.empty_like(x) if want_noise else None
    if device_state is not None:
        seed, offset, st_ptr = 0, int(extra_offset), C.c_void_p(device_state.data_ptr())
    else:
        (seed, offset), st_ptr = _claim_philox(x, generator), None
    rc = L.load().du_perturb_randn(C.c_void_p(x.data_ptr()), _DT[x.dtype], x.numel(), seed, offset, st_ptr, float(a), float(b),
                                   C.c_void_p(out.data_ptr()), _DT[out.dtype],
                                   C.c_void_p(noise.data_ptr()) if want_noise else None, _DT[x.dtype], _stream_ptr)
    L.check(rc)
    _count()
    return (out, noise) if want_noise else out


def randn_like(x: torch.Tensor, generator: Optional[torch.Generator] = None, device_state: Optional[torch.Tensor] = None,
               extra_offset: int = 0):
    """`torch.randn_like(x)` from du_perturb_randn's generator path alone (used by the tests and by DeviceRng)."""
    if not (x.is_cuda and x.dtype in _DT):
        raise RuntimeError("randn_like needs a CUDA tensor")
    out = torch.empty(x.shape, device=x.device, dtype=x.dtype)
    if x.numel() == 0:
        return out
    _stream_ptr = _stream(x)
    if device_state is not None:
        seed, offset, st_ptr = 0, int(extra_offset), C.c_void_p(device_state.data_ptr())
    else:
        (seed, offset), st_ptr = _claim_philox(out, generator), None
    L.check(L.load().du_perturb_randn(None, _DT[x.dtype], out.numel(), seed, offset, st_ptr, 0.0, 0.0, None, _DT[x.dtype],
                                      C.c_void_p(out.data_ptr()), _DT[x.dtype], _stream_ptr))
    _count()
    return out


def perturb_fresh(x: torch.Tensor, a: float, b: float, noise_like: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`a*x + b*torch.randn_like(noise_like or x)`: ONE launch with the noise drawn in registers when torch's draw can be
    reproduced in the kernel (randn_fusable), else torch.randn_like followed by du_perturb.  Same values, same generator
    state afterwards, either way."""
    like = x if noise_like is None else noise_like
    same = like.shape == x.shape and like.dtype == x.dtype and like.device == x.device
    if same and capture_rng is not None and x.is_cuda and x.is_contiguous() and torch.cuda.is_current_stream_capturing():
        return capture_rng.perturb(x, a, b)        # inside a graph capture: the device-resident Philox state
    if same and randn_fusable(x):
        return perturb_randn(x, a, b)
    noise = torch.randn_like(like)
    return perturb(x, noise if noise.shape == x.shape else noise.expand(x.shape), a, b)


class DeviceRng:
    """A Philox {seed, offset} pair in device memory, for perturbation draws inside a captured CUDA graph: every replay
    continues the stream (du_rng_advance is part of the graph), and the values equal torch's for the same (seed, offset).

    deferred=True: the draws of one graph pass their running increment as the extra offset and ONE `commit()` at the end of the
    graph moves the state past all of them (a window step then costs one tiny launch for the RNG instead of one per draw)."""

    def __init__(self, device, seed: int, offset: int = 0, deferred: bool = False):
        self.state = torch.tensor([seed, offset], dtype=torch.int64, device=device)
        self.deferred = deferred
        self._pending = 0

    def _draw(self, fn, x, *a, **kw):
        r = fn(x, *a, device_state=self.state, extra_offset=self._pending, **kw)
        inc = randn_offset_increment(x.numel())
        if self.deferred:
            self._pending += inc
        else:
            self.advance(x.numel(), x)
        return r

    def perturb(self, x: torch.Tensor, a: float, b: float, want_noise: bool = False):
        return self._draw(perturb_randn, x, a, b, want_noise=want_noise)

    def randn_like(self, x: torch.Tensor):
        return self._draw(randn_like, x)

    def commit(self, like: torch.Tensor):
        """deferred mode: advance the device state past every draw since the last commit (one launch)."""
        if self._pending:
            L.check(L.load().du_rng_advance(C.c_void_p(self.state.data_ptr()), self._pending, _stream(like)))
            _count()
            self._pending = 0

    def advance(self, numel: int, like: torch.Tensor):
        L.check(L.load().du_rng_advance(C.c_void_p(self.state.data_ptr()), randn_offset_increment(numel), _stream(like)))
        _count()

    def offset(self) -> int:
        return int(self.state[1].item())


# A DeviceRng that perturb_fresh() uses for its draws while a CUDA graph is being captured (graphed.GraphedWindowStep sets it):
# torch's own generator cannot be advanced by the host inside a capture, the device-resident state can.
capture_rng: Optional[DeviceRng] = None


def accumulate_slot(src: torch.Tensor, dst_slot: torch.Tensor):
    """Copy/convert one step's map into its slot view buffer[:, t] (du_accumulate_slot)."""
    s, d = Rows(src, "src"), Rows(dst_slot, "dst")
    if d.t is not dst_slot:
        raise ValueError("accumulate_slot: destination rows must be contiguous")
    _same_rows(s, d, "accumulate_slot")
    L.check(L.load().du_accumulate_slot(s.ptr, s.stride, s.dt, s.B, s.n, d.ptr, d.stride, d.dt, _stream(src)))
    _count()


def image_uint8(x: torch.Tensor) -> torch.Tensor:
    """uint8 images from samples in [-1, 1]: (x/2 + 0.5).clamp(0,1) * 255, rounded half to even (du_image_uint8)."""
    r = Rows(x, "x")
    out = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    L.check(L.load().du_image_uint8(r.ptr, r.stride, r.dt, r.B, r.n, C.c_void_p(out.data_ptr()), r.n, _stream(x)))
    _count()
    return out


# ------------------------------------------------------------------------------------------------ whole step
def fused_supported(n: int, dtype: torch.dtype) -> int:
    """Cluster size the fused kernel would use for rows of n elements (0 = not supported)."""
    return int(L.load().du_fused_supported(int(n), _DT[dtype]))


def fused_last_kernel() -> str:
    """Name of the kernel the last fused launch of this thread used ('fused_step_kernel': three-phase cluster kernel,
    'fused_pred_kernel': predictive single-pass kernel; '' before the first launch)."""
    return {0: "", 1: "fused_step_kernel", 2: "fused_pred_kernel"}[int(L.load().du_fused_last_kernel())]


class FusedStep:
    """A prepared du_fused_uncertainty_step call: the parameter block is built once, `launch()` is a single
    C-ABI call (one kernel launch), so a sampling loop (or a CUDA-graph capture) pays no per-step Python
    marshalling.  Outputs live in `self.res` (dict u, thr, prev, x0, eps, mask) and are overwritten by every
    launch; `set_map_out()` re-targets the map at another slot of the accumulation buffer (F8)."""

    def __init__(self, scores: Sequence[torch.Tensor], eps: torch.Tensor, sample: Optional[torch.Tensor], q: float,
                 coeffs: Optional[L.DdimCoeffs], alpha_hat_t: float, moments_mode: str = "var_with_center",
                 S: Optional[torch.Tensor] = None, S_broadcast: bool = False, higher: bool = True,
                 map_out: Optional[torch.Tensor] = None, lerp_fma: bool = False, want_x0: bool = False,
                 want_eps: bool = False, want_mask: bool = False, post_M: Optional[float] = None,
                 prev_out: Optional[torch.Tensor] = None, S_overlap: bool = False):
        M = len(scores)
        if M < 1 or M > L.DU_MAX_M:
            raise ValueError(f"fused step: M={M} must be in [1, {L.DU_MAX_M}]")
        rows = [Rows(t, f"scores[{k}]") for k, t in enumerate(scores)]
        r0 = rows[0]
        skip = sample is None       # skip_ddim: the step ends with the guided score (res["eps"]); no sample, no x_{t-1}
        e = Rows(eps, "eps")
        sm = None if skip else Rows(sample, "sample")
        for r in rows[1:] + [e]:
            _same_rows(r0, r, "fused step")
            if r.dt != r0.dt or r.stride != (r0.stride if r is not e else r.stride):
                raise RuntimeError("fused step: scores and eps must share dtype (and scores one row stride)")
        P = L.FusedParams()
        for k, r in enumerate(rows):
            P.scores[k] = r.ptr.value
        P.M, P.score_dtype, P.score_stride = M, r0.dt, r0.stride
        P.eps, P.eps_stride = e.ptr, e.stride
        if not skip:
            _same_rows(r0, sm, "fused step(sample)")
            P.sample, P.sample_stride, P.sample_dtype = sm.ptr, sm.stride, sm.dt
        P.skip_ddim = int(skip)
        P.moments_mode = _MODES[moments_mode]
        self._keep = [rows, e, sm]
        if S is not None:
            if S.dtype != torch.float32:
                S = S.float()
            sr = Rows(S, "S", batch=not S_broadcast)
            self._keep.append(sr)
            if sr.n != r0.n:
                raise ValueError("fused step: S has the wrong number of elements per image")
            P.S, P.S_stride, P.S_broadcast = sr.ptr, sr.stride, int(S_broadcast)
            # S_overlap: the caller guarantees that the du_batch_sum writing S is the launch right before this one on the
            # stream; the step then starts as its programmatic dependent (see launch_with_batch_sum)
            P.S_overlap = int(bool(S_overlap))
        P.higher, P.q, P.lerp_fma = int(higher), float(q), int(bool(lerp_fma))
        P.post_M = float(M if post_M is None else post_M)
        P.inv_alpha_hat = float(1.0 / alpha_hat_t)
        if coeffs is not None:
            P.ddim = coeffs
        elif not skip:
            raise ValueError("fused step: DDIM coefficients are required unless sample is None (skip_ddim)")
        P.B, P.n = r0.B, r0.n
        dev, shape = eps.device, eps.shape
        out_dtype = torch.promote_types(torch.promote_types(eps.dtype, eps.dtype if skip else sample.dtype), torch.float32)
        self.P, self._r0, self._dev_tensor = P, r0, eps
        thr = torch.empty(r0.B, device=dev, dtype=torch.float32)
        P.thr_out = C.c_void_p(thr.data_ptr())
        res = {"u": None, "thr": thr, "x0": None, "eps": None, "mask": None, "prev": None}
        if skip:
            want_x0, want_eps = False, True
        else:
            self.set_prev_out(prev_out if prev_out is not None else torch.empty(shape, device=dev, dtype=out_dtype), res)
        if want_x0:
            res["x0"] = torch.empty(shape, device=dev, dtype=out_dtype)
            P.x0_out, P.x0_stride = C.c_void_p(res["x0"].data_ptr()), r0.n
        if want_eps:
            res["eps"] = torch.empty(shape, device=dev, dtype=torch.float32)
            P.eps_out, P.eps_out_stride = C.c_void_p(res["eps"].data_ptr()), r0.n
        if want_mask:
            res["mask"] = torch.empty(shape, device=dev, dtype=torch.float32)
            P.mask_out, P.mask_out_stride = C.c_void_p(res["mask"].data_ptr()), r0.n
        self.res = res
        self.set_map_out(map_out if map_out is not None else torch.empty(shape, device=dev, dtype=torch.float32))
        self._fn = L.load().du_fused_uncertainty_step
        self._ref = C.byref(P)

    def set_prev_out(self, prev_out: torch.Tensor, res=None):
        """Re-target x_{t-1} (a sampling loop hands the step the tensor its next model call reads)."""
        pr = Rows(prev_out, "prev_out")
        if pr.t is not prev_out or prev_out.dtype != torch.float32:
            raise ValueError("fused step: prev_out must be a float32 tensor with contiguous rows")
        _same_rows(self._r0, pr, "fused step(prev_out)")
        (self.res if res is None else res)["prev"] = prev_out
        self.P.prev_out, self.P.prev_stride, self.P.prev_dtype = pr.ptr, pr.stride, L.F32

    def set_map_out(self, u: torch.Tensor):
        ur = Rows(u, "map_out")
        if ur.t is not u or u.dtype != torch.float32:
            raise ValueError("fused step: map_out must be a float32 tensor with contiguous rows")
        _same_rows(self._r0, ur, "fused step(map_out)")
        self.P.unc_out, self.P.unc_stride = ur.ptr, ur.stride
        self.res["u"] = u

    def set_S(self, S: torch.Tensor, broadcast: bool):
        sr = Rows(S, "S", batch=not broadcast)
        self._keep.append(sr)
        self.P.S, self.P.S_stride, self.P.S_broadcast = sr.ptr, sr.stride, int(broadcast)

    def launch(self):
        L.check(self._fn(self._ref, _stream(self._dev_tensor)))
        _count()
        return self.res

    def launch_with_batch_sum(self, eps: torch.Tensor, S_out: torch.Tensor):
        """The reference's `pred_epsilon.sum(dim=0)` (uncertainty_guidance.py:116-119) followed by the step that consumes it, as
        two back-to-back launches on one stream: du_batch_sum -> du_fused_uncertainty_step with S_overlap set, so the step
        starts as the sum's programmatic dependent and its S-independent pilot overlaps the sum.  S_out: fp32 [C,H,W] row."""
        batch_sum(eps, out=S_out)
        if self.P.S != S_out.data_ptr() or not self.P.S_broadcast:
            self.set_S(S_out, True)
        self.P.S_overlap = 1
        try:
            return self.launch()
        finally:
            self.P.S_overlap = 0


def fused_uncertainty_step(scores: Sequence[torch.Tensor], eps: torch.Tensor, sample: torch.Tensor, q: float,
                           coeffs: L.DdimCoeffs, alpha_hat_t: float, **kw):
    """ONE launch: moments -> per-image quantile -> mask -> posterior blend -> DDIM (du_fused_uncertainty_step).
    Keyword arguments as FusedStep."""
    return FusedStep(scores, eps, sample, q, coeffs, alpha_hat_t, **kw).launch()


def _fused_eligible(scores, eps, sample, map_out, S) -> bool:
    n = eps[0].numel() if eps.shape[0] > 0 else 0
    if n == 0 or eps.dtype not in _DT or not fused_supported(n, eps.dtype):
        return False
    vec = 4 if eps.dtype == torch.float32 else 8
    ts = list(scores) + [eps] + [t for t in (sample, map_out, S) if t is not None]
    for t in ts:
        if t.data_ptr() % 16 or (t.dim() > 1 and t.shape[0] > 1 and t.stride(0) % vec) or not t[0].is_contiguous():
            return False
    return all(s.dtype == eps.dtype for s in scores) and eps.shape[0] <= 65535


def uncertainty_step(scores: Sequence[torch.Tensor], eps: torch.Tensor, sample: Optional[torch.Tensor], q: float,
                     coeffs: Optional[L.DdimCoeffs], alpha_hat_t: float, moments_mode: str = "var_with_center",
                     sum_source: Optional[torch.Tensor] = None, batch_sum: bool = True, higher: bool = True,
                     map_out: Optional[torch.Tensor] = None, lerp_fma: bool = False, want_x0: bool = False,
                     want_eps: bool = False, want_mask: bool = False, fused: Optional[bool] = None,
                     precomputed_sum: Optional[torch.Tensor] = None, prev_out: Optional[torch.Tensor] = None):
    """The percentile-guided posterior step, F1 -> F2a -> F5 -> F3 (+F8 when `map_out` is a slot of the
    accumulation buffer): PU/...posterior_distribution.py:153-162 + uncertainty_guidance.py:101-120 +
    SU/...zigzag_centered.py:472-510.  Returns dict(u, thr, prev, x0, eps, mask).
    batch_sum=True reproduces the reference's `sum(dim=0)` over the batch axis (identity at B=1); `precomputed_sum` is that
    row when the caller already has it (image chunks of a larger batch, or the all-reduced row under batch sharding).
    fused=None picks the single-launch cluster kernel whenever the shape/alignment allows it.
    sample None: the step ends with the guided score (`eps` in the result; du_fused_params::skip_ddim) — the body of
    get_uncertainty_guided_score_with_percentile, whose caller applies its own scheduler."""
    global last_step_path
    M = len(scores)
    skip = sample is None
    for t in list(scores) + [eps] + ([] if skip else [sample]):
        _require_cuda(t, "uncertainty_step input")
    src = eps if sum_source is None else sum_source
    S, bcast, summed_here = (None if sum_source is None else src), False, False
    if batch_sum and precomputed_sum is not None:
        S, bcast = precomputed_sum, True
    elif batch_sum and eps.shape[0] > 1:
        S, bcast, summed_here = batch_sum_fn(src), True, True
    if fused is None:
        fused = _fused_eligible(scores, eps, sample, map_out, S)
    last_step_path = "fused" if fused else "unfused"
    if fused:
        # (the fused launch directly follows the du_batch_sum above on this stream: it may start as its dependent launch)
        return fused_uncertainty_step(scores, eps, sample, q, coeffs, alpha_hat_t, moments_mode=moments_mode, S=S,
                                      S_broadcast=bcast, higher=higher, map_out=map_out, lerp_fma=lerp_fma, want_x0=want_x0,
                                      want_eps=want_eps, want_mask=want_mask, prev_out=prev_out, S_overlap=summed_here)
    u = moments(scores, center=eps, mode=moments_mode, out=map_out)
    thr = quantile_threshold(u, q, lerp_fma=lerp_fma)
    r = guided_step(eps, sample, coeffs, guidance="posterior", u=u, thr=thr, aux=src if S is None else S, aux_broadcast=bcast,
                    higher=higher, post_M=float(M), inv_alpha_hat=float(1.0 / alpha_hat_t), want_prev=not skip,
                    want_x0=want_x0 and not skip, want_eps=want_eps or skip, want_mask=want_mask)
    r["u"], r["thr"] = u, thr
    if prev_out is not None and not skip:
        accumulate_slot(r["prev"], prev_out)
        r["prev"] = prev_out
    return r


batch_sum_fn = batch_sum
