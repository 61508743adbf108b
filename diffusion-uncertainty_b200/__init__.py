"""diffusion-uncertainty_b200 — B200-native per-step uncertainty path of Michedev/diffusion-uncertainty.

Drop-in for ONE hot path of the reference: the moments over the M perturbed / dropout score
predictions, the per-image percentile threshold and mask, the guided DDIM/DDPM x_{t-1} update and the
uncertainty-map accumulation, behind the reference's own scheduler / pipeline API.  The arithmetic runs
in hand-written sm_100a CUDA kernels (csrc/, C ABI in include/du_b200.h); PyTorch only owns memory and
streams.  The score models stay the reference's PyTorch modules.

Import name: `diffusion_uncertainty_b200` (the directory keeps the repository's hyphenated name; the
root-level shim diffusion_uncertainty_b200.py maps one to the other).
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"


def library_path() -> str:
    return _lib.LIB_PATH
