"""Drop-in for diffusion_uncertainty/schedulers_uncertainty/scheduling_ddim_flip_threshold.py (factory key: fid: flip_threshold).
Same four class names, constructor arguments and step() signature; the arithmetic runs in libdu_b200.so
(see _variants.FlipThreshold for the reference block it reproduces)."""
from ..outputs import DDIMSchedulerUncertaintyOutput  # noqa: F401
from ._families import make_family
from ._variants import FlipThreshold

globals().update(make_family(FlipThreshold, __name__))
__all__ = ["DDIMSchedulerUncertaintyOutput", "DDIMSchedulerUncertainty", "DDIMSchedulerUncertaintyImagenet",
           "DDIMSchedulerUncertaintyCifar10", "DDIMSchedulerUncertaintyImagenetClassConditioned"]
