"""Drop-in for diffusion_uncertainty/schedulers_uncertainty/scheduling_ddim_uncertainty_image.py (factory key: uncertainty_image).
Same four class names, constructor arguments and step() signature; the arithmetic runs in libdu_b200.so
(see _variants.UncertaintyImage for the reference block it reproduces)."""
from ..outputs import DDIMSchedulerUncertaintyOutput  # noqa: F401
from ._families import make_family
from ._variants import UncertaintyImage

globals().update(make_family(UncertaintyImage, __name__))
__all__ = ["DDIMSchedulerUncertaintyOutput", "DDIMSchedulerUncertainty", "DDIMSchedulerUncertaintyImagenet",
           "DDIMSchedulerUncertaintyCifar10", "DDIMSchedulerUncertaintyImagenetClassConditioned"]
