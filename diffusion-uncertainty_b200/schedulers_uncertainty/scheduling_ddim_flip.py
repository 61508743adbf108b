"""Drop-in for diffusion_uncertainty/schedulers_uncertainty/scheduling_ddim_flip.py (factory key: flip).
Same four class names, constructor arguments and step() signature; the arithmetic runs in libdu_b200.so
(see _variants.Flip for the reference block it reproduces)."""
from ..outputs import DDIMSchedulerUncertaintyOutput  # noqa: F401
from ._families import make_family
from ._variants import Flip

globals().update(make_family(Flip, __name__))
__all__ = ["DDIMSchedulerUncertaintyOutput", "DDIMSchedulerUncertainty", "DDIMSchedulerUncertaintyImagenet",
           "DDIMSchedulerUncertaintyCifar10", "DDIMSchedulerUncertaintyImagenetClassConditioned"]
