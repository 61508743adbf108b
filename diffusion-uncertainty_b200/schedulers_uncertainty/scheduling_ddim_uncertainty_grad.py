"""Drop-in for diffusion_uncertainty/schedulers_uncertainty/scheduling_ddim_uncertainty_grad.py (factory key: fid: uncertainty_grad).
Same four class names, constructor arguments and step() signature; the arithmetic runs in libdu_b200.so
(see _variants.UncertaintyGrad for the reference block it reproduces)."""
from ..outputs import DDIMSchedulerUncertaintyOutput  # noqa: F401
from ._families import make_family
from ._variants import UncertaintyGrad

globals().update(make_family(UncertaintyGrad, __name__))
__all__ = ["DDIMSchedulerUncertaintyOutput", "DDIMSchedulerUncertainty", "DDIMSchedulerUncertaintyImagenet",
           "DDIMSchedulerUncertaintyCifar10", "DDIMSchedulerUncertaintyImagenetClassConditioned"]
