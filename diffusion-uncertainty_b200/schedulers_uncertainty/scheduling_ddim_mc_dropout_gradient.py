"""Drop-in for diffusion_uncertainty/schedulers_uncertainty/scheduling_ddim_mc_dropout_gradient.py (factory key: fid: mc_dropout_gradient).
Same four class names, constructor arguments and step() signature; the arithmetic runs in libdu_b200.so
(see _variants.MCDropoutGradient for the reference block it reproduces)."""
from ..outputs import DDIMSchedulerUncertaintyOutput  # noqa: F401
from ._families import make_family
from ._variants import MCDropoutGradient

globals().update(make_family(MCDropoutGradient, __name__))
__all__ = ["DDIMSchedulerUncertaintyOutput", "DDIMSchedulerUncertainty", "DDIMSchedulerUncertaintyImagenet",
           "DDIMSchedulerUncertaintyCifar10", "DDIMSchedulerUncertaintyImagenetClassConditioned"]
