"""Drop-in for diffusion_uncertainty/schedulers_uncertainty/scheduling_ddim_mc_dropout.py (factory key: mc_dropout (factory default)).
Same four class names, constructor arguments and step() signature; the arithmetic runs in libdu_b200.so
(see _variants.MCDropout for the reference block it reproduces)."""
from ..outputs import DDIMSchedulerUncertaintyOutput  # noqa: F401
from ._families import make_family
from ._variants import MCDropout

globals().update(make_family(MCDropout, __name__))
__all__ = ["DDIMSchedulerUncertaintyOutput", "DDIMSchedulerUncertainty", "DDIMSchedulerUncertaintyImagenet",
           "DDIMSchedulerUncertaintyCifar10", "DDIMSchedulerUncertaintyImagenetClassConditioned"]
