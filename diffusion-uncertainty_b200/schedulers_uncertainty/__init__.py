"""Drop-in replacements for diffusion_uncertainty.schedulers_uncertainty (same module and class names)."""
from .get_uncertainty_scheduler import get_uncertainty_scheduler  # noqa: F401
from .mixin import SchedulerUncertaintyClassConditionedMixin, SchedulerUncertaintyMixin  # noqa: F401
from .traits import PredictorClassConditionedTrait  # noqa: F401
