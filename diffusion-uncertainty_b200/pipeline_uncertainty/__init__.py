"""Drop-in functions and classes of diffusion_uncertainty.pipeline_uncertainty that sit on the uncertainty path."""
from .threshold_guidance import calculate_threshold_map, estimate_score_update_posterior  # noqa: F401
