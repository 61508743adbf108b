"""Module-name alias of the reference's pipeline_uncertainty/pipeline_sampler_class_conditional_uncertainty_guided_posterior_distribution.py for the functions on the uncertainty path."""
from .threshold_guidance import calculate_threshold_map, estimate_score_update_posterior  # noqa: F401
