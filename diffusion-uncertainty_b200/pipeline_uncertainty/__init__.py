"""Drop-in functions and classes of diffusion_uncertainty.pipeline_uncertainty that sit on the uncertainty path."""
from .threshold_guidance import calculate_threshold_map, estimate_score_update_posterior  # noqa: F401
from .pipeline_sampler_class_conditional_uncertainty import DiffusionClassConditionalWithUncertainty  # noqa: F401,E402
from .pipeline_sampler_class_conditional_uncertainty_guided_posterior_distribution import \
    DiffusionClassConditionalGuidedPosteriorDistribution  # noqa: F401,E402
from .pipeline_sampler_class_conditional_uncertainty_guided_second_order import DiffusionClassConditionalGuidedSecondOrder  # noqa: F401,E402
