"""Module-name alias of the reference's pipeline_uncertainty/uncertainty_guidance.py:
`generate_samples_model_scheduler_class_conditioned_with_threshold` (:12-125) — implemented in guided_loops.py."""
from ..guided_loops import generate_samples_model_scheduler_class_conditioned_with_threshold  # noqa: F401
