"""Structural types the sampling loops test with `isinstance`.
Mirrors diffusion_uncertainty/schedulers_uncertainty/mixin.py:4-15: a scheduler "is" an uncertainty scheduler when it
carries the window attributes (`generate_samples.py:153,172,189`)."""
from typing import Protocol, runtime_checkable


@runtime_checkable
class SchedulerUncertaintyMixin(Protocol):
    timestep_after_step: int
    timestep_end_step: int


@runtime_checkable
class SchedulerUncertaintyClassConditionedMixin(Protocol):
    class_conditioned: bool
    timestep_after_step: int
    timestep_end_step: int
