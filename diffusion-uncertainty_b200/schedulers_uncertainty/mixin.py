"""Structural types the sampling loops test with `isinstance`.

The reference declares two runtime-checkable protocols (diffusion_uncertainty/schedulers_uncertainty/mixin.py:4-15) and its loops
ask `isinstance(scheduler, ...)` to decide whether a step returns an uncertainty map (`generate_samples.py:153,172,189`): a scheduler
"is" an uncertainty scheduler when it carries the window attributes, and a class-conditioned one when it also says so.  Same names,
same attribute sets; here the class-conditioned protocol extends the plain one instead of repeating it."""
import typing


@typing.runtime_checkable
class SchedulerUncertaintyMixin(typing.Protocol):
    """Has an uncertainty window: its first and last timestep, set by set_timesteps() (zigzag_centered.py:381-384)."""
    timestep_after_step: int
    timestep_end_step: int


@typing.runtime_checkable
class SchedulerUncertaintyClassConditionedMixin(SchedulerUncertaintyMixin, typing.Protocol):
    """Has an uncertainty window and calls its score model with class labels."""
    class_conditioned: bool
