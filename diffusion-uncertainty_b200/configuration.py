"""Scheduler configuration plumbing of the drop-in classes.

The reference schedulers inherit `ConfigMixin` / `register_to_config` from diffusers 0.31 (imports at
diffusion_uncertainty/schedulers_uncertainty/scheduling_ddim_uncertainty_zigzag_centered.py:26-34).  The hot path
needs three behaviours of that machinery and nothing else:

  * `scheduler.config.<name>` is readable AND assignable — reference callers mutate `config.after_step` /
    `config.num_steps_uc` between runs (pipeline_uncertainty/pipeline_sampler_class_conditional_uncertainty_guided_gradient.py:64-66)
    and `set_timesteps` re-reads them (…zigzag_centered.py:381-384);
  * `Cls.from_config(config, **overrides)` builds a scheduler from another scheduler's config (a dict-like or an
    object with attribute access), silently dropping keys the constructor does not accept
    (schedulers_uncertainty/get_uncertainty_scheduler.py:13-34 passes `y=`, `eta=`, … that most classes ignore);
  * every constructor argument is recorded under its name.

This module provides exactly that, with no third-party base class.
"""
from __future__ import annotations

import functools
import inspect
from typing import Any, Dict


class SchedulerConfig(dict):
    """Constructor arguments of a scheduler: a dict with attribute read / write access."""

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(f"scheduler config has no entry {name!r}") from e

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def __delattr__(self, name: str) -> None:
        del self[name]


def _ctor_parameters(cls) -> Dict[str, inspect.Parameter]:
    """Named parameters accepted by the constructor chain of `cls` (a `**variant` catch-all in a subclass forwards to
    its base, so walk the MRO)."""
    params: Dict[str, inspect.Parameter] = {}
    for klass in cls.__mro__:
        init = klass.__dict__.get("__init__")
        if init is None:
            continue
        init = getattr(init, "__wrapped__", init)
        for name, prm in inspect.signature(init).parameters.items():
            if name == "self" or prm.kind in (prm.VAR_POSITIONAL, prm.VAR_KEYWORD):
                continue
            params.setdefault(name, prm)
    return params


def records_config(init):
    """Decorator for `__init__`: store every named constructor argument (defaults included) in `self.config`."""

    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = getattr(self, "_config", None)
        if cfg is None:
            cfg = SchedulerConfig()
            object.__setattr__(self, "_config", cfg)
        for name, value in bound.arguments.items():
            prm = sig.parameters[name]
            if name == "self":
                continue
            if prm.kind == prm.VAR_KEYWORD:
                for k, v in value.items():
                    cfg.setdefault(k, v)
            elif prm.kind != prm.VAR_POSITIONAL:
                cfg[name] = value
        init(self, *args, **kwargs)

    return wrapper


class ConfigurableScheduler:
    """`config` property + `from_config` classmethod (the two ConfigMixin behaviours the callers use)."""

    config_name = "scheduler_config.json"

    @property
    def config(self) -> SchedulerConfig:
        return self._config

    @classmethod
    def from_config(cls, config=None, **kwargs):
        """Build from another scheduler's config; unknown keys (in `config` or `kwargs`) are dropped, `kwargs` win."""
        if config is None:
            config = {}
        if not isinstance(config, dict):  # FrozenDict-like objects, namespaces
            config = dict(config.items()) if hasattr(config, "items") else dict(vars(config))
        accepted = _ctor_parameters(cls)
        init = {k: v for k, v in config.items() if k in accepted and not k.startswith("_")}
        init.update({k: v for k, v in kwargs.items() if k in accepted})
        return cls(**init)
