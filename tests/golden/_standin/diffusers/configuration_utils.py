import functools
import inspect


class FrozenDict(dict):
    """dict with attribute access.  Unlike the real FrozenDict it allows attribute
    assignment, because reference callers mutate `scheduler.config.after_step`
    (pipeline_sampler_class_conditional_uncertainty_guided_gradient.py:64-66)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = [p for n, p in sig.parameters.items() if n != "self"]
        cfg = {p.name: p.default for p in params}
        for p, a in zip(params, args):
            cfg[p.name] = a
        cfg.update({k: v for k, v in kwargs.items() if k in cfg})
        init_kwargs = {k: v for k, v in kwargs.items() if k in cfg}
        self._internal_dict = FrozenDict(cfg)
        init(self, *args, **init_kwargs)

    return inner


class ConfigMixin:
    config_name = None

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__)
        accepted = {n for n in sig.parameters if n != "self"}
        init = {k: v for k, v in dict(config).items() if k in accepted}
        init.update({k: v for k, v in kwargs.items() if k in accepted})
        return cls(**init)
