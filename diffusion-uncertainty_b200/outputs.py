"""Return type of `scheduler.step()`.

The reference's `DDIMSchedulerUncertaintyOutput` is a diffusers `BaseOutput` dataclass
(schedulers_uncertainty/scheduling_ddim_uncertainty_zigzag_centered.py:37-54) on which `step()` sets extra attributes
after construction (`uncertainty`, `pred_epsilon` at :555-557; `score`, `noise` in scheduling_ddim_mc_dropout.py:53-57).
Callers use attribute access (`output.prev_sample`, generate_samples.py:187-201), item access by name
(`["prev_sample"]`, pipeline_stable_diffusion_uncertainty_guided.py:778) and tuple indexing (`[0]`).  All three work
here; fields that are None are absent from the mapping / tuple view, as with BaseOutput.
"""
from __future__ import annotations

from typing import Any, Iterator, Optional, Tuple

import torch

_FIELDS = ("prev_sample", "pred_original_sample", "uncertainty", "pred_epsilon", "score", "noise")


class DDIMSchedulerUncertaintyOutput:
    __slots__ = _FIELDS + ("_extra",)

    def __init__(self, prev_sample: torch.Tensor, pred_original_sample: Optional[torch.Tensor] = None,
                 uncertainty: Optional[torch.Tensor] = None, pred_epsilon: Optional[torch.Tensor] = None,
                 score: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None):
        object.__setattr__(self, "_extra", {})
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample
        self.uncertainty = uncertainty
        self.pred_epsilon = pred_epsilon
        self.score = score
        self.noise = noise

    # dynamic attributes, like the reference's `output.foo = ...` on a BaseOutput
    def __setattr__(self, name: str, value: Any) -> None:
        if name in _FIELDS:
            object.__setattr__(self, name, value)
        else:
            self._extra[name] = value

    def __getattr__(self, name: str) -> Any:  # only reached for names outside __slots__
        extra = object.__getattribute__(self, "_extra")
        if name in extra:
            return extra[name]
        raise AttributeError(name)

    def keys(self):
        return [k for k in _FIELDS if getattr(self, k) is not None] + [k for k, v in self._extra.items() if v is not None]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]

    def __iter__(self) -> Iterator[str]:
        return iter(self.keys())

    def __len__(self) -> int:
        return len(self.keys())

    def __contains__(self, k) -> bool:
        return k in self.keys()

    def __getitem__(self, k):
        if isinstance(k, str):
            if k not in self.keys():
                raise KeyError(k)
            return getattr(self, k)
        return self.to_tuple()[k]

    def __setitem__(self, k: str, v) -> None:
        setattr(self, k, v)

    def to_tuple(self) -> Tuple[Any, ...]:
        return tuple(self[k] for k in self.keys())

    def __repr__(self) -> str:
        parts = ", ".join(f"{k}={tuple(v.shape) if torch.is_tensor(v) else v!r}" for k, v in self.items())
        return f"DDIMSchedulerUncertaintyOutput({parts})"


# the non-uncertainty schedulers of the reference return diffusers' SchedulerOutput; same access patterns
SchedulerOutput = DDIMSchedulerUncertaintyOutput
