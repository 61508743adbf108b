from collections import OrderedDict
from dataclasses import fields, is_dataclass


class BaseOutput(OrderedDict):
    """Dataclass-backed ordered dict: attribute and ["key"] access, like diffusers'."""

    def __init_subclass__(cls) -> None:
        super().__init_subclass__()

    def __post_init__(self):
        if is_dataclass(self):
            for f in fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    OrderedDict.__setitem__(self, f.name, v)

    def __setattr__(self, name, value):
        if value is not None and (name in self.keys() or not name.startswith("_")):
            OrderedDict.__setitem__(self, name, value)
        super().__setattr__(name, value)

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return tuple(self.values())[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())
