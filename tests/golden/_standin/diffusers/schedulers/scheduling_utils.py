from dataclasses import dataclass
from enum import Enum

import torch

from ..utils import BaseOutput


class KarrasDiffusionSchedulers(Enum):
    DDIMScheduler = 1
    DDPMScheduler = 2
    PNDMScheduler = 3


@dataclass
class SchedulerOutput(BaseOutput):
    prev_sample: torch.Tensor


class SchedulerMixin:
    config_name = "scheduler_config.json"
    _compatibles = []
    has_compatibles = True
