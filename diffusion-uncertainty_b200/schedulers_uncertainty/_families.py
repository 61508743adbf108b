"""The four public classes every reference scheduler module exports, built from one variant class.

Reference pattern (e.g. scheduling_ddim_uncertainty_zigzag_centered.py:138, 653-661; scheduling_ddim_mc_dropout.py:639-654):
  DDIMSchedulerUncertainty                         diffusers-UNet call convention
  DDIMSchedulerUncertaintyImagenet                 unet(x, t)[:, :3]
  DDIMSchedulerUncertaintyCifar10                  unet(x, t).sample
  DDIMSchedulerUncertaintyImagenetClassConditioned PredictorClassConditionedTrait dispatch, class_conditioned = True
"""
import torch

from .traits import PredictorClassConditionedTrait


def make_family(variant_cls, module_name: str):
    class DDIMSchedulerUncertainty(variant_cls):
        pass

    class DDIMSchedulerUncertaintyImagenet(DDIMSchedulerUncertainty):
        def predict_model(self, x, t):
            if isinstance(t, int):
                t = torch.zeros(x.shape[0], dtype=torch.int64, device=x.device).fill_(t)
            return self.unet(x, t)[:, :3]

    class DDIMSchedulerUncertaintyCifar10(DDIMSchedulerUncertainty):
        def predict_model(self, x, t):
            return self.unet(x, t).sample

    class DDIMSchedulerUncertaintyImagenetClassConditioned(PredictorClassConditionedTrait, DDIMSchedulerUncertainty):
        class_conditioned: bool = True

    out = {}
    for cls in (DDIMSchedulerUncertainty, DDIMSchedulerUncertaintyImagenet, DDIMSchedulerUncertaintyCifar10,
                DDIMSchedulerUncertaintyImagenetClassConditioned):
        cls.__module__ = module_name
        cls.__qualname__ = cls.__name__
        cls.__doc__ = variant_cls.__doc__
        out[cls.__name__] = cls
    return out
