"""Score-model call dispatch of the class-conditioned schedulers.

Mirrors diffusion_uncertainty/schedulers_uncertainty/traits.py:6-18.  The score models stay the reference's PyTorch
modules (ADM `UNetModel`, `UViT` / `UViTAE`, diffusers `UNet2DModel`); they are recognised by class NAME anywhere in the
MRO so that this package does not have to import diffusers or the reference to dispatch."""
import torch

_UVIT_NAMES = {"UViT", "UViTAE"}
_UNET2D_NAMES = {"UNet2DModel"}


def _mro_names(obj):
    return {c.__name__ for c in type(obj).__mro__}


class PredictorClassConditionedTrait:

    def predict_model(self, x, t):
        if isinstance(t, float):
            t = round(t)
        if isinstance(t, int):
            t = torch.zeros(size=(x.shape[0],), dtype=torch.int64, device=x.device).fill_(t)
        names = _mro_names(self.unet)
        if names & _UVIT_NAMES:
            return self.unet(x, t, self.prompt_embeds)
        if names & _UNET2D_NAMES:
            return self.unet(x, t).sample
        return self.unet(x, t, y=self.prompt_embeds)[:, :3]   # ADM: learned-sigma head dropped, a strided view
