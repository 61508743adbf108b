"""String -> scheduler factory.  Mirrors diffusion_uncertainty/schedulers_uncertainty/get_uncertainty_scheduler.py:13-40
(same `--scheduler-type` keys, same argparse attribute names, unknown keys fall through to MC-dropout).  Keys whose
scheduler is outside the hot-path scope (`flip_grad`: per-parameter gradients upsampled to a map, autograd only) raise
NotImplementedError naming the row instead of silently picking another scheduler."""
from .scheduling_dpm_2_uncertainty_centered import KDPM2SchedulerUncertaintyImagenetClassConditioned as _DPM2Centered
from .scheduling_ddim_flip import DDIMSchedulerUncertaintyImagenetClassConditioned as _Flip
from .scheduling_ddim_mc_dropout import DDIMSchedulerUncertaintyImagenetClassConditioned as _MCDropout
from .scheduling_ddim_uncertainty import DDIMSchedulerUncertaintyImagenetClassConditioned as _Uncertainty
from .scheduling_ddim_uncertainty_centered import DDIMSchedulerUncertaintyImagenetClassConditioned as _Centered
from .scheduling_ddim_uncertainty_centered_d import DDIMSchedulerUncertaintyImagenetClassConditioned as _CenteredD
from .scheduling_ddim_uncertainty_image import DDIMSchedulerUncertaintyImagenetClassConditioned as _Image
from .scheduling_ddim_uncertainty_zigzag_centered import DDIMSchedulerUncertaintyImagenetClassConditioned as _ZigZagCentered

_NOT_ON_PATH = {"flip_grad"}


def get_uncertainty_scheduler(args, y, unet, scheduler):
    kind = args.scheduler_type
    cfg = scheduler.config
    common = dict(after_step=args.start_step_uc, num_steps_uc=args.num_steps_uc, unet=unet, y=y, eta=getattr(args, "eta", 0.0))
    if kind in _NOT_ON_PATH:
        raise NotImplementedError(f"scheduler type {kind!r} is outside the accelerated uncertainty path (SURVEY.md §8f)")
    if kind == "flip":
        return _Flip.from_config(cfg, after_step=args.start_step_uc, num_steps_uc=args.num_steps_uc, unet=unet,
                                 eta=getattr(args, "eta", 0.0), prompt_embeds=y)
    if kind == "uncertainty":
        return _Uncertainty.from_config(cfg, M=args.M, predict_next=args.predict_next, **common)
    if kind == "uncertainty_image":
        return _Image.from_config(cfg, M=args.M, predict_next=args.predict_next, **common)
    if kind == "uncertainty_centered":
        return _Centered.from_config(cfg, M=args.M, predict_next=args.predict_next, **common)
    if kind == "uncertainty_original":
        return _Uncertainty.from_config(cfg, M=args.M, predict_next=False, **common)
    if kind == "uncertainty_centered_d":
        return _CenteredD.from_config(cfg, M=args.M, uncertainty_distance=args.uncertainty_distance, **common)
    if kind == "dpm_2_uncertainty_centered":
        # (the reference passes y= and eta= too; the DPM-2 constructor accepts neither, from_config drops them — :31-32)
        return _DPM2Centered.from_config(cfg, M=args.M, **common)
    if kind == "uncertainty_zigzag_centered":
        return _ZigZagCentered.from_config(cfg, M=args.M, num_zigzag=args.num_zigzag, **common)
    return _MCDropout.from_config(cfg, prompt_embeds=y, M=args.M, **{k: v for k, v in common.items() if k != "y"})


# aliases kept by the reference (get_uncertainty_scheduler.py:38-40)
instatiate_uc_scheduler = get_uncertainty_scheduler
instatiate_uncertainty_scheduler = get_uncertainty_scheduler
