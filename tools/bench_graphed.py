#!/usr/bin/env python
"""Eager against CUDA-graph-replayed window steps at the launch-bound BASELINE shapes: the zig-zag-centred
scheduler, M = 5, num_zigzag = 3, a cheap elementwise score model (so that what is measured is the scheduler's own launches and
host work, not the model).  One JSON line per shape."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_uncertainty_b200 import ops  # noqa: E402
from diffusion_uncertainty_b200.graphed import GraphedScheduler  # noqa: E402
from diffusion_uncertainty_b200.schedulers_uncertainty.scheduling_ddim_uncertainty_zigzag_centered import \
    DDIMSchedulerUncertaintyImagenetClassConditioned as Sched  # noqa: E402
from tests.toy_models import ToyADM  # noqa: E402

BASE = dict(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", clip_sample=True, set_alpha_to_one=True,
            steps_offset=0, prediction_type="epsilon", timestep_spacing="leading")


def main():
    dev = torch.device("cuda:0")
    for name, shape in (("cifar10_b16", (16, 3, 32, 32)), ("sd_latent_b1", (1, 3, 64, 64)), ("imagenet64_b128", (128, 3, 64, 64))):
        model = ToyADM(3, seed=5).eval().to(dev)
        sched = Sched.from_config(BASE, unet=model, M=5, after_step=40, num_steps_uc=10, num_zigzag=3)
        sched.set_timesteps(50)
        sched.prompt_embeds = torch.arange(shape[0], device=dev) % 10
        x = torch.randn(shape, device=dev)
        t = sched.uncertainty_timesteps()[0]
        t_tensor = torch.full((shape[0],), t, device=dev, dtype=torch.long)
        eps = model(x, t_tensor, y=sched.prompt_embeds)[:, :3]
        gs = GraphedScheduler(sched, seed=1)

        def timed(fn, reps=50):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps * 1e6

        n0 = ops.launch_count
        sched.step(eps, t, x)
        launches = ops.launch_count - n0
        eager_us = timed(lambda: sched.step(eps, t, x))
        graph_us = timed(lambda: gs.step(eps, t, x))
        print(json.dumps({"shape": name, "dims": shape, "window_step_eager_us": round(eager_us, 1), "window_step_graphed_us": round(graph_us, 1),
                          "speedup": round(eager_us / graph_us, 2), "library_launches_per_step": launches,
                          "model_forwards_per_step": 15, "timing": "wall clock around 50 steps, device synchronised"}), flush=True)


if __name__ == "__main__":
    main()
