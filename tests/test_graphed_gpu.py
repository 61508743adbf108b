"""CUDA-graph capture of whole window steps (diffusion_uncertainty_b200/graphed.py): one graph per timestep
holding the M x num_zigzag re-noise launches (noise from the device-resident Philox state), the model forwards, the reduction
and the DDIM update."""
import pytest
import torch

from tests.toy_models import ToyADM

pytestmark = pytest.mark.gpu

BASE = dict(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", clip_sample=True, set_alpha_to_one=True,
            steps_offset=0, prediction_type="epsilon", timestep_spacing="leading")


def build(dev):
    from diffusion_uncertainty_b200.schedulers_uncertainty.scheduling_ddim_uncertainty_zigzag_centered import \
        DDIMSchedulerUncertaintyImagenetClassConditioned as Sched
    model = ToyADM(3, seed=5).eval().to(dev)
    sched = Sched.from_config(BASE, unet=model, M=5, after_step=2, num_steps_uc=4, num_zigzag=3)
    sched.set_timesteps(10)
    sched.prompt_embeds = torch.arange(16, device=dev) % 10
    return model, sched


def test_graphed_window_step_matches_eager_and_continues_the_noise_stream():
    from diffusion_uncertainty_b200 import ops
    from diffusion_uncertainty_b200.graphed import GraphedScheduler
    dev = torch.device("cuda:0")
    model, sched = build(dev)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(16, 3, 32, 32, generator=g).to(dev)            # CIFAR-10 shape, batch 16 (BASELINE configs[0])
    t = sched.uncertainty_timesteps()[0]
    t_tensor = torch.full((16,), t, device=dev, dtype=torch.long)
    eps = model(x, t_tensor, y=sched.prompt_embeds)[:, :3]
    eager = sched.step(eps, t, x)
    gs = GraphedScheduler(sched, seed=123)
    n0 = ops.launch_count
    out1 = gs.step(eps, t, x)
    u1, p1 = out1.uncertainty.clone(), out1.prev_sample.clone()
    captured_launches = ops.launch_count - n0
    # the DDIM update does not depend on the noise: bit-identical to the eager step; the map is a different draw of the same law
    assert torch.equal(p1, eager.prev_sample)
    assert torch.isfinite(u1).all() and (u1 >= 0).all()
    assert abs(float(u1.mean()) / float(eager.uncertainty.mean()) - 1.0) < 0.2
    # a replay launches nothing from Python and continues the stream: new noise, new map, same x_(t-1)
    n1 = ops.launch_count
    out2 = gs.step(eps, t, x)
    assert ops.launch_count == n1 and gs.replays == 2
    assert torch.equal(out2.prev_sample, p1) and not torch.equal(out2.uncertainty, u1)
    u2 = out2.uncertainty.clone()
    # the same seed and offset reproduce the maps bit for bit (the state lives on the device; one advance per graph)
    gs2 = GraphedScheduler(sched, seed=123)
    assert torch.equal(gs2.step(eps, t, x).uncertainty, u1)
    assert torch.equal(gs2.step(eps, t, x).uncertainty, u2)
    # M x num_zigzag re-noise launches + one reduction + one DDIM launch + ONE rng advance (plus the warm-up runs before the capture)
    per_step = 5 * 3 + 1 + 1 + 1
    assert captured_launches == per_step * (1 + gs.warmup) - gs.warmup, captured_launches     # (warm-up steps run eagerly: no advance)
    # out-of-window steps and other inputs go through the same interface
    t_out = sched._host_timesteps[-1]
    o3 = gs.step(eps, t_out, x)
    assert torch.equal(o3.prev_sample, sched.step(eps, t_out, x).prev_sample)
