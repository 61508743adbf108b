"""SURVEY.md §8f N1 — perturbation with the noise drawn in the kernel (du_perturb_randn).

Contract: the variates are BIT-IDENTICAL to what `torch.randn_like` yields on the same CUDA device for the same generator
state, and the generator is left in the same state; so a sampling loop that draws its noise in the kernel replays the
trajectory of the loop that calls torch.randn_like (the reference's order of draws is unchanged).  torch's generator is the
library that defines the stream here (as torch.quantile defines F2a); the combination a*x + b*n is checked against
du_perturb, which the oracle pins.
"""
import pytest
import torch

from tests.helpers import l4_sampling_loop
from tests.test_oracle_golden import SCHED_CASES, T, load
from tests.toy_models import ToyADM

pytestmark = pytest.mark.gpu


def dev():
    torch.cuda.init()      # torch.cuda.default_generators is empty until CUDA is initialised
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from diffusion_uncertainty_b200 import ops as o
    return o


def bits(a, b):
    a, b = a.detach().cpu().contiguous(), b.detach().cpu().contiguous()
    if a.dtype == torch.float32:
        return a.shape == b.shape and torch.equal(a.view(torch.int32), b.view(torch.int32))
    return a.shape == b.shape and torch.equal(a.view(torch.int16), b.view(torch.int16))


def grid_cap():
    p = torch.cuda.get_device_properties(0)
    return p.multi_processor_count * (p.max_threads_per_multi_processor // 256)


SIZES = [1, 3, 7, 255, 256, 257, 1000, 4096, 16384, 49152, 3 * 64 * 64 * 5 + 1]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_randn_equals_torch_bitwise_small_and_ragged(ops, dtype):
    dev()
    gen = torch.cuda.default_generators[0]
    for n in SIZES:
        x = torch.empty(n, device=dev(), dtype=dtype)
        torch.manual_seed(1234 + n)
        want = torch.randn_like(x)
        off_want = gen.get_offset()
        torch.manual_seed(1234 + n)
        got = ops.randn_like(x)
        assert bits(got, want), (n, dtype)
        assert gen.get_offset() == off_want, "the generator must be advanced exactly as torch's launch advances it"


def test_randn_equals_torch_beyond_one_grid_stride(ops):
    """sizes around G*256*4 (one full pass of the capped grid) and the BASELINE shapes: several curand_normal4 calls per
    virtual thread, ragged last pass"""
    dev()
    cap = grid_cap() * 256 * 4
    gen = torch.cuda.default_generators[0]
    for n in [cap - 1, cap, cap + 1, 2 * cap + 12345, 128 * 3 * 64 * 64, 128 * 3 * 128 * 128]:
        x = torch.empty(n, device=dev())
        torch.manual_seed(7)
        want = torch.randn_like(x)
        off = gen.get_offset()
        torch.manual_seed(7)
        got = ops.randn_like(x)
        assert bits(got, want), n
        assert gen.get_offset() == off


def test_draws_interleave_with_torch_draws(ops):
    """torch draw, kernel draw, torch draw == three torch draws (nonzero starting offsets, 4-D shapes)"""
    shape = (8, 3, 32, 32)
    torch.manual_seed(99)
    w = [torch.randn(shape, device=dev()) for _ in range(3)]
    torch.manual_seed(99)
    a = torch.randn(shape, device=dev())
    b = ops.randn_like(a)
    c = torch.randn(shape, device=dev())
    assert bits(a, w[0]) and bits(b, w[1]) and bits(c, w[2])


def test_explicit_generator(ops):
    g1 = torch.Generator(device=dev()).manual_seed(5)
    g2 = torch.Generator(device=dev()).manual_seed(5)
    x = torch.empty(4, 4, 16, 16, device=dev())
    torch.randn(3, device=dev(), generator=g1); torch.randn(3, device=dev(), generator=g2)
    want = torch.randn(x.shape, device=dev(), generator=g1)
    got = ops.randn_like(x, generator=g2)
    assert bits(got, want) and g1.get_offset() == g2.get_offset()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_perturb_randn_equals_randn_then_perturb(ops, dtype):
    x = torch.randn(16, 3, 32, 32, device=dev()).to(dtype)
    a, b = 0.9949874, 0.1
    torch.manual_seed(11)
    noise = torch.randn_like(x)
    want = ops.perturb(x, noise, a, b)
    torch.manual_seed(11)
    got, n2 = ops.perturb_randn(x, a, b, want_noise=True)
    assert bits(n2, noise) and bits(got, want)
    torch.manual_seed(11)
    assert bits(ops.perturb_fresh(x, a, b), want)
    # a strided view cannot take the in-kernel draw: perturb_fresh falls back to torch.randn_like + du_perturb, same values
    big = torch.randn(16, 6, 32, 32, device=dev()).to(dtype)
    view = big[:, :3]
    torch.manual_seed(12)
    nz = torch.randn_like(view)
    want_v = ops.perturb(view, nz, a, b)
    torch.manual_seed(12)
    assert bits(ops.perturb_fresh(view, a, b), want_v)


def test_replaced_randn_like_is_honoured(ops):
    """a caller that swaps torch.randn_like (the tests' seeded CPU noise) must get ITS noise"""
    from tests.toy_models import seeded_noise
    x = torch.randn(4, 3, 8, 8, device=dev())
    with seeded_noise(3):
        assert not ops.randn_fusable(x)
        got = ops.perturb_fresh(x, 0.5, 2.0)
    with seeded_noise(3):
        want = ops.perturb(x, torch.randn_like(x), 0.5, 2.0)
    assert bits(got, want)
    assert ops.randn_fusable(x)


def test_device_rng_replays_inside_a_cuda_graph(ops):
    """draws captured in a CUDA graph: the {seed, offset} pair lives on the device and is advanced inside the graph, so each
    replay continues torch's stream for that seed"""
    x = torch.randn(8, 3, 32, 32, device=dev())
    seed = 4242
    rng = ops.DeviceRng(dev(), seed)
    out = torch.empty(2, *x.shape, device=dev())
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        rng.perturb(x, 0.75, 0.5)                  # warm-up outside capture
    torch.cuda.current_stream().wait_stream(s)
    rng.state[1] = 0
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out[0].copy_(rng.perturb(x, 0.75, 0.5))
        out[1].copy_(rng.perturb(x, 0.75, 0.5))
    torch.manual_seed(seed)
    for replay in range(3):
        g.replay()
        torch.cuda.synchronize()
        for k in range(2):
            want = ops.perturb(x, torch.randn_like(x), 0.75, 0.5)
            assert bits(out[k], want), (replay, k)
    assert rng.offset() == torch.cuda.default_generators[0].get_offset()


def _run_loop(module, cls, kw, fuse, seed):
    import importlib
    from diffusion_uncertainty_b200 import ops as o
    mod = importlib.import_module("diffusion_uncertainty_b200.schedulers_uncertainty." + module)
    model = ToyADM(3, seed=1).to(dev())
    sched = getattr(mod, cls).from_config(
        dict(num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear", clip_sample=True,
             set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon", timestep_spacing="leading"), unet=model, **kw)
    sched.set_timesteps(10)
    g = torch.Generator().manual_seed(seed)
    x_T = torch.randn(4, 3, 16, 16, generator=g).to(dev())
    y = torch.randint(0, 10, (4,), generator=g).to(dev())
    orig = o.randn_fusable
    if not fuse:
        o.randn_fusable = lambda *a, **k: False
    try:
        torch.manual_seed(seed)
        launches = o.launch_count
        res = l4_sampling_loop(sched, model, x_T, y)
        return res, o.launch_count - launches
    finally:
        o.randn_fusable = orig


@pytest.mark.parametrize("module,kw", [
    ("scheduling_ddim_uncertainty_zigzag_centered", dict(M=3, after_step=4, num_steps_uc=4, num_zigzag=2)),
    ("scheduling_ddim_uncertainty_centered", dict(M=3, after_step=4, num_steps_uc=4)),
    ("scheduling_ddim_uncertainty_centered_d", dict(M=2, after_step=4, num_steps_uc=3, uncertainty_distance=2)),
    ("scheduling_ddim_uncertainty_image", dict(M=3, after_step=4, num_steps_uc=3)),
])
def test_scheduler_trajectory_is_unchanged_by_the_in_kernel_draw(module, kw):
    """the whole loop with torch's CUDA generator: in-kernel draws vs torch.randn_like + du_perturb — identical bits"""
    fused, n_f = _run_loop(module, "DDIMSchedulerUncertaintyImagenetClassConditioned", kw, True, 21)
    plain, n_p = _run_loop(module, "DDIMSchedulerUncertaintyImagenetClassConditioned", kw, False, 21)
    assert bits(fused["final"], plain["final"]) and bits(fused["uncertainty"], plain["uncertainty"])
    assert n_f == n_p, "one du_ launch per perturbation either way (the torch generator launch is what disappears)"


def test_percentile_guidance_function_with_in_kernel_draws(ops):
    """uncertainty_guidance.py:83-120 with torch's CUDA generator: the M draws inside du_perturb_randn leave the result unchanged"""
    import diffusion_uncertainty_b200.uncertainty_guidance as ug
    from tests.toy_models import ToySDUNet
    sd = ToySDUNet(4, seed=12).eval().to(dev())
    g = torch.Generator().manual_seed(8)
    lat2 = torch.cat([torch.randn(1, 4, 32, 32, generator=g)] * 2).to(dev())
    emb = torch.randn(2, 7, 16, generator=g).to(dev())
    t_tensor = torch.tensor(500, device=dev())
    un, tx = sd(lat2, t_tensor, emb)[0].chunk(2)
    eps = (un + 7.5 * (tx - un)).detach()
    outs = []
    for fuse in (True, False):
        orig = ops.randn_fusable
        if not fuse:
            ops.randn_fusable = lambda *a, **k: False
        try:
            torch.manual_seed(31)
            ug.use_posterior = True
            outs.append(ug.get_uncertainty_guided_score_with_percentile(eps.clone(), lat2.clone(), t_tensor, emb.clone(), sd,
                                                                        torch.tensor(0.5), 0.9, "stable-diffusion",
                                                                        num_uncertainty_samples=3, guidance_scale=7.5))
        finally:
            ops.randn_fusable = orig
    assert bits(outs[0], outs[1])
