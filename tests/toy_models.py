"""Deterministic toy score models for parity tests.

They only use elementwise mul/add/roll/flip/clamp with one rounding per op, so the SAME bits come
out on CPU and on CUDA (no conv / matmul / transcendental whose rounding differs between devices).
All randomness goes through `torch.randn_like`, which `seeded_noise()` replaces with a seeded CPU
generator so that golden vectors recorded on CPU replay bit-for-bit on the GPU.
"""
import contextlib

import torch


@contextlib.contextmanager
def seeded_noise(seed: int):
    """Route torch.randn_like through a seeded CPU generator (device-independent noise)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    orig = torch.randn_like

    def randn_like(x, **kw):
        return torch.randn(x.shape, generator=g, dtype=torch.float32).to(device=x.device, dtype=kw.get("dtype", x.dtype))

    torch.randn_like = randn_like
    try:
        yield g
    finally:
        torch.randn_like = orig


class DetDropout(torch.nn.Dropout):
    """nn.Dropout subclass (so the reference's isinstance check passes) whose mask is drawn with
    torch.randn_like -> reproducible under seeded_noise() on any device.  p = 0.5 only."""

    def forward(self, x):
        if not self.training:
            return x
        keep = (torch.randn_like(x) > 0).to(x.dtype)
        return x * keep * 2.0


class ToyADM(torch.nn.Module):
    """ADM-shaped stand-in: forward(x[B,C,H,W], t[B] int64, y=[B] int64) -> [B, 2C, H, W]
    (callers slice [:, :C], giving the strided view the real ADM gives)."""

    def __init__(self, channels: int = 3, seed: int = 0, dropout: bool = False, scale: float = 1.0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        c2 = 2 * channels
        self.register_buffer("a", torch.randn(c2, 1, 1, generator=g) * 0.5 * scale)
        self.register_buffer("b", torch.randn(c2, 1, 1, generator=g) * 0.25 * scale)
        self.register_buffer("c", torch.randn(c2, 1, 1, generator=g) * 0.5)
        self.register_buffer("d", torch.randn(c2, 1, 1, generator=g) * 0.125)
        self.drop = DetDropout(0.5) if dropout else None

    def forward(self, x, t, y=None, **kw):
        if not torch.is_tensor(t):
            t = torch.full((x.shape[0],), int(t), device=x.device, dtype=torch.long)
        x = x.float()
        x6 = torch.cat([x, x.flip(-1)], dim=1)
        h = self.a * x6
        h = h + self.b * (torch.roll(x6, 1, -1) * torch.roll(x6, 1, -2))
        if self.drop is not None:
            h = self.drop(h)
        tt = (t.float() * (1.0 / 1024.0)).view(-1, 1, 1, 1)
        h = h + self.c * tt
        if y is not None and y.dtype in (torch.int64, torch.int32):
            h = h + self.d * (y.float() * 0.125).view(-1, 1, 1, 1)
        return h.clamp(-3.0, 3.0)


class ToyADMWithParameter(ToyADM):
    """ToyADM plus one trainable (zero) parameter in the graph, so that the output requires grad even when the input does not —
    like a real score model; the reference's guided-gradient pipeline back-propagates the raw prediction first
    (pipeline_sampler_class_conditional_uncertainty_guided_gradient.py:182-183)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.p = torch.nn.Parameter(torch.zeros(1))

    def forward(self, x, t, y=None, **kw):
        return super().forward(x, t, y=y, **kw) + self.p * 0.0


class ToySDUNet(torch.nn.Module):
    """diffusers-UNet2DConditionModel-shaped stand-in: called by keyword, returns a 1-tuple."""

    def __init__(self, channels: int = 4, seed: int = 1):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("a", torch.randn(channels, 1, 1, generator=g) * 0.5)
        self.register_buffer("b", torch.randn(channels, 1, 1, generator=g) * 0.25)
        self.register_buffer("c", torch.randn(channels, 1, 1, generator=g) * 0.5)

    def forward(self, sample, timestep, encoder_hidden_states=None, **kw):
        x = sample.float()
        h = self.a * x + self.b * (torch.roll(x, 1, -1) * torch.roll(x, 2, -2))
        tt = torch.as_tensor(timestep, device=x.device).float().reshape(-1)[:1] * (1.0 / 1024.0)
        h = h + self.c * tt
        if encoder_hidden_states is not None:
            e = encoder_hidden_states.float().reshape(encoder_hidden_states.shape[0], -1)[:, :1]
            h = h + (e * 0.25).view(-1, 1, 1, 1)
        return (h.clamp(-3.0, 3.0),)


class _SampleOutput:
    def __init__(self, sample):
        self.sample = sample


class ToyUNet2D(ToyADM):
    """diffusers-UNet2DModel-shaped stand-in for the unconditioned (CIFAR-10) loop: `model(x, t).sample` is the 6-channel
    ToyADM output (generate_samples.py:414 slices `[:, :3]`; the Cifar10 scheduler classes call `unet(x, t).sample`)."""

    def forward(self, x, t, y=None, **kw):
        return _SampleOutput(super().forward(x, t, y=None))


class ToyUNet2D3(ToyUNet2D):
    """the same with a 3-channel `.sample` (what the scheduler's own forwards need: Cifar10.predict_model does not slice)"""

    def forward(self, x, t, y=None, **kw):
        return _SampleOutput(ToyADM.forward(self, x, t, y=None)[:, :3])


class ToyUViTMixin:
    """U-ViT-with-autoencoder-shaped stand-in: `model(z[B,4,h,w], t[B], y[B]) -> [B,4,h,w]` (class label positional, no channel
    slice: traits.py:12-13) and `decode(z) -> [B,3,2h,2w]`.  Mixed into a class NAMED UViTAE (the drop-in dispatches on the
    class name) or into a subclass of the reference's UViTAE (tests/golden/make_golden.py)."""

    def toy_init(self, seed: int = 0):
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("ta", torch.randn(4, 1, 1, generator=g) * 0.5)
        self.register_buffer("tb", torch.randn(4, 1, 1, generator=g) * 0.25)
        self.register_buffer("tc", torch.randn(4, 1, 1, generator=g) * 0.5)
        self.register_buffer("td", torch.randn(4, 1, 1, generator=g) * 0.125)
        self.register_buffer("dec", torch.randn(3, 4, 1, 1, generator=g) * 0.5)

    def forward(self, x, t, y=None, **kw):
        x = x.float()
        h = self.ta * x + self.tb * (torch.roll(x, 1, -1) * torch.roll(x, 1, -2))
        h = h + self.tc * (t.float() * (1.0 / 1024.0)).view(-1, 1, 1, 1)
        if y is not None:
            h = h + self.td * (y.float() * 0.125).view(-1, 1, 1, 1)
        return h.clamp(-3.0, 3.0)

    def decode(self, z):
        img = (self.dec.unsqueeze(0) * z.float().unsqueeze(1)).sum(dim=2)          # 1x1 mix of the 4 latent channels -> 3
        return torch.repeat_interleave(torch.repeat_interleave(img, 2, dim=-1), 2, dim=-2)


class UViTAE(ToyUViTMixin, torch.nn.Module):
    def __init__(self, seed: int = 0):
        torch.nn.Module.__init__(self)
        self.toy_init(seed)
