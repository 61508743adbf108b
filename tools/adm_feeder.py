"""ADM-shaped random-init score model — BENCH FEEDER, not part of the product.

north_star keeps the score models "the reference PyTorch modules that feed" the uncertainty path; those modules
(`diffusion_uncertainty/guided_diffusion/unet_openai.py`) live in /root/reference, which does not exist on the GPU box and may
not be copied.  The img/s metric (BASELINE.json: "ImageNet-128 M=5 img/s at 1/2/4/8 GPU") needs a model of the named shape in
the loop, so this file builds one from the published architecture description (Dhariwal & Nichol 2021, "Diffusion Models
Beat GANs", app. I: residual UNet, BigGAN up/down-sampling residual blocks, adaptive group norm = scale-shift conditioning,
multi-head self-attention at the listed resolutions, timestep + class embedding) with the hyper-parameters the reference
passes (`diffusion_uncertainty/init_model.py:21` ImageNet-128: model_channels 256, channel_mult (1,1,2,3,4), 2 residual blocks
per level, attention at down-sampling rates 4/8/16, 4 heads, learned sigma -> 6 output channels, 1000 classes;
`init_model.py:45-47` ImageNet-64: 192 channels, (1,2,3,4), 3 blocks, rates 2/4/8, 64 channels per head, dropout).
Weights are random (no checkpoint; the zero-initialised output layers of the original are given small random weights so
that the scores are not identically zero).  Call convention: `model(x, t[B] int64, y=[B] int64) -> [B, 6, H, W]`.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def sinusoidal_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, device=t.device, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class ResidualBlock(nn.Module):
    """GN-SiLU-conv, scale-shift conditioning on the embedding, GN-SiLU-dropout-conv, learned skip when channels change;
    `resample` in {None, 'up', 'down'} resamples both branches (BigGAN-style block)."""

    def __init__(self, c_in, c_out, c_emb, dropout, resample=None):
        super().__init__()
        self.resample = resample
        self.norm1 = GroupNorm32(32, c_in)
        self.conv1 = nn.Conv2d(c_in, c_out, 3, padding=1)
        self.emb = nn.Linear(c_emb, 2 * c_out)
        self.norm2 = GroupNorm32(32, c_out)
        self.drop = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(c_out, c_out, 3, padding=1)
        self.skip = nn.Identity() if c_in == c_out else nn.Conv2d(c_in, c_out, 1)

    def _resample(self, x):
        if self.resample == "up":
            return F.interpolate(x, scale_factor=2, mode="nearest")
        if self.resample == "down":
            return F.avg_pool2d(x, 2)
        return x

    def forward(self, x, emb):
        h = F.silu(self.norm1(x))
        h, x = self._resample(h), self._resample(x)
        h = self.conv1(h)
        scale, shift = self.emb(F.silu(emb)).type(h.dtype)[:, :, None, None].chunk(2, dim=1)
        h = self.norm2(h) * (1 + scale) + shift
        h = self.conv2(self.drop(F.silu(h)))
        return self.skip(x) + h


class SelfAttention(nn.Module):
    def __init__(self, c, heads):
        super().__init__()
        self.heads = heads
        self.norm = GroupNorm32(32, c)
        self.qkv = nn.Conv1d(c, 3 * c, 1)
        self.proj = nn.Conv1d(c, c, 1)

    def forward(self, x, emb=None):
        B, C, H, W = x.shape
        qkv = self.qkv(self.norm(x).reshape(B, C, H * W))
        q, k, v = qkv.reshape(B, 3, self.heads, C // self.heads, H * W).unbind(1)
        a = F.scaled_dot_product_attention(q.transpose(-1, -2), k.transpose(-1, -2), v.transpose(-1, -2))
        a = a.transpose(-1, -2).reshape(B, C, H * W)
        return x + self.proj(a).reshape(B, C, H, W)


class Stage(nn.ModuleList):
    def forward(self, x, emb):
        for m in self:
            x = m(x, emb)
        return x


class ADMFeeder(nn.Module):
    def __init__(self, image_size=128, in_channels=3, model_channels=256, out_channels=6, num_res_blocks=2,
                 attention_rates=(4, 8, 16), dropout=0.0, channel_mult=(1, 1, 2, 3, 4), num_classes=1000, num_heads=4,
                 num_head_channels=-1, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        mc, c_emb = model_channels, 4 * model_channels
        self.model_channels = mc
        heads = (lambda c: num_heads) if num_head_channels == -1 else (lambda c: c // num_head_channels)
        self.time_embed = nn.Sequential(nn.Linear(mc, c_emb), nn.SiLU(), nn.Linear(c_emb, c_emb))
        self.label_emb = nn.Embedding(num_classes, c_emb)
        self.stem = nn.Conv2d(in_channels, mc, 3, padding=1)
        self.down = nn.ModuleList()
        skips, ch, rate = [mc], mc, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                blk = [ResidualBlock(ch, mult * mc, c_emb, dropout)]
                ch = mult * mc
                if rate in attention_rates:
                    blk.append(SelfAttention(ch, heads(ch)))
                self.down.append(Stage(blk))
                skips.append(ch)
            if level != len(channel_mult) - 1:
                self.down.append(Stage([ResidualBlock(ch, ch, c_emb, dropout, resample="down")]))
                skips.append(ch)
                rate *= 2
        self.mid = Stage([ResidualBlock(ch, ch, c_emb, dropout), SelfAttention(ch, heads(ch)), ResidualBlock(ch, ch, c_emb, dropout)])
        self.up = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                blk = [ResidualBlock(ch + skips.pop(), mult * mc, c_emb, dropout)]
                ch = mult * mc
                if rate in attention_rates:
                    blk.append(SelfAttention(ch, heads(ch)))
                if level and i == num_res_blocks:
                    blk.append(ResidualBlock(ch, ch, c_emb, dropout, resample="up"))
                    rate //= 2
                self.up.append(Stage(blk))
        self.out_norm = GroupNorm32(32, ch)
        self.out_conv = nn.Conv2d(ch, out_channels, 3, padding=1)

    def forward(self, x, timesteps, y=None, **kw):
        if not torch.is_tensor(timesteps):
            timesteps = torch.full((x.shape[0],), int(timesteps), device=x.device, dtype=torch.long)
        timesteps = timesteps.reshape(-1).expand(x.shape[0])
        emb = self.time_embed(sinusoidal_embedding(timesteps, self.model_channels))
        if y is not None:
            emb = emb + self.label_emb(y)
        h = self.stem(x)
        hs = [h]
        for stage in self.down:
            h = stage(h, emb)
            hs.append(h)
        h = self.mid(h, emb)
        for stage in self.up:
            h = stage(torch.cat([h, hs.pop()], dim=1), emb)
        return self.out_conv(F.silu(self.out_norm(h)))


def adm_imagenet128(seed=0):
    """hyper-parameters of diffusion_uncertainty/init_model.py:21"""
    return ADMFeeder(128, 3, 256, 6, 2, (4, 8, 16), 0.0, (1, 1, 2, 3, 4), 1000, 4, -1, seed)


def adm_imagenet64(dropout=0.5, seed=0):
    """hyper-parameters of diffusion_uncertainty/init_model.py:45-47 (dropout 0.5: BASELINE.json configs[1])"""
    return ADMFeeder(64, 3, 192, 6, 3, (2, 4, 8), dropout, (1, 2, 3, 4), 1000, 4, 64, seed)


if __name__ == "__main__":
    m = adm_imagenet128()
    print("ADM-128 feeder parameters: %.1f M" % (sum(p.numel() for p in m.parameters()) / 1e6))
    m64 = adm_imagenet64()
    print("ADM-64 feeder parameters: %.1f M" % (sum(p.numel() for p in m64.parameters()) / 1e6))
    with torch.no_grad():
        o = m64(torch.randn(1, 3, 64, 64), torch.tensor([10]), y=torch.tensor([3]))
    print(o.shape, float(o.std()))
