"""CUDA-graph capture of whole scheduler steps (second half).

A window step of an uncertainty scheduler is M (x num_zigzag) small launches around M model forwards: re-noise, forward, reduce,
DDIM.  At the small BASELINE shapes (CIFAR-10 b16: 49 k elements; the SD latent: 16 k) every one of those kernels runs for 2-4 us
while its launch from Python costs 5-10 us, and the reference adds a `t.item()` sync per step (generate_samples.py:178-179).
`GraphedScheduler` captures `scheduler.step(model_output, t, sample)` ONCE PER TIMESTEP (the host scalars of a step are kernel
arguments, so each timestep is its own graph) into a CUDA graph with static input / output buffers — model forwards included,
they are ordinary torch modules — and replays it: one launch per step from the host.

Noise: inside a capture torch's generator cannot be advanced by the host, so the perturbation draws come from a device-resident
Philox state (ops.DeviceRng, deferred mode: one du_rng_advance at the end of the graph).  Every replay continues that stream.  The
values are torch's normal variates for (seed, offset) — the same generator algorithm, a stream of its own.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops


class GraphedScheduler:
    """Wraps an uncertainty scheduler; `step()` has the scheduler's signature (model_output, timestep, sample) and returns the same
    output object, whose tensors are STATIC buffers overwritten by the next replay of the same timestep."""

    def __init__(self, scheduler, seed: int = 0, warmup: int = 2):
        self.scheduler = scheduler
        self.seed = seed
        self.warmup = warmup
        self._graphs: Dict[tuple, tuple] = {}
        self._rng: Optional[ops.DeviceRng] = None
        self.replays = 0

    def __getattr__(self, name):
        return getattr(self.scheduler, name)

    def _capture(self, key, model_output, t, sample):
        dev = sample.device
        if self._rng is None:
            self._rng = ops.DeviceRng(dev, self.seed, 0, deferred=True)
        s_mo, s_x = model_output.clone(), sample.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up on a side stream (allocator, lazy module init, autotuning)
            for _ in range(self.warmup):
                self.scheduler.step(s_mo, t, s_x)
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        prev_rng, ops.capture_rng = ops.capture_rng, self._rng
        try:
            with torch.cuda.graph(g):
                out = self.scheduler.step(s_mo, t, s_x)
                self._rng.commit(s_x)
        finally:
            ops.capture_rng = prev_rng
        self._graphs[key] = (g, s_mo, s_x, out)
        return self._graphs[key]

    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor, **kw):
        if kw.get("eta", 0.0) or kw.get("generator") is not None or kw.get("variance_noise") is not None:
            return self.scheduler.step(model_output, timestep, sample, **kw)      # stochastic DDIM steps stay eager
        t = int(timestep)
        key = (t, tuple(sample.shape), sample.dtype, tuple(model_output.shape), model_output.dtype, tuple(model_output.stride()))
        hit = self._graphs.get(key)
        if hit is None:
            hit = self._capture(key, model_output, t, sample)
        g, s_mo, s_x, out = hit
        s_mo.copy_(model_output)
        s_x.copy_(sample)
        g.replay()
        self.replays += 1
        return out
