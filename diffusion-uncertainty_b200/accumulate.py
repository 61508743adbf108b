"""F8 — uncertainty-map accumulation.

The reference appends `output.uncertainty.cpu()` (and `pred_epsilon.cpu()`) every in-window step — a synchronous
pageable D2H of the whole map — then `torch.stack(dim=1)` per batch and `torch.cat(dim=0)` over batches
(diffusion_uncertainty/generate_samples.py:189-201, 229-231; pipeline_uncertainty/pipeline_sampler_class_conditional_uncertainty.py:49-63,146-147).

Here the `[B, T_uc, C, H, W]` result exists once, on the device: the moments kernel writes each step's map straight
into slot `[:, k]` (a strided rows view — no copy at all), `stash()` places other per-step tensors (the scores) with one
du_accumulate_slot launch, and the whole batch leaves the GPU in ONE asynchronous copy into pinned host memory on a
side stream when the batch is done.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import ops


class UncertaintyMapAccumulator:
    """Preallocated `[B, T_uc, *map_shape]` device buffer with slot views."""

    def __init__(self, batch: int, num_slots: int, map_shape: Sequence[int], device, dtype: torch.dtype = torch.float32):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(f"UncertaintyMapAccumulator lives on a CUDA device (got {device}): the path has no CPU fallback")
        self.batch, self.num_slots, self.map_shape = int(batch), int(num_slots), tuple(int(s) for s in map_shape)
        self.buffer = torch.empty((self.batch, self.num_slots) + self.map_shape, device=device, dtype=dtype)
        self.cursor = 0
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._host: Optional[torch.Tensor] = None
        self._done: Optional[torch.cuda.Event] = None

    # ---------------------------------------------------------------------------------------------- slots
    def slot(self, k: int) -> torch.Tensor:
        if not 0 <= k < self.num_slots:
            raise IndexError(f"slot {k} outside [0, {self.num_slots})")
        return self.buffer[:, k]

    def accepts(self, shape, dtype) -> bool:
        """True when a kernel can write a map of this shape / dtype straight into the next slot."""
        return (self.cursor < self.num_slots and tuple(shape) == (self.batch,) + self.map_shape and dtype == self.buffer.dtype)

    def fit(self, map_shape: Sequence[int]) -> None:
        """Re-shape the (still empty) buffer for maps of another per-sample shape than the one it was built for
        (flip_threshold returns [B,1,H,W] maps for [B,C,H,W] samples)."""
        map_shape = tuple(int(s) for s in map_shape)
        if map_shape == self.map_shape:
            return
        if self.cursor != 0:
            raise ValueError(f"map of shape {map_shape} does not fit accumulator slots of shape {self.map_shape}")
        self.map_shape = map_shape
        self.buffer = torch.empty((self.batch, self.num_slots) + map_shape, device=self.buffer.device, dtype=self.buffer.dtype)

    def next_slot(self, shape=None, dtype=None) -> torch.Tensor:
        """The view the next in-window step writes its map into (advances the cursor)."""
        if self.cursor >= self.num_slots:
            raise IndexError(f"accumulator full: {self.num_slots} uncertainty steps already stored")
        view = self.slot(self.cursor)
        if shape is not None and tuple(shape) != tuple(view.shape):
            raise ValueError(f"map of shape {tuple(shape)} does not fit accumulator slots of shape {tuple(view.shape)}")
        if dtype is not None and dtype != view.dtype:
            raise ValueError(f"map dtype {dtype} does not match the accumulator's {view.dtype}")
        self.cursor += 1
        return view

    def stash(self, tensor: torch.Tensor, k: Optional[int] = None) -> torch.Tensor:
        """Copy (and convert) a per-step tensor into slot k (default: next) with one du_accumulate_slot launch."""
        if k is None:
            self.fit(tensor.shape[1:])
        view = self.next_slot(tensor.shape) if k is None else self.slot(k)
        ops.accumulate_slot(tensor, view)
        return view

    def reset(self) -> None:
        self.cursor = 0

    # ---------------------------------------------------------------------------------------------- leaving the GPU
    def filled(self) -> torch.Tensor:
        return self.buffer[:, :self.cursor]

    def to_host_async(self, out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.cuda.Event]:
        """One asynchronous D2H of the filled `[B, T_uc, ...]` block into pinned memory on a side stream.
        Returns (host tensor, event to wait on).  `out` may be a pinned slice of the final `[N_samples, T_uc, ...]` tensor,
        which removes the reference's torch.cat over batches as well."""
        src = self.filled()
        if out is None:
            if self._host is None or self._host.shape != src.shape or self._host.dtype != src.dtype:
                self._host = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            out = self._host
        elif tuple(out.shape) != tuple(src.shape):
            raise ValueError(f"host destination {tuple(out.shape)} != filled block {tuple(src.shape)}")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.buffer.device)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.buffer.device))
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            out.copy_(src, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        src.record_stream(self._copy_stream)
        self._done = done
        return out, done

    def to_host(self) -> torch.Tensor:
        out, done = self.to_host_async()
        done.synchronize()
        return out
