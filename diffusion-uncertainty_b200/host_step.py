"""The percentile-guided uncertainty step for callers whose tensors live in HOST memory.

This is the reference-facing entry for the end-to-end measurement (bench.py `e2e`): scores, eps and sample arrive in
pinned host buffers, x_{t-1} and the map are wanted back on the host.  A naive "copy everything in, run, copy everything
out" serialises three PCIe/NVLink-C2C transfers around a 60 µs kernel.  Here the batch is cut into image chunks and
pipelined over three streams:

    copy-in stream    H2D of chunk k+1 (M score tensors, sample)
    compute stream    du_fused_uncertainty_step on chunk k           (per-image work: chunks are independent)
    copy-out stream   D2H of chunk k-1 (x_{t-1}, map)

so the step costs about max(H2D, D2H) instead of H2D + kernel + D2H.  Consecutive calls pipeline too: the next call's H2D
starts as soon as this call's last kernel has consumed the staging buffers (it does not wait for this call's D2H — PCIe is
full duplex), and its first kernel waits for this call's last D2H before it overwrites the output staging buffers.  The
one thing that couples images — the
reference's posterior sum over the batch axis (`pred_epsilon.sum(dim=0)`, uncertainty_guidance.py:119) — is handled
by sending eps for the whole batch first and reducing it (du_batch_sum) before the first chunk is updated.
Device staging buffers are allocated once and reused by every call.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops


class HostStreamedUncertaintyStep:
    def __init__(self, batch: int, shape: Sequence[int], M: int, device, score_dtype: torch.dtype = torch.float32,
                 chunks: int = 8):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"device {self.device}: the uncertainty path has no CPU fallback")
        self.B, self.shape, self.M = int(batch), tuple(int(s) for s in shape), int(M)
        self.chunks = max(1, min(int(chunks), self.B))
        full = (self.B,) + self.shape
        d = self.device
        self.d_scores = [torch.empty(full, device=d, dtype=score_dtype) for _ in range(self.M)]
        self.d_eps = torch.empty(full, device=d, dtype=score_dtype)
        self.d_sample = torch.empty(full, device=d, dtype=torch.float32)
        self.d_prev = torch.empty(full, device=d, dtype=torch.float32)
        self.d_map = torch.empty(full, device=d, dtype=torch.float32)
        self.d_S = torch.empty(self.shape, device=d, dtype=torch.float32)
        self.s_in, self.s_out = torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)
        self.bounds = [(k * self.B // self.chunks, (k + 1) * self.B // self.chunks) for k in range(self.chunks)]
        self._inputs_free: Optional[torch.cuda.Event] = None    # previous call's last kernel is done with d_scores / d_eps / d_sample
        self._outputs_free: Optional[torch.cuda.Event] = None   # previous call's last D2H is done with d_prev / d_map

    def __call__(self, h_scores: List[torch.Tensor], h_eps: torch.Tensor, h_sample: torch.Tensor, q: float, coeffs,
                 alpha_hat_t: float, out_prev: torch.Tensor, out_map: torch.Tensor, batch_sum: bool = True,
                 map_slot: Optional[torch.Tensor] = None):
        """h_*: pinned host tensors [B, ...]; out_prev / out_map: pinned host tensors that receive x_{t-1} and the map.
        map_slot: optional device slot view `[B, ...]` of the accumulation buffer that also receives the map (F8).
        Returns the event that marks the last D2H (call `.synchronize()` before reading the outputs)."""
        if len(h_scores) != self.M:
            raise ValueError(f"expected {self.M} score tensors, got {len(h_scores)}")
        main = torch.cuda.current_stream(self.device)
        if self._inputs_free is None:
            self.s_in.wait_stream(main)
            self.s_out.wait_stream(main)
        else:
            self.s_in.wait_event(self._inputs_free)
        u_dst = map_slot if map_slot is not None else self.d_map
        with torch.cuda.stream(self.s_in):
            self.d_eps.copy_(h_eps, non_blocking=True)           # whole batch first: the batch-axis sum needs all of it
            eps_ready = torch.cuda.Event()
            eps_ready.record(self.s_in)
        S = None
        if batch_sum and self.B > 1:
            main.wait_event(eps_ready)
            S = ops.batch_sum(self.d_eps, out=self.d_S)
        in_ready = []
        with torch.cuda.stream(self.s_in):
            for a, b in self.bounds:
                for m in range(self.M):
                    self.d_scores[m][a:b].copy_(h_scores[m][a:b], non_blocking=True)
                self.d_sample[a:b].copy_(h_sample[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_in)
                in_ready.append(ev)
        last = None
        for k, (a, b) in enumerate(self.bounds):
            main.wait_event(in_ready[k])
            if k == 0:
                main.wait_event(eps_ready)
                if self._outputs_free is not None:
                    main.wait_event(self._outputs_free)
            ops.uncertainty_step([s[a:b] for s in self.d_scores], self.d_eps[a:b], self.d_sample[a:b], q, coeffs, alpha_hat_t,
                                 precomputed_sum=S, batch_sum=batch_sum, map_out=u_dst[a:b], prev_out=self.d_prev[a:b])
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                out_prev[a:b].copy_(self.d_prev[a:b], non_blocking=True)
                out_map[a:b].copy_(u_dst[a:b], non_blocking=True)
                last = torch.cuda.Event()
                last.record(self.s_out)
        self._inputs_free = torch.cuda.Event()
        self._inputs_free.record(main)
        self._outputs_free = last
        return last

    def synchronize(self):
        """Wait for every transfer of every call so far (the staging buffers and the host outputs are then quiescent)."""
        if self._outputs_free is not None:
            self._outputs_free.synchronize()
