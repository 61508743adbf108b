"""Multi-GPU plumbing of the uncertainty path: one process per GPU, `torch.distributed` (NCCL over NVLink 5 / NVSwitch on
the B200 box) for the few places where the path has a real exchange step.

  * BATCH sharding (ImageNet-64/128, U-ViT): moments are per element, quantiles and masks per image, the DDIM update
    elementwise, the accumulation per sample — contiguous image ranges per rank and NO collective, exactly like the
    reference's `mp.spawn` slicing (scripts/generate_dataset_score_uncertainty_imagenet.py:51, 137-144).  Two reference
    quirks couple images and therefore need one tiny exchange when the batch is sharded:
        - the whole-batch z-normalisation of the in-scheduler threshold variants (F2c): all-gather of the per-rank
          (mean, std, count, M2) block, Chan-combined by du_znorm_stats_combine;
        - the posterior score's sum over the batch axis (F5, `pred_epsilon.sum(dim=0)`): all-reduce(sum) of one [C,H,W] row.
  * M sharding (Stable Diffusion, B = 1, M = 16): every rank draws its own M/R perturbations and reduces them to
    per-element partial moments (count, mean, M2) with du_moments(DU_MOM_PARTIAL_M2); one all-gather of the packed
    (mean, M2) pair and du_moments_merge (pairwise Chan update in rank order, deterministic) give the exact M-sample
    variance on every rank; the latent-sized rest of the step is replicated.

The collective calls are `torch.distributed` ones so the same code runs over NCCL on GPUs and over gloo in the CPU
tests of the host logic (tests/test_distributed_cpu.py), where the kernels are replaced by injected checker functions.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------ sharding arithmetic
def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `n_items` owned by `rank`; the first n_items % world ranks get one extra item."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(n_items), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    """The rank's image range of every tensor (views, no copy)."""
    n = tensors[0].shape[0]
    for t in tensors:
        if t.shape[0] != n:
            raise ValueError("shard_batch: tensors disagree on the batch size")
    a, b = shard_range(n, rank, world)
    return [t[a:b] for t in tensors]


def shard_samples(M: int, rank: int, world: int) -> int:
    """How many of the M perturbed forwards this rank runs."""
    a, b = shard_range(M, rank, world)
    return b - a


def _world(group) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


# ------------------------------------------------------------------------------------------------ M sharding
class ShardedMoments:
    """Exact moments over M samples that are spread over the ranks of `group`.

    partial_fn(scores, center, as_extra_sample) -> (m2, mean): per-element partial moments of the LOCAL samples
        (default: du_moments in DU_MOM_PARTIAL_M2 mode);
    merge_fn(means, m2s, counts, mode) -> map: Chan merge of the gathered partials (default: du_moments_merge).
    Both are injectable so that the gloo tests can run the plumbing on CPU tensors with the oracle as the checker."""

    def __init__(self, group=None, partial_fn: Optional[Callable] = None, merge_fn: Optional[Callable] = None):
        self.group = group
        self.rank, self.world = _world(group)
        self._partial = partial_fn or self._kernel_partial
        self._merge = merge_fn or self._kernel_merge

    @staticmethod
    def _kernel_partial(scores, center, as_extra_sample):
        from . import ops
        if as_extra_sample and center is not None:
            scores, center = list(scores) + [center], None
        m2, mean = ops.moments(scores, center=center, mode="partial", return_mean=True)
        return m2, mean

    @staticmethod
    def _scale(x, a):
        if x.is_cuda:
            from . import ops
            return ops.scale(x, a)
        return x * a          # (gloo tests of the host logic run on CPU tensors)

    @staticmethod
    def _kernel_merge(means, m2s, counts, mode):
        from . import ops
        return ops.moments_merge(means, m2s, counts, mode=mode)

    def reduce(self, local_scores: Sequence[torch.Tensor], center: Optional[torch.Tensor], mode: str,
               total_M: Optional[int] = None, empty_like: Optional[torch.Tensor] = None) -> torch.Tensor:
        """mode: 'var' (F1b), 'var_with_center' (F1c: the centre counts as one more sample, contributed by rank 0 only)
        or 'centered' (F1a: sum of squared deviations about the common centre / total M).
        total_M: the number of samples over all ranks when they were dealt out with `shard_samples` — the per-rank counts
        are then known arithmetically, the step is ONE collective (the packed partials) and never reads a device value
        on the host; without it the counts are exchanged too (one more small all-gather and a host read)."""
        if mode not in ("var", "var_with_center", "centered"):
            raise ValueError(f"ShardedMoments: unsupported mode {mode!r}")
        n_local = len(local_scores)
        # everything that can raise is checked BEFORE the collective: a rank that raised after its peers had entered the
        # all-gather would leave them waiting forever
        counts = None
        if total_M is not None:
            counts = [shard_samples(total_M, r, self.world) + (1 if (mode == "var_with_center" and r == 0) else 0)
                      for r in range(self.world)]
            if shard_samples(total_M, self.rank, self.world) != n_local:
                raise ValueError(f"ShardedMoments: rank {self.rank} holds {n_local} samples, shard_samples({total_M}) says "
                                 f"{shard_samples(total_M, self.rank, self.world)}")
        if n_local == 0:
            # a rank without samples (M < world): a zero partial with count 0 is the identity of the Chan merge.  With
            # 'var_with_center' rank 0 still contributes the centre as its one sample (mean = centre, M2 = 0).
            like = center if center is not None else empty_like
            if like is None:
                raise ValueError("ShardedMoments: a rank without samples needs `center` or `empty_like` for the shape")
            m2 = torch.zeros(like.shape, device=like.device, dtype=torch.float32)
            if mode == "var_with_center" and self.rank == 0:
                mean, count = center.to(torch.float32), 1
            else:
                mean, count = torch.zeros_like(m2), 0
        elif mode == "centered":
            m2, mean = self._partial(local_scores, center, False)          # about the centre; mean unused
            count = n_local
        elif mode == "var_with_center" and self.rank == 0:
            m2, mean = self._partial(local_scores, center, True)
            count = n_local + 1
        else:
            m2, mean = self._partial(local_scores, None, False)
            count = n_local
        if self.world == 1:
            return self._merge([mean], [m2], [count], "centered" if mode == "centered" else "var")
        if mode == "centered" and counts is not None:
            # F1a needs no means: the partial sums of squared deviations about the COMMON centre add up — ONE all-reduce(sum) of N
            # floats, then one scale by 1 / M.  (NCCL's sum order is fixed for a given communicator, so every
            # rank holds the same bits.)
            m2 = m2.float().contiguous()
            dist.all_reduce(m2, op=dist.ReduceOp.SUM, group=self.group)
            return self._scale(m2, 1.0 / float(sum(counts)))
        packed = torch.stack([mean.float(), m2.float()], dim=0).contiguous()          # [2, ...]
        gathered = torch.empty((self.world,) + tuple(packed.shape), device=packed.device, dtype=packed.dtype)
        dist.all_gather_into_tensor(gathered.view(self.world * packed.shape[0], *packed.shape[1:]), packed, group=self.group)
        if counts is None:
            counts_t = torch.tensor([count], device=packed.device, dtype=torch.int64)
            all_counts = torch.empty(self.world, device=packed.device, dtype=torch.int64)
            dist.all_gather_into_tensor(all_counts, counts_t, group=self.group)
            counts = [int(c) for c in all_counts.tolist()]
        means = [gathered[r, 0] for r in range(self.world)]
        m2s = [gathered[r, 1] for r in range(self.world)]
        return self._merge(means, m2s, counts, "centered" if mode == "centered" else "var")


# ------------------------------------------------------------------------------------------------ batch-sharding exchanges
def allgather_znorm_stats(stats: torch.Tensor, group=None, combine_fn: Optional[Callable] = None) -> torch.Tensor:
    """F2c under batch sharding: per-rank [mean, std, count, M2] blocks -> whole-batch block on every rank."""
    rank, world = _world(group)
    if combine_fn is None:
        from . import ops
        combine_fn = ops.znorm_stats_combine
    if world == 1:
        return stats
    blocks = torch.empty(world * 4, device=stats.device, dtype=stats.dtype)
    dist.all_gather_into_tensor(blocks, stats.reshape(4).contiguous(), group=group)
    return combine_fn(blocks.view(world, 4))


def allreduce_batch_sum(S: torch.Tensor, group=None) -> torch.Tensor:
    """F5 under batch sharding: the reference's sum over the batch axis spans all ranks' images."""
    rank, world = _world(group)
    if world > 1:
        dist.all_reduce(S, op=dist.ReduceOp.SUM, group=group)
    return S


def gather_maps_to_rank0(local_maps: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """Optional epilogue: the reference writes one `uncertainty_<type>_<gpu>.pth` per rank
    (scripts/generate_dataset_score_uncertainty_imagenet.py:92); this returns the concatenation on rank 0 instead."""
    rank, world = _world(group)
    if world == 1:
        return local_maps
    sizes = torch.empty(world, device=local_maps.device, dtype=torch.int64)
    dist.all_gather_into_tensor(sizes, torch.tensor([local_maps.shape[0]], device=local_maps.device, dtype=torch.int64), group=group)
    sizes = [int(s) for s in sizes.tolist()]
    out = [torch.empty((s,) + tuple(local_maps.shape[1:]), device=local_maps.device, dtype=local_maps.dtype) for s in sizes] if rank == 0 else None
    dist.gather(local_maps.contiguous(), out, dst=0, group=group)
    return torch.cat(out, dim=0) if rank == 0 else None
