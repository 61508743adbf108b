"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic in diffusion_uncertainty_b200/distributed.py: sharding
arithmetic, the M-shard partial-moment exchange, the whole-batch z-norm statistics exchange and the batch-axis sum.
The CUDA kernels cannot run here, so the per-rank kernels are replaced by injected CPU checker functions that follow
the kernels' contracts (partial = (sum of squared deviations about the local mean, local mean); merge = pairwise Chan
update in rank order); the GPU parity of those kernels is tests/test_ops_gpu.py::test_moments_partial_merge."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import du_oracle as O


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


# ---- CPU stand-ins with the kernels' contracts ----------------------------------------------------------------------
def cpu_partial(scores, center, as_extra_sample):
    xs = [s.double() for s in scores]
    if as_extra_sample and center is not None:
        xs, center = xs + [center.double()], None
    st = torch.stack(xs, 0)
    if center is not None:
        return ((st - center.double().unsqueeze(0)) ** 2).sum(0).float(), st.mean(0).float()
    mean = st.mean(0)
    return ((st - mean.unsqueeze(0)) ** 2).sum(0).float(), mean.float()


def cpu_merge(means, m2s, counts, mode):
    if mode == "centered":
        return (sum(m.double() for m in m2s) / sum(counts)).float()
    mean, m2, na = means[0].double(), m2s[0].double(), float(counts[0])
    for r in range(1, len(m2s)):
        nr = float(counts[r])
        if nr == 0:
            continue
        delta = means[r].double() - mean
        tot = na + nr
        m2 = m2 + m2s[r].double() + delta * delta * (na * nr / tot)
        mean = mean + delta * (nr / tot)
        na = tot
    return (m2 / (na - 1.0)).float()


def cpu_combine(blocks):
    n, mean, m2 = 0.0, 0.0, 0.0
    for mu, _sd, cnt, q in blocks.double().tolist():
        if cnt == 0:
            continue
        delta = mu - mean
        tot = n + cnt
        m2 = m2 + q + delta * delta * n * cnt / tot
        mean = mean + delta * cnt / tot
        n = tot
    return torch.tensor([mean, (m2 / (n - 1)) ** 0.5, n, m2], dtype=torch.float32)


def worker(rank, world, port, M, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from diffusion_uncertainty_b200 import distributed as D
        g = torch.Generator().manual_seed(0)                      # same data on every rank, each takes its share
        eps = torch.randn(2, 4, 8, 8, generator=g)
        scores = [eps + 0.05 * torch.randn(2, 4, 8, 8, generator=g) for _ in range(M)]
        a, b = D.shard_range(M, rank, world)
        assert D.shard_samples(M, rank, world) == b - a
        sm = D.ShardedMoments(partial_fn=cpu_partial, merge_fn=cpu_merge)
        assert (sm.rank, sm.world) == (rank, world)
        for mode, want in (("var", O.variance_unbiased(scores)), ("var_with_center", O.variance_with_center(scores, eps)),
                           ("centered", O.centered_second_moment(scores, eps))):
            got = sm.reduce(scores[a:b], eps, mode)
            assert torch.allclose(got, want, rtol=1e-5, atol=1e-9, equal_nan=True), (mode, (got - want).abs().max())   # (var of ONE sample is NaN)
            got_m = sm.reduce(scores[a:b], eps, mode, total_M=M)      # counts known arithmetically: one collective only
            if mode == "centered":      # ONE all-reduce(sum) of the partial sums about the common centre, then a scale by 1/M
                assert torch.allclose(got_m, want, rtol=1e-5, atol=1e-9), mode
            else:
                assert torch.equal(got_m.view(torch.int32), got.view(torch.int32)), mode
        with pytest.raises(ValueError, match="shard_samples"):
            sm.reduce(scores[a:b], eps, "var", total_M=M + 2 * world + 1)
        # batch sharding: whole-batch z-norm statistics and the batch-axis sum
        u = torch.rand(6, 3, 8, 8, generator=g) ** 2
        (mine,) = D.shard_batch([u], rank, world)
        local = torch.tensor([mine.mean(), mine.std(), mine.numel(), ((mine - mine.mean()) ** 2).sum()], dtype=torch.float32)
        allst = D.allgather_znorm_stats(local, combine_fn=cpu_combine)
        assert torch.allclose(allst[:2], torch.tensor([u.mean(), u.std()]), rtol=1e-5)
        assert float(allst[2]) == u.numel()
        S = D.allreduce_batch_sum(mine.sum(0))
        assert torch.allclose(S, u.sum(0), rtol=1e-5, atol=1e-6)
        full = D.gather_maps_to_rank0(mine)
        if rank == 0:
            assert torch.equal(full, u)
        else:
            assert full is None
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("M", [16, 5, 1])     # M = 1: rank 1 owns no sample (a zero partial with count 0, not a deadlock)
def test_two_rank_exchanges_over_gloo(tmp_path, M):
    world = 2
    mp.spawn(worker, args=(world, free_port(), M, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_range_properties():
    from diffusion_uncertainty_b200.distributed import shard_batch, shard_range
    for n in (0, 1, 5, 16, 128, 131):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(128, r, 8) for r in range(8)][3] == (48, 64)   # the reference's X_T[gpu_idx*n : (gpu_idx+1)*n]
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
    x, y = torch.arange(10).view(10, 1), torch.arange(10)
    xs, ys = shard_batch([x, y], 1, 3)
    assert xs.flatten().tolist() == [4, 5, 6] and ys.tolist() == [4, 5, 6]
    with pytest.raises(ValueError):
        shard_batch([x, y[:3]], 0, 2)


def test_single_process_paths_need_no_process_group():
    from diffusion_uncertainty_b200 import distributed as D
    sm = D.ShardedMoments(partial_fn=cpu_partial, merge_fn=cpu_merge)
    g = torch.Generator().manual_seed(1)
    scores = [torch.randn(1, 4, 4, 4, generator=g) for _ in range(6)]
    assert torch.allclose(sm.reduce(scores, None, "var"), O.variance_unbiased(scores), rtol=1e-5, atol=1e-9)
    st = torch.tensor([0.5, 1.0, 10.0, 9.0])
    assert D.allgather_znorm_stats(st, combine_fn=cpu_combine) is st
    assert D.gather_maps_to_rank0(scores[0]) is scores[0]
