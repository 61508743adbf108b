"""Two-GPU NCCL tests of the exchanges in diffusion_uncertainty_b200/distributed.py with the real kernels (skipped on a
single-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_distributed_gpu.py -m gpu`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import du_oracle as O

pytestmark = pytest.mark.gpu


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    d = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=d)
    try:
        from diffusion_uncertainty_b200 import distributed as D
        from diffusion_uncertainty_b200 import ops
        g = torch.Generator().manual_seed(0)
        # ---- M sharding: SD-512 latent, M = 16 (BASELINE config 5)
        eps = torch.randn(1, 4, 64, 64, generator=g)
        scores = [eps + 0.05 * torch.randn(1, 4, 64, 64, generator=g) for _ in range(16)]
        a, b = D.shard_range(16, rank, world)
        sm = D.ShardedMoments()
        local = [s.to(d) for s in scores[a:b]]
        for mode, want in (("var", O.variance_unbiased(scores)), ("var_with_center", O.variance_with_center(scores, eps)),
                           ("centered", O.centered_second_moment(scores, eps))):
            got = sm.reduce(local, eps.to(d), mode).cpu()
            assert torch.allclose(got, want, rtol=1e-5, atol=1e-12), (mode, float((got - want).abs().max()))
            got_m = sm.reduce(local, eps.to(d), mode, total_M=16).cpu()
            if mode == "centered":     # ONE all-reduce(sum) of the partial sums about the common centre, then a scale by 1/M
                assert torch.allclose(got_m, want, rtol=1e-5, atol=1e-12), mode
            else:
                assert torch.equal(got_m, got), mode
        # every rank then runs the replicated rest of the step on identical maps
        u = sm.reduce(local, eps.to(d), "var_with_center")
        gathered = [torch.empty_like(u) for _ in range(world)]
        dist.all_gather(gathered, u)
        assert all(torch.equal(gathered[0], x) for x in gathered), "the merged map must be bit-identical on every rank"
        # ---- batch sharding: whole-batch z-norm statistics and the posterior's batch-axis sum
        umap = (torch.rand(8, 3, 16, 16, generator=g) ** 2)
        (mine,) = D.shard_batch([umap], rank, world)
        stats = D.allgather_znorm_stats(ops.znorm_stats(mine.to(d)))
        want = torch.tensor([umap.mean(), umap.std()])
        assert torch.allclose(stats[:2].cpu(), want, rtol=1e-5), (stats, want)
        S = D.allreduce_batch_sum(ops.batch_sum(mine.to(d)) if mine.shape[0] >= 1 else None)
        assert torch.allclose(S.cpu(), umap.sum(0), rtol=1e-5, atol=1e-6)
        full = D.gather_maps_to_rank0(mine.to(d))
        if rank == 0:
            assert torch.equal(full.cpu(), umap)
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_exchanges_over_nccl(tmp_path):
    world = 2
    mp.spawn(worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
