"""Parity on the configurations the numbers are quoted for: every `bench.WORKLOADS` entry at its FULL batch —
the headline (128, 3x128x128, M=5, q=0.9) in fp32 and fp16 included — through the very objects bench.py times: the prepared
`FusedStep`, `launch_with_batch_sum` (du_batch_sum + the step as its programmatic dependent), replayed from a CUDA graph, x_{t-1}
written into the ring of output buffers.  Checked against (a) the CPU oracle's whole chain and (b) the reference's eager torch
expressions on the GPU (bench.StepBench.parity, the check bench.py runs before it times anything).

Bars: map within 1e-5 relative (fp32 arithmetic on the same — possibly 16-bit — inputs); thresholds bit-identical to
torch.quantile of the kernel's map; masks identical to the oracle's wherever the oracle's map is not within 1e-5 of its
threshold; x_{t-1} within 1e-5 of max(|x|, 0.4) on the pixels whose masks agree (the update `x - sqrt(1-abar) eps` cancels, so the
bound is relative to the operand scale: |x|, |eps| <= ~5 gives 6 roundings x 2^-24 x 5 < 2e-6; the fused kernel's two approximate
reciprocals add <= 2 ulp each to the guided score, which enters x_{t-1} with weight sqrt(1-abar) <= 1).
"""
import numpy as np
import pytest
import torch

import bench
from oracle import du_oracle as O

pytestmark = pytest.mark.gpu

CASES = [(w, "fp32") for w in bench.WORKLOADS] + [("imagenet128_adm_b128_m5", "fp16"), ("imagenet128_adm_b128_m5", "bf16"),
                                                  ("imagenet64_adm_b128_m5", "fp16")]


@pytest.fixture(scope="module")
def ops():
    from diffusion_uncertainty_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("workload,dtype", CASES)
def test_bench_workload_full_batch(ops, workload, dtype):
    dev = torch.device("cuda:0")
    B = bench.WORKLOADS[workload][0]
    sb = bench.StepBench(ops, workload, dtype, B, dev, 1234)
    assert sb.fused, "every BASELINE shape takes the single-launch path"
    sb.warm(3)
    par = sb.parity()                       # (b): eager torch on the GPU; raises on failure
    assert par["thr_bit_exact"] and par["map_max_rel_err_vs_exact"] < 1e-5

    # the timed object: K steps in ONE CUDA graph; the replay must reproduce the eager launches bit for bit
    sb.step(0)
    torch.cuda.synchronize()
    prev_eager, map_eager = sb.prevs[0].clone(), sb.maps[:, 0].clone()
    thr_eager = sb.plan.res["thr"].clone()
    g = sb.capture(sb.step, 3)
    sb.prevs[0].zero_(); sb.maps.zero_()
    g.replay()
    torch.cuda.synchronize()
    bits = lambda t: t.view(torch.int32)      # noqa: E731  (u = 0 gives NaN updates, as in the reference: compare bit patterns)
    assert torch.equal(bits(sb.prevs[0]), bits(prev_eager)) and torch.equal(bits(sb.maps[:, 0]), bits(map_eager))
    assert torch.equal(bits(sb.plan.res["thr"]), bits(thr_eager))
    assert torch.equal(bits(sb.prevs[2]), bits(prev_eager)) and torch.equal(bits(sb.maps[:, 2]), bits(map_eager))
    assert ops.fused_last_kernel() == sb.kernel
    if workload == "imagenet128_adm_b128_m5":
        assert sb.kernel == "fused_pred_kernel"

    # (a): the CPU oracle's chain on the same inputs (16-bit scores upcast exactly)
    h_eps, h_scores, h_sample = sb.h
    ef, sf = h_eps.float(), [s.float() for s in h_scores]
    M, q = sb.M, sb.q
    ac = torch.cumprod(1 - O.make_betas(), 0)
    c = O.DDIMCoeffs(ac, torch.tensor(1.0), bench.TIMESTEP, bench.TIMESTEP - bench.STEP_RATIO, 0.0)
    u_o, mask_o, _, prev_o, _ = O.uncertainty_step_posterior(sf, ef, h_sample, q, M, ac[bench.TIMESTEP], c, batch_sum=sb.batch_sum)
    u_k = map_eager.cpu()
    # the map against the EXACT variance of the fp32 inputs (fp64): 1e-5 relative everywhere; against the oracle's fp32 torch.var
    # the same plus the oracle's own deviation from the exact value (small-variance pixels: see bench.StepBench.parity)
    u64 = torch.var(torch.stack([s.double() for s in sf] + [ef.double()], dim=0), dim=0)
    assert float(((u_k.double() - u64).abs() / u64.clamp_min(1e-300)).max()) < 1e-5
    ref_err = float(((u_o.double() - u64).abs() / u64.clamp_min(1e-300)).max())
    assert float(((u_k - u_o).abs() / u_o.abs().clamp_min(1e-30)).max()) < 1e-5 + 1.01 * ref_err
    del u64
    thr_k = thr_eager.cpu()
    assert np.array_equal(thr_k.numpy(), torch.quantile(u_k.flatten(1), q, dim=1).numpy())
    mask_k = (u_k > thr_k.view(-1, 1, 1, 1)).float()
    thr_o = torch.quantile(u_o.flatten(1), q, dim=1).view(-1, 1, 1, 1)
    safe = (u_o - thr_o).abs() > 1e-5 * thr_o.abs()
    assert torch.equal(mask_k[safe], mask_o[safe])
    agree = (mask_k == mask_o) & torch.isfinite(prev_o)
    err = (prev_eager.cpu() - prev_o).abs()[agree]
    assert bool((err <= 1e-5 * prev_o.abs().clamp_min(0.4)[agree]).all()), float(err.max())
    assert float(agree.float().mean()) > 0.9995


def test_fused_step_on_another_device_keeps_the_callers_device(ops):
    """an op on a cuda:1 tensor must neither move torch.cuda.current_device() nor launch on a stale device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    x = torch.randn(4, 3, 32, 32, device="cuda:1")
    n = torch.randn(4, 3, 32, 32, device="cuda:1")
    out = ops.perturb(x, n, 0.5, 0.25)
    assert torch.cuda.current_device() == 0 and out.device.index == 1
    assert torch.equal(out, 0.5 * x + 0.25 * n)
    y = torch.randn(4, 3, 32, 32, device="cuda:0")
    assert torch.equal(ops.perturb(y, y, 1.0, 1.0), y + y)
    torch.cuda.set_device(1)            # the caller changes device behind the library's back
    assert torch.equal(ops.perturb(y, y, 2.0, 0.0), 2.0 * y + 0.0 * y) and torch.cuda.current_device() == 1
    torch.cuda.set_device(0)
