#!/usr/bin/env python
"""Every kernel of the library once, at the BASELINE shapes — the driver for the per-kernel evidence (north_star: "every kernel's
achieved HBM GB/s against B200 peak is evidenced by a committed ncu capture").

  python tools/kernel_zoo.py            CUDA-event timing of each kernel (30 launches back to back), one JSON line per kernel with
                                        its algorithmic bytes and the fraction of the measured HBM peak
  ncu --set full -k regex:... python tools/kernel_zoo.py --once     one launch of each (after one warm-up), for the profiler

Shapes: ImageNet-128 b128 fp32 M=5 (6.29 M elements per tensor) unless the row says otherwise.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from diffusion_uncertainty_b200 import ops  # noqa: E402

ONCE = "--once" in sys.argv
dev = torch.device("cuda:0")
peak = bench.peaks()[0]


def timeit(fn, reps=30):
    fn(); fn()
    torch.cuda.synchronize()
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return None
    # `reps` launches captured in ONE CUDA graph: the wrappers' Python (5-40 us per call) is not what is being measured
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def row(kernel, what, us, alg_bytes, note=""):
    if us is None:
        return
    print(json.dumps({"kernel": kernel, "call": what, "us": round(us, 2), "algorithmic_MB": round(alg_bytes / 1e6, 1),
                      "GBps": round(alg_bytes / us / 1e3, 1), "frac_of_measured_hbm_peak": round(alg_bytes / us / 1e3 / peak, 3),
                      "note": note}), flush=True)


def main():
    sb = bench.StepBench(ops, "imagenet128_adm_b128_m5", "fp32", 128, dev, 1234)
    n = sb.n_el
    M = sb.M
    sc = sb.sc
    u = torch.empty_like(sb.eps)
    # F1
    row("moments_kernel", "du_moments(var_with_center)", timeit(lambda: ops.moments(sb.scores, center=sb.eps, mode="var_with_center", out=u)),
        (M + 1) * 4 * n + 4 * n)
    row("moments_kernel", "du_moments(centered) into a slot of the accumulation buffer",
        timeit(lambda: ops.moments(sb.scores, center=sb.eps, mode="centered", out=sb.maps[:, 3])), (M + 1) * 4 * n + 4 * n)
    # F2a
    row("quantile_rows_kernel", "du_quantile_threshold(q=0.9)", timeit(lambda: ops.quantile_threshold(u, 0.9)), 4 * n,
        "u (25 MB) is L2-resident after the moments kernel; 4 B/element algorithmic, re-read per radix pass from L2")
    thr = ops.quantile_threshold(u, 0.9)
    row("rows_kernel<ThrMaskF>", "du_threshold_mask", timeit(lambda: ops.threshold_mask(u, thr)), 8 * n)
    # F5 + F3
    ops.batch_sum(sb.eps, out=sb.S)
    row("batch_sum_rows_kernel", "du_batch_sum", timeit(lambda: ops.batch_sum(sb.eps, out=sb.S)), 4 * n)
    row("rows_kernel<GuidedF>", "du_guided_step(posterior, thr)", timeit(
        lambda: ops.guided_step(sb.eps, sb.sample, sb.coeffs, guidance="posterior", u=u, thr=thr, aux=sb.S, aux_broadcast=True,
                                post_M=float(M), inv_alpha_hat=1.0 / sc["alpha_hat"], want_eps=False)), 16 * n,
        "reads eps, sample, u; writes x_(t-1)")
    row("rows_kernel<DdimF>", "du_ddim_step(prev + x0)", timeit(lambda: ops.ddim_step(sb.eps, sb.sample, sb.coeffs)), 16 * n)
    # F7
    row("rows_kernel<PerturbF>", "du_perturb", timeit(lambda: ops.perturb(sb.sample, sb.eps, 0.99, 0.1)), 12 * n)
    rng = ops.DeviceRng(dev, 1234, 0, deferred=True)     # (a graph capture cannot advance torch's generator from the host)
    row("perturb_randn_kernel", "du_perturb_randn (noise drawn in the kernel)", timeit(lambda: rng.perturb(sb.sample, 0.99, 0.1)), 8 * n,
        "Philox4x32-10 + Box-Muller per element, fixed by bit-parity with torch's generator")
    # F8 / N4
    row("rows_kernel<CopyF>", "du_accumulate_slot", timeit(lambda: ops.accumulate_slot(u, sb.maps[:, 5])), 8 * n)
    row("rows_kernel<ImageU8F>", "du_image_uint8", timeit(lambda: ops.image_uint8(sb.sample)), 5 * n)
    # F2c
    row("znorm_partial_kernel", "du_znorm_stats", timeit(lambda: ops.znorm_stats(u)), 4 * n)
    stats = ops.znorm_stats(u)
    row("rows_kernel<ZnormF>", "du_znorm_weights", timeit(lambda: ops.znorm_weights(u, stats, mode="max", thr=1.0)), 12 * n)
    # N2 / N3
    maps = torch.rand(256, 3 * 64 * 64, device=dev)
    row("column_kth_kernel", "du_column_kth (256 samples x 12288 pixels)", timeit(lambda: ops.column_kth(maps, 230)), maps.numel() * 4,
        "offline tool: 32 counting passes over the slice")
    row("row_sum_kernel", "du_row_sum", timeit(lambda: ops.row_sum(sb.maps[:, 0])), 4 * n)
    # the fused step, both kernels
    row("fused_pred_kernel", "du_fused_uncertainty_step (ImageNet-128 b128)", timeit(lambda: sb.kernel_only(0)), sb.alg_bytes(),
        "x_(t-1) into one fixed buffer here (bench.py rotates a ring of 8)")
    s64 = bench.StepBench(ops, "imagenet64_adm_b128_m5", "fp32", 128, dev, 1234)
    row("fused_step_kernel", "du_fused_uncertainty_step (ImageNet-64 b128)", timeit(lambda: s64.kernel_only(0)), s64.alg_bytes(),
        "56.6 MB: L2-resident working set")
    ssd = bench.StepBench(ops, "sd512_latent_b1_m16", "fp32", 1, dev, 1234)
    row("fused_step_kernel", "du_fused_uncertainty_step (SD latent 1x4x64x64, M=16)", timeit(lambda: ssd.kernel_only(0)), ssd.alg_bytes(),
        "1.3 MB: latency-bound")


if __name__ == "__main__":
    main()
