#!/usr/bin/env python
"""bench.py — the uncertainty-step benchmark (BASELINE.json metric: uncertainty-step Mpix/s + % HBM peak).

A "step" is one pass of the hot path over one batch of synthetic score tensors:
  F1c variance over the M perturbed predictions + eps  ->  F2a per-image percentile threshold + mask
  ->  F5 posterior score blend  ->  F3 DDIM x_{t-1}  (+F8: the map is written into its accumulation slot)
Workload (config.workload): ImageNet-128 ADM shapes, batch 128 per GPU, M=5, fp32, q=0.9 — the configuration the
metric is quoted on (226.5 MB of algorithmic traffic per step, larger than the 126 MB L2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--dtype fp32|fp16|bf16]

N>1 is launched by torchrun (one rank per GPU).  Default: the workload's batch is SHARDED over the ranks (strong scaling, BASELINE
configs[2]; every shard a batch of its own, no data-path collective: moments are per element, quantiles per image), with the weak
number (the whole batch on every GPU) as a sub-record; --scaling weak swaps the two, --allreduce-batch-sum makes the ranks share one
posterior batch-axis sum (an NCCL all-reduce inside every step).  The default line also carries the ImageNet-128 sampling-loop img/s
(BASELINE metric iii) as the `sampling_loop` sub-record.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, C, H, W, M, q)
    "imagenet128_adm_b128_m5": (128, 3, 128, 128, 5, 0.9),
    "imagenet64_adm_b128_m5": (128, 3, 64, 64, 5, 0.9),
    "cifar10_ddpm_b16_m5": (16, 3, 32, 32, 5, 0.95),
    "uvit256_latent_b128_m5": (128, 4, 32, 32, 5, 0.9),
    "sd512_latent_b1_m16": (1, 4, 64, 64, 16, 0.9),
}
DTYPES = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}
T_UC = 10            # accumulation slots (num_steps_uc of the README command)
TIMESTEP = 180       # first uncertainty timestep of `--start-step-uc 40` at 50 steps
STEP_RATIO = 20


L2_BYTES = 126e6
IN_RING_MAX = 16


def input_ring(alg_bytes):
    """Timing rule: a step must not find its inputs in L2.  A working set above the 126 MB L2 needs nothing; a smaller one (the shards
    of the strong split, the small workloads) rotates over this many resident copies of the input set, so that at least 2 x L2 of
    other traffic passes between two uses of a copy (capped: below ~16 MB per step the numbers are L2-resident and latency-bound,
    and the line says so)."""
    if alg_bytes > L2_BYTES:
        return 1
    return min(IN_RING_MAX, int(math.ceil(2 * L2_BYTES / alg_bytes)) + 1)


def l2_note(alg_bytes):
    R, mb = input_ring(alg_bytes), alg_bytes / 1e6
    if R == 1:
        return "no flush: per-step working set %.1f MB per GPU > 126 MB L2" % mb
    if (R - 1) * alg_bytes >= 2 * L2_BYTES:
        return ("no flush: the steps rotate over %d resident copies of the input set (%.1f MB per step per GPU; %.0f MB pass between two "
                "uses of a copy, > 2 x 126 MB L2) and over rings of output buffers" % (R, mb, (R - 1) * mb))
    return "L2-resident, latency-bound: %.1f MB per step per GPU, %d input copies rotate (%.0f MB in all, < L2)" % (mb, R, R * mb)


def ulp_distance(a, b):
    """largest distance, in units of the last place, between two fp32 tensors of one shape (inf if a NaN faces a number)"""
    a, b = a.detach().cpu().float().flatten(), b.detach().cpu().float().flatten()
    nan = torch.isnan(a)
    if a.shape != b.shape or not torch.equal(nan, torch.isnan(b)):
        return float("inf")
    if bool(nan.all()):
        return 0

    def line(x):       # sign-magnitude bit patterns -> one monotone integer line (-0 and +0 coincide)
        i = x[~nan].contiguous().view(torch.int32).long()
        return torch.where(i < 0, -(i & 0x7FFFFFFF), i)
    return int((line(a) - line(b)).abs().max())


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel, workload=None, images_per_gpu=None, dtype=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed `ncu --set full` summary of this
    round (profiles/ncu_traffic.json; written by tools/ncu_summary.py from the capture).  None when there is no capture of this
    kernel at this workload, per-GPU batch and score dtype."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            rec = json.load(f).get(kernel, {})
        for key, have in (("workload", workload), ("images_per_gpu", images_per_gpu), ("score_dtype", dtype)):
            if have is not None and rec.get(key) is not None and rec[key] != have:
                return None
        return rec.get("dram_bytes_per_launch")
    except Exception:
        return None


def algorithmic_bytes_per_element(M, score_bytes, sample_bytes=4):
    """(M+1)*s_in (scores + eps) + s_x (sample) + 4 (u -> map slot) + s_x (x_{t-1})."""
    return (M + 1) * score_bytes + sample_bytes + 4 + sample_bytes


def ddim_scalars(t=TIMESTEP, ratio=STEP_RATIO):
    """Scalars of one DDIM step with the reference's fp32 expressions (linear betas 1e-4..0.02,
    SU/scheduling_ddim_uncertainty_zigzag_centered.py:462-468,507)."""
    betas = torch.linspace(1e-4, 0.02, 1000, dtype=torch.float32)
    ac = torch.cumprod(1.0 - betas, dim=0)
    a_t, a_prev = ac[t], ac[t - ratio] if t - ratio >= 0 else torch.tensor(1.0)
    return dict(sqrt_alpha_t=(a_t ** 0.5).item(), sqrt_beta_t=((1 - a_t) ** 0.5).item(),
                sqrt_alpha_prev=(a_prev ** 0.5).item(), dir_coef=((1 - a_prev) ** 0.5).item(), alpha_hat=a_t.item())


def synth_host(B, C, H, W, M, dtype, seed, pin):
    g = torch.Generator().manual_seed(seed)
    eps = torch.randn(B, C, H, W, generator=g)
    scores = [(eps + 0.05 * torch.randn(B, C, H, W, generator=g)).to(dtype) for _ in range(M)]
    sample = torch.randn(B, C, H, W, generator=g)
    eps = eps.to(dtype)
    if pin:
        eps, sample, scores = eps.pin_memory(), sample.pin_memory(), [s.pin_memory() for s in scores]
    return eps, scores, sample


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self, window=None):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm),
                "window": window or "sustained replays of the timed graph around the K timed steps (the K-step region is shorter "
                                    "than one nvidia-smi sampling period)"}


# ------------------------------------------------------------------------------------------- CPU arms
def cpu_step_fn(B, C, H, W, M, q):
    """The reference's torch CPU expressions for the step (oracle/du_oracle.py restates them line for line)."""
    from oracle import du_oracle as O
    sc = ddim_scalars()
    ac = torch.cumprod(1.0 - O.make_betas(), dim=0)
    c = O.DDIMCoeffs(ac, torch.tensor(1.0), TIMESTEP, TIMESTEP - STEP_RATIO, 0.0)
    a_hat = ac[TIMESTEP]

    def step(eps, scores, sample):
        return O.uncertainty_step_posterior(scores, eps, sample, q, M, a_hat, c, clip_sample=True, batch_sum=True)
    del sc
    return step


def time_cpu(workload, sample_images, budget_s, min_reps=2):
    B, C, H, W, M, q = WORKLOADS[workload]
    b = min(B, sample_images)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    eps, scores, sample = synth_host(b, C, H, W, M, torch.float32, 1234, pin=False)
    step = cpu_step_fn(b, C, H, W, M, q)
    step(eps, scores, sample)  # warm-up
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < min_reps or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter()
        step(eps, scores, sample)
        times.append(time.perf_counter() - t0)
    best = min(times)
    return {"value": b * H * W / best / 1e6, "unit": "Mpix/s", "cores": cores, "kind": "port",
            "sample": f"{b} of {B} images of {workload}, fp32, best of {len(times)} steps ({best * 1e3:.1f} ms/step)"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (torch eager expressions, restated in
    oracle/du_oracle.py and pinned bit-exact to the reference by tests/golden) on the box's host cores, on the SAME
    configuration as the GPU arm: the whole batch of the workload per step.  (The reference cannot be pip-installed here — its
    build backend, hatchling, is not in the wheelhouse — and the arithmetic of this path is inline in functions that also call
    the score model, so the port is what is timed)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B, C, H, W, M, q = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    eps, scores, sample = synth_host(B, C, H, W, M, torch.float32, 1234, pin=False)
    step = cpu_step_fn(B, C, H, W, M, q)
    for _ in range(max(1, min(args.warmup, 3))):
        step(eps, scores, sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(eps, scores, sample)
    dt = time.perf_counter() - t0
    val = B * H * W * args.steps / dt / 1e6
    line = {"impl": "reference", "metric": "uncertainty_step_throughput", "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": step_config(args.workload, B, world, True, "f32", "strong" if world > 1 else "weak"),
            "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "port",
                             "sample": f"all {B} images of {args.workload} per step, {args.steps} steps, torch {torch.__version__} CPU, "
                                       f"{cores} threads"},
            "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def step_config(workload, B_total, world, batch_sum, dtype_tag, scaling="weak", allreduce=False):
    """`config` of a step line — the same keys and values in both arms (--impl reference prints the GPU arm's configuration: it is a
    function of the workload, the rank count and the scaling mode only; what is specific to one run sits in the line's `run`)."""
    _, C, H, W, M, q = WORKLOADS[workload]
    sb = {"f32": 4, "f16": 2, "bf16": 2}[dtype_tag]
    strong = scaling == "strong" and world > 1
    B = B_total // world if strong else B_total
    alg = algorithmic_bytes_per_element(M, sb) * B * C * H * W
    if strong and allreduce:
        par = (f"batch of {B_total} sharded x{world} ({B} images per GPU); one NCCL all-reduce(sum) of the posterior's batch-axis sum row "
               f"({C * H * W * 4 // 1024} KB) per step")
    elif strong:
        par = (f"batch of {B_total} sharded x{world}: {B} images per GPU, each shard a batch of its own (the reference's mp.spawn slicing), "
               "no collective")
    else:
        par = f"{world} independent batches of {B} images, no collective"
    return {"workload": workload, "global_batch": B_total if strong else B_total * world, "shape": [C, H, W], "M": M, "q": q,
            "score_dtype": dtype_tag, "chain": "F1c var(M+1) -> F2a quantile mask -> F5 posterior -> F3 DDIM (+F8 slot write)",
            "batch_sum": bool(batch_sum), "images_per_gpu": B, "parallelism": par,
            "l2": l2_note(alg)}


# ------------------------------------------------------------------------------------------- GPU arm
def eager_reference_step(eps, scores, sample, q, M, sc, batch_sum=True, allreduce=None):
    """The reference's eager torch expressions of the step, on whatever device the tensors live on (no library of this
    repository involved): uncertainty_guidance.py:101-120 (F1c, F2a, F5) + SU/scheduling_ddim_uncertainty_zigzag_centered.py:
    472-510 (F3, epsilon prediction, clip_sample).  What a user of the reference runs on the same B200."""
    pe = eps.float() if eps.dtype != torch.float32 else eps
    stacked = torch.stack([s.float() if s.dtype != torch.float32 else s for s in scores] + [pe], dim=0)
    u = torch.var(stacked, dim=0)
    shp = u.shape
    thr = torch.quantile(u.flatten(1).to(torch.float32), q, dim=1, keepdim=True).view(shp[0], *([1] * (len(shp) - 1)))
    mask = (u > thr).float()
    inv_var = 1 / u
    S = pe.sum(dim=0) if batch_sum else pe
    if batch_sum and allreduce is not None:      # batch sharded over ranks: the reference's sum runs over the WHOLE batch
        allreduce(S)
    post = (1 / (M * inv_var + 1 / sc["alpha_hat"])) * (inv_var * S)
    eg = pe * (1 - mask) + mask * post
    x0 = ((sample - sc["sqrt_beta_t"] * eg) / sc["sqrt_alpha_t"]).clamp(-1.0, 1.0)
    prev = sc["sqrt_alpha_prev"] * x0 + sc["dir_coef"] * eg
    return u, thr.flatten(), mask, prev


class StepBench:
    """One (workload, score dtype, per-GPU batch) instance of the step on a device: resident inputs, the prepared fused launch,
    CUDA graphs of K steps, parity check and timings."""
    PREV_RING = 8     # x_{t-1} goes to a ring of buffers (8 x 25 MB at ImageNet-128): the 126 MB L2 cannot absorb the writes

    def __init__(self, ops, workload, dtype_name, B, dev, seed, batch_sum=True, unfused=False, allreduce=None, in_ring=None):
        self.ops, self.dev, self.workload, self.dtype_name = ops, dev, workload, dtype_name
        _, C, H, W, M, q = WORKLOADS[workload]
        # allreduce: a callable that sums a tensor over the ranks in place — given when the workload's batch is SHARDED over the
        # ranks (strong scaling): the posterior's batch-axis sum then spans every rank's images (distributed.allreduce_batch_sum)
        self.allreduce = allreduce
        self.B, self.C, self.H, self.W, self.M, self.q = B, C, H, W, M, q
        self.batch_sum = bool(batch_sum) and (B > 1 or allreduce is not None)
        self.dtype = DTYPES[dtype_name]
        self.sc = ddim_scalars()
        self.coeffs = ops.make_coeffs(self.sc["sqrt_alpha_t"], self.sc["sqrt_beta_t"], self.sc["sqrt_alpha_prev"], self.sc["dir_coef"],
                                      clip_sample=True)
        self.h = synth_host(B, C, H, W, M, self.dtype, seed, pin=True)
        h_eps, h_scores, h_sample = self.h
        self.eps, self.scores, self.sample = h_eps.to(dev), [s.to(dev) for s in h_scores], h_sample.to(dev)
        self.maps = torch.zeros(B, T_UC, C, H, W, device=dev, dtype=torch.float32)      # F8 accumulation buffer
        self.S = torch.empty(C, H, W, device=dev, dtype=torch.float32)
        self.prevs = [torch.empty(B, C, H, W, device=dev, dtype=torch.float32) for _ in range(self.PREV_RING)]
        self.n_el = B * C * H * W
        self.sb = 4 if self.dtype == torch.float32 else 2
        # input sets the steps rotate over (input_ring): set 0 is the one above, the others are bit-identical copies of it
        self.in_ring = input_ring(self.alg_bytes()) if in_ring is None else max(1, int(in_ring))
        self.inputs = [(self.eps, self.scores, self.sample)]
        for _ in range(self.in_ring - 1):
            self.inputs.append((self.eps.clone(), [s.clone() for s in self.scores], self.sample.clone()))
        self.fused = (not unfused) and ops.fused_supported(C * H * W, self.dtype) > 0
        self.plan, self.plans = None, []
        if self.fused:
            for eps_r, scores_r, sample_r in self.inputs:
                self.plans.append(ops.FusedStep(scores_r, eps_r, sample_r, q, self.coeffs, self.sc["alpha_hat"],
                                                S=self.S if self.batch_sum else None, S_broadcast=self.batch_sum,
                                                map_out=self.maps[:, 0], prev_out=self.prevs[0]))
            self.plan = self.plans[0]
        self.kernel = "moments_kernel"

    def alg_bytes(self):
        return algorithmic_bytes_per_element(self.M, self.sb) * self.n_el

    def step(self, i):
        ops, slot = self.ops, self.maps[:, i % T_UC]
        eps, scores, sample = self.inputs[i % self.in_ring]
        if self.fused:
            plan = self.plans[i % self.in_ring]
            plan.set_map_out(slot)
            plan.set_prev_out(self.prevs[i % self.PREV_RING])
            if self.batch_sum and self.allreduce is not None:
                ops.batch_sum(eps, out=self.S)           # this rank's images ...
                self.allreduce(self.S)                   # ... summed over the ranks (NCCL all-reduce of one [C,H,W] row)
                return plan.launch()["prev"]
            if self.batch_sum:                   # the reference's `pred_epsilon.sum(dim=0)` (uncertainty_guidance.py:119)
                return plan.launch_with_batch_sum(eps, self.S)["prev"]      # du_batch_sum, then the step as its dependent launch
            return plan.launch()["prev"]
        if self.batch_sum:
            ops.batch_sum(eps, out=self.S)
        u = ops.moments(scores, center=eps, mode="var_with_center", out=slot)
        thr = ops.quantile_threshold(u, self.q)
        r = ops.guided_step(eps, sample, self.coeffs, guidance="posterior", u=u, thr=thr,
                            aux=self.S if self.batch_sum else eps, aux_broadcast=self.batch_sum, post_M=float(self.M),
                            inv_alpha_hat=1.0 / self.sc["alpha_hat"], want_eps=False)
        return r["prev"]

    def kernel_only(self, i):
        """the dominant kernel alone (roofline.achieved = its algorithmic bytes / its average launch duration)"""
        if self.fused:
            plan = self.plans[i % self.in_ring]
            plan.set_map_out(self.maps[:, i % T_UC])
            plan.set_prev_out(self.prevs[i % self.PREV_RING])
            plan.launch()
        else:
            eps, scores, _ = self.inputs[i % self.in_ring]
            self.ops.moments(scores, center=eps, mode="var_with_center", out=self.maps[:, i % T_UC])

    def capture(self, fn, steps):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(steps):
                fn(i)
        return g

    def warm(self, n):
        for i in range(n):
            self.step(i)
        torch.cuda.synchronize()
        if self.fused:
            self.kernel = self.ops.fused_last_kernel() or "fused_step_kernel"

    def parity(self):
        """The timed configuration once, outside the timed region, against the reference's eager torch expressions on the same
        device: map within 1e-5 relative, thresholds bit-equal to torch.quantile of the kernel's own map, masks identical away
        from the threshold (and different on < 0.1 % of the pixels overall), x_{t-1} within 1e-5 of max(|x|, 0.4) where the masks
        agree.  Returns a dict for the bench line; raises on failure."""
        prev = self.step(0)
        u_k = self.maps[:, 0]
        u, thr, mask, prev_ref = eager_reference_step(self.eps, self.scores, self.sample, self.q, self.M, self.sc, self.batch_sum,
                                                      self.allreduce)
        torch.cuda.synchronize()
        # The bar for the map is the EXACT variance of the fp32 inputs (fp64 on the same device): within 1e-5 relative.  torch.var's
        # own fp32 result is reported next to it: on pixels whose variance is far below the typical one it deviates from the exact
        # value by more than that (its running mean is rounded at the magnitude of the samples, ~1e-7 absolute, against deviations
        # of ~1e-3), so "equal to the reference's fp32 value to 1e-5" is not a property any implementation can have there; the
        # kernel's shifted sums subtract nearby fp32 numbers exactly (Sterbenz) and stay at the 1e-6 level everywhere.
        u64 = torch.var(torch.stack([s.double() for s in self.scores] + [self.eps.double()], dim=0), dim=0)
        den = u64.abs().clamp_min(1e-300)
        rel_u = float(((u_k.double() - u64).abs() / den).max())
        rel_ref = float(((u.double() - u64).abs() / den).max())
        out = {"map_max_rel_err_vs_exact": rel_u, "reference_fp32_max_rel_err_vs_exact": rel_ref,
               "map_max_rel_diff_vs_reference_fp32": float(((u_k - u).abs() / u.abs().clamp_min(1e-30)).max())}
        del u64, den
        if rel_u > 1e-5:
            raise AssertionError(f"parity: map differs from the exact variance by {rel_u:.3e} relative")
        # against the reference's fp32 value: within 1e-5 plus the reference's own deviation from the exact value
        if out["map_max_rel_diff_vs_reference_fp32"] > 1e-5 + 1.01 * rel_ref:
            raise AssertionError(f"parity: map differs from torch.var by {out['map_max_rel_diff_vs_reference_fp32']:.3e} relative")
        if self.fused:
            thr_k = self.plan.res["thr"]
            # torch.quantile of the kernel's own map, on the host.  The two order statistics are exact; the final lerp is rounded
            # once per operation by the kernel (lerp_fma = 0), while torch contracts it into an FMA in its CUDA kernel and in the
            # CPU kernels of FMA-capable builds: at most one unit in the last place apart, and for rows of thousands of elements
            # (adjacent order statistics close together) practically always identical — anything beyond that last bit is a wrong
            # order statistic and fails.
            thr_c = torch.quantile(u_k.flatten(1).cpu(), self.q, dim=1)
            out["thr_bit_exact"] = bool(torch.equal(thr_k.cpu(), thr_c))
            out["thr_max_ulp_vs_host_torch"] = ulp_distance(thr_k, thr_c)
            if out["thr_max_ulp_vs_host_torch"] > 1:
                raise AssertionError("parity: thresholds differ from torch.quantile of the map by more than the lerp's last bit")
            mask_k = (u_k > thr_k.view(-1, 1, 1, 1)).float()
        else:
            mask_k = (u_k > thr.view(-1, 1, 1, 1)).float()
        agree = mask_k == mask
        out["mask_agreement"] = float(agree.float().mean())
        near = (u - thr.view(-1, 1, 1, 1)).abs() <= 1e-5 * thr.view(-1, 1, 1, 1).abs()
        if bool((~agree & ~near).any()) or out["mask_agreement"] < 0.999:
            raise AssertionError("parity: masks differ away from the threshold")
        fin = torch.isfinite(prev_ref) & agree
        err = (prev - prev_ref).abs()[fin]
        bound = 1e-5 * prev_ref.abs().clamp_min(0.4)[fin]
        out["prev_max_abs_err"] = float(err.max()) if err.numel() else 0.0
        if bool((err > bound).any()):
            raise AssertionError(f"parity: x_(t-1) off by {out['prev_max_abs_err']:.3e}")
        if not bool(torch.equal(torch.isfinite(prev)[agree], torch.isfinite(prev_ref)[agree])):
            raise AssertionError("parity: non-finite values in different places")
        return out

    def time_eager_reference(self, reps=5):
        """the reference's eager expressions on this GPU, CUDA events, best of `reps` after one warm-up"""
        best = None
        for r in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eager_reference_step(self.eps, self.scores, self.sample, self.q, self.M, self.sc, self.batch_sum, self.allreduce)
            e1.record()
            torch.cuda.synchronize()
            if r > 0:
                ms = e0.elapsed_time(e1)
                best = ms if best is None or ms < best else best
        return best

    def quick(self, steps):
        """(ms per step, ms per launch of the dominant kernel) from CUDA graphs of `steps` steps, one warm replay each"""
        self.warm(3)
        use_graph = self.fused
        out = []
        for fn in (self.step, self.kernel_only):
            if use_graph:
                g = self.capture(fn, steps)
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if use_graph:
                g.replay()
            else:
                for i in range(steps):
                    fn(i)
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / steps)
        return out


def run_ours(args):
    import torch.distributed as dist
    from diffusion_uncertainty_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the uncertainty path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B_total, C, H, W, M, q = WORKLOADS[args.workload]
    # BASELINE.json configs[2]: "batch 128 sharded across 1/2/4/8" — the default for N > 1 is the STRONG split (B_total / N images
    # per GPU); --scaling weak keeps B_total images on every GPU (N independent replicas of the N=1 job)
    scaling = args.scaling or ("strong" if world > 1 else "weak")
    if scaling == "strong" and B_total % world != 0:
        raise SystemExit(f"--scaling strong needs the batch ({B_total}) to divide over {world} ranks")
    B = B_total // world if scaling == "strong" else B_total

    def allreduce_sum(t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    # Sharded batch: by default every rank treats its shard as a batch of its own — the reference's multi-GPU mode
    # (scripts/generate_dataset_score_uncertainty_imagenet.py:51, 137-144: mp.spawn, one process per GPU, each with its own batches, so
    # its `pred_epsilon.sum(dim=0)` runs over the process's batch) — and the step has no collective.  --allreduce-batch-sum makes the
    # ranks share ONE posterior sum over the whole batch of 128 instead (an NCCL all-reduce of the [C,H,W] row inside every step).
    sharded = world > 1 and scaling == "strong" and bool(args.batch_sum) and args.allreduce_batch_sum
    sb_ = StepBench(ops, args.workload, args.dtype, B, dev, 1234 + rank + args.seed_offset, batch_sum=args.batch_sum, unfused=args.unfused,
                    allreduce=allreduce_sum if sharded else None)
    fused = sb_.fused

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sb_.warm(args.warmup)
    barrier()
    dominant = sb_.kernel
    parity = None if args.no_parity else sb_.parity()
    step, kernel_only = sb_.step, sb_.kernel_only

    # The K timed steps are captured into ONE CUDA graph (every C-ABI call is capturable: no host reads, no allocation),
    # so the timed region holds exactly K steps of GPU work and no Python / launch latency between them.  The unfused
    # chain allocates per call and is timed eagerly.
    use_graph = fused and not args.eager
    launches0 = ops.launch_count
    if use_graph:
        g_step, g_kernel = sb_.capture(step, args.steps), sb_.capture(kernel_only, args.steps)
        n_step_launches = args.steps * (2 if sb_.batch_sum else 1)        # (library launches; the NCCL all-reduce is not counted)
        g_step.replay(); g_kernel.replay()            # one untimed replay each
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        # The K-step region lasts a few milliseconds, less than one nvidia-smi sampling period, so the sampler runs over a
        # sustained window of the SAME work: untimed replays before (until the first sample has arrived and the clocks have
        # ramped), the timed K steps, untimed replays after.  Throttle reasons seen anywhere in the window are reported.
        # (with a collective inside the step every rank must run the SAME number of replays: the count for a given duration is agreed
        # on once — slowest rank's replay time, all-reduced — instead of being timed per rank)
        replay_s = [None]

        def sustain(seconds):
            if sharded:
                if replay_s[0] is None:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    g_step.replay() if use_graph else [step(i) for i in range(args.steps)]
                    torch.cuda.synchronize()
                    t = torch.tensor([time.perf_counter() - t0], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    replay_s[0] = max(t.item(), 1e-5)
                for _ in range(max(1, min(2000, int(seconds / replay_s[0])))):
                    g_step.replay() if use_graph else [step(i) for i in range(args.steps)]
                torch.cuda.synchronize()
                return
            t_end = time.perf_counter() + seconds
            while time.perf_counter() < t_end:
                if use_graph:
                    g_step.replay()
                else:
                    for i in range(args.steps):
                        step(i)
                torch.cuda.synchronize()

        if sharded:
            sustain(0.6)       # (a fixed, rank-independent amount of work; the sampler has its first rows by then)
        else:
            t_wait = time.perf_counter() + 3.0
            while not clocks.rows and time.perf_counter() < t_wait:
                sustain(0.05)
            sustain(0.3)
        barrier()
        # The ranks leave sustain() at different moments, so a GPU may have idled for tens of milliseconds at the barrier and
        # dropped its clocks: one untimed replay of the same K steps re-warms it, back to back with the timed one (no host
        # synchronisation in between; the events still bracket exactly K steps).
        if use_graph:
            g_step.replay()
        else:
            for i in range(min(args.steps, 10)):
                step(i)
        e0.record()
        if use_graph:
            g_step.replay()
        else:
            launches0 = ops.launch_count
            for i in range(args.steps):
                out = step(i)
            n_step_launches = ops.launch_count - launches0
        e1.record()
        barrier()
        # the dominant kernel alone, K launches back to back on the same stream
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        if use_graph:
            g_kernel.replay()
        else:
            for i in range(args.steps):
                kernel_only(i)
        k1.record()
        barrier()
        sustain(0.3)
    ms = e0.elapsed_time(e1)
    launches = n_step_launches
    k_ms = k0.elapsed_time(k1) / args.steps
    # eager cross-check: the same K steps launched from Python (includes per-launch host latency)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for i in range(args.steps):
        out = step(i)
    b1.record()
    torch.cuda.synchronize()
    eager_ms = b0.elapsed_time(b1) / args.steps
    ms_ranks = [ms]
    if world > 1:
        t = torch.tensor([ms], device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        ms_ranks = [x.item() for x in allt]
        ms = max(ms_ranks)
    ms_per_step = ms / args.steps
    value = world * B * H * W / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the public API with HOST buffers (pinned): H2D of the step's inputs and D2H of its results
    # inside the timed region (diffusion_uncertainty_b200.host_step pipelines image chunks over three streams)
    from diffusion_uncertainty_b200.host_step import HostStreamedUncertaintyStep
    h_eps, h_scores, h_sample = sb_.h
    h_prev = torch.empty(B, C, H, W, dtype=torch.float32).pin_memory()
    h_map = torch.empty(B, C, H, W, dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    host_step = HostStreamedUncertaintyStep(B, (C, H, W), M, dev, score_dtype=sb_.dtype, chunks=args.e2e_chunks)

    def e2e_step(i):
        return host_step(h_scores, h_eps, h_sample, q, sb_.coeffs, sb_.sc["alpha_hat"], h_prev, h_map, batch_sum=sb_.batch_sum,
                         map_slot=sb_.maps[:, i % T_UC])

    e2e_step(0).synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        last = e2e_step(i)
    last.synchronize()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_val = world * B * H * W * e2e_steps / e2e_s / 1e6
    sb = sb_.sb
    n_el = sb_.n_el
    h2d = n_el * ((M + 1) * sb + 4)
    d2h = n_el * 8

    # ---- sub-records measured by every rank (collectives inside): the other scaling mode, the sampling loop
    other = None
    if world > 1 and not args.no_extras:
        oB = B_total if scaling == "strong" else B_total // world
        ob = StepBench(ops, args.workload, args.dtype, oB, dev, 1234 + rank + args.seed_offset, batch_sum=args.batch_sum,
                       allreduce=allreduce_sum if (scaling == "weak" and bool(args.batch_sum) and args.allreduce_batch_sum) else None)
        o_ms, o_k = ob.quick(args.steps)
        t = torch.tensor([o_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        other = {"scaling": "weak" if scaling == "strong" else "strong", "images_per_gpu": oB, "ms_per_step": t.item(),
                 "value": world * oB * H * W / (t.item() * 1e-3) / 1e6, "unit": "Mpix/s", "kernel": ob.kernel, "kernel_ms": o_k,
                 "l2": l2_note(ob.alg_bytes())}
        del ob
    loop = None
    if args.with_loop and args.workload == "imagenet128_adm_b128_m5":
        del host_step
        torch.cuda.empty_cache()
        try:
            loop = loop_record(args, "imagenet128_adm_loop", world, rank, local, dev, dist)
        except Exception as ex:          # a sub-record must not take the headline line down
            loop = {"error": repr(ex)[:300]} if rank == 0 else None

    if rank == 0:
        peak, peak_kind = peaks()
        alg_step = sb_.alg_bytes()
        # dominant kernel: the fused step moves exactly the step's algorithmic bytes; unfused: moments = M scores + eps in, u out
        alg_kernel = alg_step if fused else ((M + 1) * sb + 4) * n_el
        achieved = alg_kernel / (k_ms * 1e-3) / 1e9
        dt_tag = {"fp32": "f32", "fp16": "f16", "bf16": "bf16"}[args.dtype]
        cfg = step_config(args.workload, B_total, world, sb_.batch_sum, dt_tag, scaling, sharded)
        run = {"fused_single_launch": bool(fused), "input_copies": sb_.in_ring,
               "prev_out": f"ring of {StepBench.PREV_RING} buffers ({StepBench.PREV_RING * n_el * 4 / 1e6:.0f} MB): x_(t-1) is written to HBM, not absorbed by L2"}
        line = {
            "metric": "uncertainty_step_throughput", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": dt_tag, "data": "synthetic", "config": cfg, "run": run,
            "parity_checked": parity is not None, "parity": parity,
            "step_hbm_frac": alg_step / (ms_per_step * 1e-3) / 1e9 / peak,
            "step_algorithmic_GBps": alg_step / (ms_per_step * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_kind": peak_kind, "traffic": ncu_traffic(dominant, args.workload, B, args.dtype),
                         "kernel_ms": k_ms, "algorithmic_bytes": alg_kernel,
                         "timing": "CUDA events around %d back-to-back launches%s" % (args.steps, " replayed from one CUDA graph" if use_graph else "")},
            "ms_per_step_by_rank": [m / args.steps for m in ms_ranks],
            "ms_per_step_eager_python_loop": eager_ms,
            "timed_region": "one CUDA graph of K steps" if use_graph else "K eager steps",
            "e2e": {"value": e2e_val, "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / e2e_steps * 1e3},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
        }
        if other is not None:
            line["other_scaling"] = other
        if loop is not None:
            line["sampling_loop"] = loop
        if world == 1 and not args.no_extras:
            # the reference's eager torch expressions on this GPU: the bar a user of the reference compares against
            try:
                eg_ms = sb_.time_eager_reference()
                line["gpu_eager_baseline"] = {"value": B * H * W / (eg_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": eg_ms,
                                              "what": "the reference's eager torch expressions (stack / var / quantile / blend / DDIM) on the "
                                                      "same GPU and tensors, CUDA events, best of 5", "speedup_of_value": eg_ms / ms_per_step}
            except Exception as ex:      # informational legs must not take the headline line down
                line["gpu_eager_baseline"] = {"error": repr(ex)[:200]}
            # the other BASELINE shapes and the autocast score dtype, same measurement in short form (parity-checked each)
            subs = {}
            for wl, dtn in ([(args.workload, "fp16")] if args.dtype == "fp32" else []) + [(w, "fp32") for w in WORKLOADS if w != args.workload]:
                try:
                    x = StepBench(ops, wl, dtn, WORKLOADS[wl][0], dev, 1234, batch_sum=args.batch_sum)
                    x.warm(3)
                    par = x.parity()
                    s_ms, kk_ms = x.quick(args.steps)
                    subs[f"{wl}:{dtn}"] = {"ms_per_step": s_ms, "kernel": x.kernel, "kernel_ms": kk_ms,
                                           "value": x.B * x.H * x.W / (s_ms * 1e-3) / 1e6, "unit": "Mpix/s",
                                           "roofline_frac": x.alg_bytes() / (kk_ms * 1e-3) / 1e9 / peak, "parity_checked": True,
                                           "mask_agreement": par["mask_agreement"], "l2": l2_note(x.alg_bytes())}
                    del x
                except Exception as ex:   # a sub-record must not take the headline line down
                    subs[f"{wl}:{dtn}"] = {"error": repr(ex)[:200]}
            line["other_workloads"] = subs
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = time_cpu(args.workload, sample_images=B_total, budget_s=12.0)
        print(json.dumps(line), flush=True)
    del out
    if world > 1:
        # the graphs may hold captured NCCL kernels: release them (and drain the device) BEFORE the communicator goes away
        if use_graph:
            g_step = g_kernel = None
        other = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- M-sharded arm
def run_msharded(args):
    """BASELINE config 5 (Stable Diffusion 512 px: one 4x64x64 latent, M = 16 perturbed predictions SHARDED over the ranks):
    every rank reduces its M/R predictions to per-element partial moments (du_moments, DU_MOM_PARTIAL_M2), ONE all-gather of
    the packed (mean, M2) pairs over NCCL, du_moments_merge (Chan, rank order), then the latent-sized rest of the step
    (quantile, posterior blend, DDIM) replicated on every rank.  Strong scaling of one step; latency-bound by design
    (1.3 MB of scores in total) — the line reports microseconds per step next to the Mpix/s."""
    import torch.distributed as dist
    from diffusion_uncertainty_b200 import distributed as D
    from diffusion_uncertainty_b200 import ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, C, H, W, M, q = WORKLOADS[args.workload]
    sc = ddim_scalars()
    coeffs = ops.make_coeffs(sc["sqrt_alpha_t"], sc["sqrt_beta_t"], sc["sqrt_alpha_prev"], sc["dir_coef"], clip_sample=False)
    eps_h, scores_h, sample_h = synth_host(B, C, H, W, M, torch.float32, 1234, pin=False)     # same data on every rank
    a, b = D.shard_range(M, rank, world)
    eps, sample, mine = eps_h.to(dev), sample_h.to(dev), [s.to(dev) for s in scores_h[a:b]]
    sm = D.ShardedMoments()

    def step(i):
        u = sm.reduce(mine, eps, "var_with_center", total_M=M)
        thr = ops.quantile_threshold(u, q)
        return ops.guided_step(eps, sample, coeffs, guidance="posterior", u=u, thr=thr, aux=eps, post_M=float(M),
                               inv_alpha_hat=1.0 / sc["alpha_hat"], want_eps=False)["prev"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 10)):
        out = step(i)
    barrier()
    # K steps in ONE CUDA graph (the kernels and the NCCL all-gather are capturable: no host reads with total_M given), so the
    # timed region holds GPU work only; --eager times K steps launched from Python instead
    graph = None
    if not args.eager:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                step(i)
        torch.cuda.current_stream().wait_stream(side)
        barrier()
        launches0 = ops.launch_count
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(args.steps):
                out = step(i)
        launches = ops.launch_count - launches0
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        for i in range(3 if graph is not None else 200):
            graph.replay() if graph is not None else step(i)
        barrier()
        launches0 = ops.launch_count
        e0.record()
        if graph is not None:
            graph.replay()
        else:
            for i in range(args.steps):
                out = step(i)
        e1.record()
        if graph is None:
            launches = ops.launch_count - launches0
        barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        # all ranks must hold the same x_{t-1}
        ref = out.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([int(torch.equal(ref, out))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        identical = bool(same.item())
    else:
        identical = True
    if rank == 0:
        ms_per_step = ms / args.steps
        n_el = B * C * H * W
        line = {"metric": "uncertainty_step_throughput", "value": B * H * W / (ms_per_step * 1e-3) / 1e6, "unit": "Mpix/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 10), "ms_per_step": ms_per_step,
                "us_per_step": ms_per_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "shape": [B, C, H, W], "M": M, "q": q,
                           "parallelism": f"M sharded x{world} ({b - a} predictions per rank), one NCCL all-gather of (mean, M2), "
                                          "replicated quantile / posterior / DDIM",
                           "chain": "du_moments(partial) -> all_gather -> du_moments_merge -> du_quantile_threshold -> du_guided_step",
                           "l2": "1.3 MB of scores: latency-bound, L2-resident"},
                "timed_region": "one CUDA graph of K steps (kernels + NCCL all-gather)" if graph is not None else "K eager steps",
                "collective_bytes_per_step": 2 * n_el * 4 * world, "x_prev_identical_on_all_ranks": identical,
                "gpu_launches": launches, "clocks": clocks.summary("200 untimed steps + the K timed steps")}
        print(json.dumps(line), flush=True)
    # the graph holds captured NCCL kernels: release it (and drain the device) BEFORE the communicator goes away, otherwise
    # the teardown can wait forever
    graph = None
    del out
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- sampling-loop arm
LOOPS = {
    # name: (feeder factory, total batch, C, H, betas, scheduler kwargs, generation steps)
    "imagenet128_adm_loop": ("adm_imagenet128", 128, 3, 128, dict(beta_start=1e-4, beta_end=0.02, beta_schedule="linear"),
                             dict(M=5, after_step=40, num_steps_uc=10, num_zigzag=3), 50),
    # BASELINE.json configs[1]: ImageNet-64 ADM (dropout 0.5, cosine schedule: init_model.py:45-47, 134-137), same window
    "imagenet64_adm_loop": ("adm_imagenet64", 128, 3, 64, dict(beta_schedule="squaredcos_cap_v2"),
                            dict(M=5, after_step=40, num_steps_uc=10, num_zigzag=3), 50),
}


def loop_record(args, workload, world, rank, local, dev, dist):
    """BASELINE.json metric (iii): ImageNet-128 M=5 img/s over the FULL sampling loop of the README command
    (scripts/generate_dataset_score_uncertainty_imagenet.py: 50 DDIM steps, uncertainty window = the last 10, M=5 x num_zigzag=3
    perturbed forwards per window step, maps accumulated and copied to the host), through the drop-in
    `generate_samples_model_scheduler_class_conditioned_from_tensor` with the zigzag-centred scheduler, under torch.autocast as
    the reference runs it.  The score model is a random-init ADM-128-shaped feeder (tools/adm_feeder.py).  The batch of 128 is
    SHARDED over the ranks (strong scaling, no collective) — the reference's mp.spawn slicing.  One step = one whole loop.
    Called by every rank (the process group, if any, is already up); returns the record on rank 0, None elsewhere."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import adm_feeder
    from diffusion_uncertainty_b200 import ops
    from diffusion_uncertainty_b200.generate_samples import generate_samples_model_scheduler_class_conditioned_from_tensor as gen_loop
    from diffusion_uncertainty_b200.schedulers_uncertainty.scheduling_ddim_uncertainty_zigzag_centered import \
        DDIMSchedulerUncertaintyImagenetClassConditioned as Sched

    factory, B_total, C, H, betas, skw, n_steps = LOOPS[workload]
    B = B_total // world
    model = getattr(adm_feeder, factory)().to(dev).eval()
    sched = Sched.from_config(dict(num_train_timesteps=1000, clip_sample=True, set_alpha_to_one=True, steps_offset=0,
                                   prediction_type="epsilon", timestep_spacing="leading", **betas), unet=model, **skw)
    sched.set_timesteps(n_steps)
    g = torch.Generator().manual_seed(49394 + rank)
    X_T = torch.randn(B, C, H, H, generator=g).pin_memory()
    y = torch.randint(0, 1000, (B,), generator=g)
    torch.manual_seed(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_loop():
        with torch.autocast("cuda"):
            return gen_loop(X_T, y, B, dev, model, sched)

    warm = max(1, min(args.warmup, 1)) if args.loop_warmup is None else args.loop_warmup
    for _ in range(warm):
        res = one_loop()
    barrier()
    steps = args.loop_steps
    launches0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            res = one_loop()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), wall * 1e3)     # the loop ends with host-side waits on the pinned copies: wall covers them
    launches = ops.launch_count - launches0
    # the share of the score model: one forward of the same batch, timed alone
    t_tensor = torch.full((B,), 180, device=dev, dtype=torch.long)
    xg, yg = X_T.to(dev), y.to(dev)
    with torch.no_grad(), torch.autocast("cuda"):
        for _ in range(2):
            model(xg, t_tensor, y=yg)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(5):
            model(xg, t_tensor, y=yg)
        f1.record()
    torch.cuda.synchronize()
    fwd_ms = f0.elapsed_time(f1) / 5
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    n_fwd = n_steps + skw["num_steps_uc"] * skw["M"] * skw["num_zigzag"]
    if rank != 0:
        return None
    ms_per_loop = ms / steps
    unc = res["uncertainty"]
    share = n_fwd * fwd_ms / ms_per_loop
    return {
        "metric": "imagenet%d_m5_sampling_loop_throughput" % H, "value": B_total / (ms_per_loop * 1e-3), "unit": "img/s",
        "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": ms_per_loop, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f16 autocast (model) / f32 (uncertainty path)", "data": "synthetic",
        "config": {"workload": workload, "global_batch": B_total, "batch_per_gpu": B, "generation_steps": n_steps,
                   "start_step_uc": skw["after_step"], "num_steps_uc": skw["num_steps_uc"], "M": skw["M"],
                   "num_zigzag": skw["num_zigzag"], "scheduler": "uncertainty_zigzag_centered",
                   "model": "random-init ADM-%d-shaped feeder (tools/adm_feeder.py, %.1f M parameters)"
                            % (H, sum(p.numel() for p in model.parameters()) / 1e6),
                   "parallelism": f"batch-sharded x{world}, no collective", "step": "one full sampling loop of the batch"},
        "model_forwards_per_loop": n_fwd, "model_forward_ms": fwd_ms,
        "model_share_of_loop": share,
        "non_model_ms_per_loop": ms_per_loop - n_fwd * fwd_ms,
        "limiter": "the reference's PyTorch score model: %d forwards x %.1f ms at %d images per GPU = %.1f %% of the loop"
                   % (n_fwd, fwd_ms, B, 100 * share),
        "e2e": {"value": B_total / (ms_per_loop * 1e-3), "unit": "img/s", "h2d_bytes_per_step": X_T.numel() * 4 * world,
                "d2h_bytes_per_step": (2 * unc.numel() * unc.element_size() + res["gen_images"].numel()) * world,
                "note": "the loop itself is end to end: X_T comes from pinned host memory, maps / scores / uint8 images end in host memory"},
        "gpu_launches": launches, "map_shape": list(unc.shape), "map_finite": bool(torch.isfinite(unc).all()),
        "clocks": clocks.summary("the timed loops"),
    }


def run_loop(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the uncertainty path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = loop_record(args, args.workload, world, rank, local, dev, dist)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="imagenet128_adm_b128_m5", choices=list(WORKLOADS) + list(LOOPS))
    ap.add_argument("--loop-steps", type=int, default=1, help="sampling-loop workloads: timed loops (one loop = one step)")
    ap.add_argument("--loop-warmup", type=int, default=None, help="sampling-loop workloads: untimed warm-up loops (default 1)")
    ap.add_argument("--dtype", default="fp32", choices=list(DTYPES))
    ap.add_argument("--batch-sum", type=int, default=1, help="1 = reference behaviour (posterior sum over the batch axis)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--unfused", action="store_true", help="time the 3-kernel chain instead of the single fused launch")
    ap.add_argument("--e2e-chunks", type=int, default=1, help="image chunks of the host-buffer pipeline (e2e leg)")
    ap.add_argument("--shard-m", action="store_true", help="shard the M predictions over the ranks (SD-512 latent workload) instead of the batch")
    ap.add_argument("--seed-offset", type=int, default=0, help="added to the data seed 1234 + rank (to replay another rank's data on one GPU)")
    ap.add_argument("--eager", action="store_true", help="time K eager launches from Python instead of one CUDA graph of K steps")
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"],
                    help="N > 1: strong = the workload's batch split over the ranks (default, BASELINE configs[2]); weak = the whole batch on every GPU")
    ap.add_argument("--allreduce-batch-sum", action="store_true",
                    help="N > 1, strong: one posterior batch-axis sum over ALL ranks' images (NCCL all-reduce of the [C,H,W] row in every step)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check of the timed configuration")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records (other scaling mode, other workloads, fp16, eager-torch GPU baseline)")
    ap.add_argument("--no-loop", dest="with_loop", action="store_false", help="skip the ImageNet-128 sampling-loop sub-record (img/s, ~1 min per loop at N=1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload in LOOPS:
        if args.impl == "reference":
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the sampling loop needs the score model on a GPU; the reference "
                                  "CPU arm exists for the uncertainty-step workloads only"}))
            return
        run_loop(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.shard_m:
        run_msharded(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
