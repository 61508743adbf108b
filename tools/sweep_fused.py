"""Tuning sweep of the fused step kernel in ONE process: every configuration is selected through the DU_FUSED_* environment
knobs (read by the library on every call), timed as a CUDA graph of K back-to-back launches (CUDA events, best of R replays).
Optionally writes the per-CTA phase timeline of each configuration (DU_FUSED_TIMELINE) and summarises it.

  python tools/sweep_fused.py "mapg:cluster:threads:tmem[:extra=val,...]" ... [--timeline] [--workload NAME] [--dtype fp32]
"""
import argparse
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from diffusion_uncertainty_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--timeline", action="store_true")
    ap.add_argument("--workload", default="imagenet128_adm_b128_m5")
    ap.add_argument("--dtype", default="fp32")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--replays", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/sweep_fused.jsonl")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, C, H, W, M, q = bench.WORKLOADS[a.workload]
    dtype = bench.DTYPES[a.dtype]
    sc = bench.ddim_scalars()
    coeffs = ops.make_coeffs(sc["sqrt_alpha_t"], sc["sqrt_beta_t"], sc["sqrt_alpha_prev"], sc["dir_coef"], clip_sample=True)
    h_eps, h_scores, h_sample = bench.synth_host(B, C, H, W, M, dtype, 1234, pin=False)
    eps, scores, sample = h_eps.to(dev), [s.to(dev) for s in h_scores], h_sample.to(dev)
    maps = torch.zeros(B, bench.T_UC, C, H, W, device=dev)
    S_buf = torch.empty(C, H, W, device=dev)
    ops.batch_sum(eps, out=S_buf)
    plan = ops.FusedStep(scores, eps, sample, q, coeffs, sc["alpha_hat"], S=S_buf, S_broadcast=True, map_out=maps[:, 0])
    sb = 4 if dtype == torch.float32 else 2
    alg = bench.algorithmic_bytes_per_element(M, sb) * B * C * H * W
    peak, _ = bench.peaks()
    knobs = ["DU_FUSED_MAPG", "DU_FUSED_CLUSTER", "DU_FUSED_THREADS", "DU_FUSED_TMEM"]
    ref_prev = None
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "a") as fo:
        for cfg in [c for arg in a.configs for c in arg.split()]:
            parts = cfg.split(":")
            for k in list(os.environ):
                if k.startswith("DU_FUSED_"):
                    del os.environ[k]
            for k, v in zip(knobs, parts[:4]):
                if v not in ("", "-"):
                    os.environ[k] = v
            if len(parts) > 4:
                for kv in parts[4].split(","):
                    k, v = kv.split("=")
                    os.environ["DU_FUSED_" + k] = v
            try:
                for i in range(3):
                    plan.set_map_out(maps[:, i % bench.T_UC])
                    out = plan.launch()
                torch.cuda.synchronize()
                prev = out["prev"].clone()
                if ref_prev is None:
                    ref_prev = prev
                same = bool(torch.equal(torch.nan_to_num(prev), torch.nan_to_num(ref_prev)))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(a.steps):
                        plan.set_map_out(maps[:, i % bench.T_UC])
                        plan.launch()
                g.replay()
                torch.cuda.synchronize()
                best = 1e9
                for _ in range(a.replays):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    g.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / a.steps)
                rec = {"cfg": cfg, "kernel_us": round(best * 1e3, 2), "frac": round(alg / (best * 1e-3) / 1e9 / peak, 4),
                       "same_prev_as_first_cfg": same}
            except Exception as ex:  # a configuration the library refuses is reported, not fatal
                rec = {"cfg": cfg, "error": str(ex)[:200]}
            print(json.dumps(rec), flush=True)
            fo.write(json.dumps(rec) + "\n")
            if a.timeline and "error" not in rec:
                tl = os.path.join(os.path.dirname(a.out), "timeline_" + cfg.replace(":", "_").replace("=", "").replace(",", "_") + ".txt")
                os.environ["DU_FUSED_TIMELINE"] = tl
                plan.launch()
                torch.cuda.synchronize()
                del os.environ["DU_FUSED_TIMELINE"]
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "timeline.py"), tl], capture_output=True, text=True)
                print(r.stdout, flush=True)
                with open(tl.replace(".txt", "_summary.txt"), "w") as f:
                    f.write(r.stdout)


if __name__ == "__main__":
    main()
