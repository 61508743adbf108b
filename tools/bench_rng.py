#!/usr/bin/env python
"""F7 perturbation builders on the GPU: torch.randn_like + du_perturb (two launches, 16 B/element) against du_perturb_randn
(one launch, 8 B/element), ImageNet-128 b128 fp32.  Prints one JSON line per variant (CUDA events, L2 flushed by size)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_uncertainty_b200 import ops  # noqa: E402


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    for shape in [(128, 3, 128, 128), (128, 3, 64, 64), (1, 4, 64, 64)]:
        # rotate over enough distinct inputs that the working set exceeds the 126 MB L2
        n = 1
        for s in shape:
            n *= s
        copies = max(1, int(300e6 // (n * 4)) + 1)
        xs = [torch.randn(shape, device="cuda") for _ in range(copies)]
        i = [0]

        def two():
            x = xs[i[0] % copies]; i[0] += 1
            return ops.perturb(x, torch.randn_like(x), 0.99, 0.1)

        def one():
            x = xs[i[0] % copies]; i[0] += 1
            return ops.perturb_randn(x, 0.99, 0.1)

        def gen_only():
            x = xs[i[0] % copies]; i[0] += 1
            return torch.randn_like(x)

        t2, t1, tg = timeit(two), timeit(one), timeit(gen_only)
        print(json.dumps({"shape": shape, "randn_like+du_perturb_us": round(t2, 2), "du_perturb_randn_us": round(t1, 2),
                          "torch_randn_like_alone_us": round(tg, 2), "fused_GBps_algorithmic(8B/el)": round(n * 8 / t1 / 1e3, 1),
                          "frac_of_measured_hbm_peak": round(n * 8 / t1 / 1e3 / peak, 3), "speedup": round(t2 / t1, 2)}))


if __name__ == "__main__":
    main()
