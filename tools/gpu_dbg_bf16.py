"""debug: is the fused step deterministic on the bf16 ImageNet-128 batch? which output differs between runs?"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from diffusion_uncertainty_b200 import ops
dev = torch.device("cuda:0")
for dt in ("bf16", "fp16", "fp32"):
    sb = bench.StepBench(ops, "imagenet128_adm_b128_m5", dt, 128, dev, 1234)
    sb.warm(2)
    outs = []
    for r in range(6):
        sb.prevs[0].zero_(); sb.maps[:, 0].zero_()
        sb.step(0); torch.cuda.synchronize()
        outs.append((sb.prevs[0].clone(), sb.maps[:, 0].clone(), sb.plan.res["thr"].clone()))
    u = outs[0][1]
    print(dt, "kernel", ops.fused_last_kernel(), "zeros in map", int((u == 0).sum()), "nan thr", int(torch.isnan(outs[0][2]).sum()))
    for r in range(1, 6):
        dp = outs[r][0] != outs[0][0]
        dp &= ~(torch.isnan(outs[r][0]) & torch.isnan(outs[0][0]))
        dm = (outs[r][1] != outs[0][1]).sum().item()
        dthr = ((outs[r][2] != outs[0][2]) & ~(torch.isnan(outs[r][2]) & torch.isnan(outs[0][2]))).sum().item()
        print("  run", r, "prev diffs", int(dp.sum()), "map diffs", dm, "thr diffs", dthr)
        if dp.any():
            idx = dp.nonzero()[:5]
            for i in idx:
                i = tuple(i.tolist())
                print("    at", i, outs[0][0][i].item(), outs[r][0][i].item(), "u", u[i].item(), "thr", outs[0][2][i[0]].item())
            imgs = dp.flatten(1).any(1).nonzero().flatten().tolist()
            print("    images with diffs:", imgs[:20], "count", len(imgs))
