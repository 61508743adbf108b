"""debug: determinism of the three-phase fused kernel on the SD latent shape (B=1, 4x64x64, M=16, skip_ddim)"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_uncertainty_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
eps = torch.randn(1, 4, 64, 64, generator=g).to(dev)
preds = [(eps.cpu() + 0.05 * torch.randn(1, 4, 64, 64, generator=g)).to(dev) for _ in range(16)]
a_hat = torch.tensor(0.3)
for skip in (True, False):
    outs = []
    for r in range(8):
        if skip:
            f = ops.uncertainty_step(preds, eps, None, 0.9, None, a_hat, fused=True, want_mask=True)
        else:
            c = ops.make_coeffs(0.5, 0.8, 0.6, 0.7)
            f = ops.uncertainty_step(preds, eps, eps.clone(), 0.9, c, a_hat, fused=True, want_mask=True, want_eps=True)
        torch.cuda.synchronize()
        outs.append({k: v.clone() for k, v in f.items() if v is not None})
    print("skip_ddim", skip, "kernel", ops.fused_last_kernel())
    for r in range(1, 8):
        print("  run", r, {k: int((outs[r][k].view(torch.int32) != outs[0][k].view(torch.int32)).sum()) for k in outs[0]})
