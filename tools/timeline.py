"""Summarise the per-CTA phase stamps written by DU_FUSED_TIMELINE (debug aid for the fused step kernel)."""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.float64)
t0 = a[:, 0].min()
names = ["start", "A done", "barrier1 passed", "locate0 done", "lists complete", "thr known", "C done"]
print(f"{len(a)} CTAs; all times in us relative to the first CTA start")
for k, nm in enumerate(names):
    col = a[:, k]
    col = col[col > 0]
    if len(col):
        r = (col - t0) / 1e3
        print(f"  {nm:16s} min {r.min():7.2f}  mean {r.mean():7.2f}  p90 {np.percentile(r, 90):7.2f}  max {r.max():7.2f}")
if a.shape[1] > 7 and (a[:, 7] > 0).any():
    r = (a[:, 7][a[:, 7] > 0] - t0) / 1e3
    print(f"  {'locate1 done':16s} min {r.min():7.2f}  mean {r.mean():7.2f}  p90 {np.percentile(r, 90):7.2f}  max {r.max():7.2f}")
    ok = (a[:, 4] > 0) & (a[:, 7] > 0)
    print(f"  {'lists complete':>16s} -> {'locate1 done':16s} mean {((a[ok, 7] - a[ok, 4]) / 1e3).mean():7.2f}")
d = np.diff(a[:, :7], axis=1) / 1e3
for k in range(6):
    ok = (a[:, k] > 0) & (a[:, k + 1] > 0)
    if ok.any():
        print(f"  {names[k]:>16s} -> {names[k+1]:16s} mean {d[ok, k].mean():7.2f}  max {d[ok, k].max():7.2f}")

# extra debug stamps (slots 8..15), relative to "lists complete"
for k in range(8, a.shape[1]):
    ok = (a[:, k] > 0) & (a[:, 4] > 0)
    if ok.any():
        print(f"  slot {k}: after lists complete mean {((a[ok, k] - a[ok, 4]) / 1e3).mean():7.2f}  max {((a[ok, k] - a[ok, 4]) / 1e3).max():7.2f}")
