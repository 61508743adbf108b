"""Summarise the per-CTA phase stamps written by DU_FUSED_TIMELINE (debug aid for the fused step kernels).
Slot names follow fused_pred_kernel (du_fused_pred.cu); the last column is the SM id."""
import sys

import numpy as np

a = np.loadtxt(sys.argv[1], dtype=np.float64)
a = a[a[:, 0] > 0]          # (the buffer is sized for the three-phase plan's cluster; unused rows stay zero)
t0 = a[:, 0].min()
ORDER = [(0, "start"), (1, "band known (pilot done)"), (2, "streaming done"), (3, "band verified (cluster barrier passed)"),
         (7, "fine bin located"), (8, "bin keys picked"), (10, "order statistics ranked"), (5, "threshold known"), (6, "patched, done")]
print(f"{len(a)} CTAs; all times in us relative to the first CTA start")
prev = None
for k, nm in ORDER:
    col = a[:, k]
    ok = col > 0
    if not ok.any():
        continue
    r = (col[ok] - t0) / 1e3
    line = f"  {nm:40s} min {r.min():7.2f}  mean {r.mean():7.2f}  p90 {np.percentile(r, 90):7.2f}  max {r.max():7.2f}"
    if prev is not None:
        both = ok & (a[:, prev] > 0)
        d = (a[both, k] - a[both, prev]) / 1e3
        line += f"   | since previous: mean {d.mean():6.2f} max {d.max():6.2f}"
    print(line)
    prev = k
sm = a[:, -1].astype(int)
if (a[:, -1] < 4096).all():
    cnt = np.bincount(sm, minlength=148)
    end, done = (a[:, 2] - t0) / 1e3, (a[:, 6] - t0) / 1e3
    for k in sorted(set(cnt[cnt > 0])):
        sel = np.isin(sm, np.where(cnt == k)[0])
        print(f"  CTAs on SMs hosting {k} CTA(s): {sel.sum():4d}  streaming done mean {end[sel].mean():6.2f} max {end[sel].max():6.2f};"
              f"  exit mean {done[sel].mean():6.2f} max {done[sel].max():6.2f}")
