#!/usr/bin/env bash
# round-2 first GPU visit: whole GPU suite, smoke, the new bench line (no loop), timeline with SM ids
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -25 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-loop > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
DU_FUSED_TIMELINE=gpurun_out/r2_timeline_raw.txt timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-loop --no-parity --eager > /dev/null 2>&1
python - <<'PY'
import numpy as np
a = np.loadtxt("gpurun_out/r2_timeline_raw.txt")
t0 = a[:, 0].min()
sm = a[:, -1].astype(int)
end = (a[:, 2] - t0) / 1e3
cnt = np.bincount(sm, minlength=148)
for k in (1, 2):
    sel = np.isin(sm, np.where(cnt == k)[0])
    if sel.any():
        print(f"CTAs on SMs hosting {k} CTA(s): {sel.sum():4d}  streaming done mean {end[sel].mean():6.2f} min {end[sel].min():6.2f} max {end[sel].max():6.2f} us;  exit mean {((a[sel,6]-t0)/1e3).mean():6.2f} max {((a[sel,6]-t0)/1e3).max():6.2f}")
print("SMs used", (cnt > 0).sum(), "with 2 CTAs", (cnt == 2).sum(), "with 1", (cnt == 1).sum())
PY
python tools/timeline.py gpurun_out/r2_timeline_raw.txt > gpurun_out/r2_timeline_summary.txt; cat gpurun_out/r2_timeline_summary.txt | head -12
