python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench
from diffusion_uncertainty_b200 import ops
dev = torch.device("cuda:0")
for wl, Bs in (("imagenet128_adm_b128_m5", (128, 64, 32, 16, 8, 4)), ("imagenet64_adm_b128_m5", (128, 64, 32, 16))):
  for B in Bs:
    sb = bench.StepBench(ops, wl, "fp32", B, dev, 1234)
    sb.warm(3); par = sb.parity()
    s_ms, k_ms = sb.quick(50)
    print(f"{wl} B={B:4d} kernel {sb.kernel} step {s_ms*1e3:7.2f} us kernel {k_ms*1e3:7.2f} us  frac {sb.alg_bytes()/(k_ms*1e-3)/1e9/6547.5:.3f}  mask_agree {par['mask_agreement']:.6f}")
PY
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_bench_configs_gpu.py -m gpu -q --timeout 800 2>&1 | tail -3
