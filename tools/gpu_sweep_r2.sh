#!/usr/bin/env bash
# sweep of the predictive kernel's launch shape on the ImageNet-128 step (fp32 and fp16): kernel us, frac
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  for dt in fp32 fp16; do
    r=$(env "$@" timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu --no-extras --no-loop --no-parity --dtype $dt 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel'], round(d['roofline']['kernel_ms']*1e3,2), round(d['roofline']['frac'],4), round(d['ms_per_step']*1e3,2))" 2>&1 | tail -1)
    echo "$label $dt -> $r"
  done
}
run "default(2:512)" A=1
run "1:1024" DU_FUSED_CLUSTER=1 DU_FUSED_THREADS=1024
run "1:1024:smem100" DU_FUSED_CLUSTER=1 DU_FUSED_THREADS=1024 DU_FUSED_SMEM_KB=50
run "2:512:nosteal" DU_FUSED_STEAL=0
run "2:512:sigma4" DU_FUSED_BAND_SIGMA=4
run "2:512:sigma8" DU_FUSED_BAND_SIGMA=8
run "2:512:smem48" DU_FUSED_SMEM_KB=48
run "2:512:nohints" DU_L2_HINTS=0
