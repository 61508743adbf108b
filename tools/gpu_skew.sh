#!/usr/bin/env bash
mkdir -p gpurun_out
for sk in ${1:-0 6 9 12 15 18}; do
  r=$(DU_FUSED_SKEW_US=$sk DU_FUSED_CLUSTER=2 DU_FUSED_THREADS=512 timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['ms_per_step'])" 2>&1 | tail -1)
  echo "skew=$sk -> $r"
done | tee gpurun_out/skew.txt
DU_FUSED_SKEW_US=12 bash tools/gpu_timeline.sh "2:512"
