#!/usr/bin/env bash
mkdir -p gpurun_out
for cfg in ${1:-"2:512"}; do
  c=${cfg%%:*}; t=${cfg##*:}
  DU_FUSED_TIMELINE=gpurun_out/timeline_c${c}_t${t}.txt DU_FUSED_CLUSTER=$c DU_FUSED_THREADS=$t timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
  echo "== cluster=$c threads=$t"; python tools/timeline.py gpurun_out/timeline_c${c}_t${t}.txt
done
