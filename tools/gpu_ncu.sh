#!/usr/bin/env bash
# ncu full capture of the fused step kernel for the configurations named in $1 (e.g. "1:1024 2:512")
mkdir -p gpurun_out
for cfg in ${1:-"2:512"}; do
  c=${cfg%%:*}; t=${cfg##*:}
  DU_FUSED_CLUSTER=$c DU_FUSED_THREADS=$t timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step_kernel -s 4 -c 1 \
    -o gpurun_out/fused_c${c}_t${t} -f python bench.py --steps 3 --warmup 3 --no-cpu --eager ${2:-} > gpurun_out/ncu_c${c}_t${t}.log 2>&1
  tail -2 gpurun_out/ncu_c${c}_t${t}.log | cut -c1-300
done
ls -la gpurun_out/*.ncu-rep
