#!/usr/bin/env bash
# A/B of environment switches on the default bench: tools/gpu_ab.sh "VAR=a VAR=b ..." (each run twice, interleaved)
mkdir -p gpurun_out
for rep in 1 2; do
for kv in $1; do
  env $kv timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$kv step_us', round(d['ms_per_step']*1e3,2), 'kernel_us', round(d['roofline']['kernel_ms']*1e3,2), 'frac', round(d['roofline']['frac'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value'],1))"
done
done | tee gpurun_out/ab_$(date +%H%M%S).txt
