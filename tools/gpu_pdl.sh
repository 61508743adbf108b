#!/usr/bin/env bash
# fused-step parity suite + A/B of the programmatic-dependent-launch overlap (du_batch_sum -> fused step)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -m gpu -x -q 2>&1 | tail -8
for v in 1 0 1 0; do
  DU_FUSED_PDL=$v timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('PDL=$v step_us', round(d['ms_per_step']*1e3,2), 'kernel_us', round(d['roofline']['kernel_ms']*1e3,2), 'value', round(d['value']), 'e2e', round(d['e2e']['value'],1))"
done | tee gpurun_out/r1_v5_pdl_ab.txt
