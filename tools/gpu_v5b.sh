#!/usr/bin/env bash
# after the list-capacity fix: fused parity suite, step time for the data of ranks 0..7 on one GPU, fp16 trip-threshold A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -m gpu -x -q 2>&1 | tail -4
for r in 0 1 2 3 4 5 6 7; do
  timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu --seed-offset $r 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('seed-offset $r step_us', round(d['ms_per_step']*1e3,2), 'kernel_us', round(d['roofline']['kernel_ms']*1e3,2), d['roofline']['kernel'])"
done | tee gpurun_out/r1_v5_seed_sweep.txt
for mt in 8 6 4; do
  DU_FUSED_PRED_MIN_TRIPS=$mt timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu --dtype fp16 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('fp16 min_trips $mt step_us', round(d['ms_per_step']*1e3,2), 'kernel_us', round(d['roofline']['kernel_ms']*1e3,2), d['roofline']['kernel'], round(d['roofline']['frac'],3))"
done | tee gpurun_out/r1_v5_fp16_trips.txt
