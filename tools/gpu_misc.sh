#!/usr/bin/env bash
mkdir -p gpurun_out
DU_FUSED_CLUSTER=8 python bench.py --workload uvit256_latent_b128_m5 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -5 | cut -c1-400
for ch in 1 2 4 8 16; do python bench.py --steps 10 --warmup 3 --no-cpu --e2e-chunks $ch 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunks', $ch, 'e2e', round(d['e2e']['value'],1), 'Mpix/s', round(d['e2e']['ms_per_step'],3), 'ms')"; done
python bench.py --workload sd512_latent_b1_m16 --steps 50 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sd', d['roofline']['kernel_ms'], d['ms_per_step'])"
