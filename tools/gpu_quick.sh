#!/usr/bin/env bash
# quick GPU iteration: parity tests of the kernels, the fused-step tuning sweep, one ncu capture
# usage: gpu_quick.sh "<cluster:threads> ..." "<cluster:threads for ncu>" [pytest-targets]
mkdir -p gpurun_out
timeout 900 python -m pytest ${3:-tests/test_fused_gpu.py tests/test_ops_gpu.py} -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -6 gpurun_out/pytest_quick.log
for cfg in ${1:-"1:1024 2:512"}; do
  c=${cfg%%:*}; t=${cfg##*:}
  r=$(DU_FUSED_CLUSTER=$c DU_FUSED_THREADS=$t timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['ms_per_step'])" 2>&1 | tail -1)
  echo "cluster=$c threads=$t -> $r"
done | tee gpurun_out/sweep2.txt
timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu --unfused 2>&1 | tail -1 > gpurun_out/bench_unfused.json
timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu --dtype fp16 2>&1 | tail -1 > gpurun_out/bench_fp16.json
python -c "
import json
for f in ['bench_unfused','bench_fp16']:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms'], d['roofline']['frac'])
"
if [ -n "$2" ]; then bash tools/gpu_ncu.sh "$2"; fi
