#!/usr/bin/env bash
# One GPU visit: parity tests, bench, tuning sweep, ncu launch list and one full capture of the fused kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_ref.json
for c in 2 4 8; do for k in 0 1; do for t in 256 512; do
  r=$(DU_FUSED_CLUSTER=$c DU_FUSED_KEEP_EPS=$k DU_FUSED_THREADS=$t timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu --batch-sum 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['roofline']['kernel_ms_back_to_back'], d['roofline']['frac'])" 2>&1 | tail -1)
  echo "cluster=$c keep=$k threads=$t -> $r"
done; done; done | tee gpurun_out/sweep.txt
timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu --unfused 2>&1 | tail -1 > gpurun_out/bench_unfused.json
timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu --dtype fp16 2>&1 | tail -1 > gpurun_out/bench_fp16.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_step_kernel -c 2 -o gpurun_out/fused_full -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
