#!/usr/bin/env bash
# short GPU visit: parity suite, smoke, both bench arms (no ncu)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_ref.json
