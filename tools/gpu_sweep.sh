#!/usr/bin/env bash
# GPU visit: fused-kernel parity tests, then the configuration sweep with per-CTA timelines (tools/sweep_fused.py)
# usage: gpu_sweep.sh "<-:cluster:threads:tmem[:KNOB=v,...]> ..." [bench]
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py -m gpu -x -q > gpurun_out/pytest_fused.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fused.log
tail -15 gpurun_out/pytest_fused.log
rm -f gpurun_out/sweep_fused.jsonl
timeout 900 python tools/sweep_fused.py --timeline -- "${1:--:-:-:- -:-:-:-:PRED=0 -:1:1024:- -:-:-:-:BAND_SIGMA=4}" > gpurun_out/sweep_fused.log 2>&1
timeout 300 python tools/sweep_fused.py --dtype fp16 -- "-:-:-:- -:-:-:-:PRED=0" >> gpurun_out/sweep_fused.log 2>&1
cat gpurun_out/sweep_fused.log
if [ -n "$2" ]; then
  timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json
fi
