mkdir -p gpurun_out
timeout 75 python -m pytest tests -m gpu -q -x --timeout 60 > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_final_pytest_gpu.log; tail -4 gpurun_out/r2_final_pytest_gpu.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -2 gpurun_out/r2_final_smoke.log
timeout 60 python bench.py --no-loop --steps 20 --warmup 3 > gpurun_out/r2_final_bench_noloop.json 2> gpurun_out/r2_final_bench.err; tail -c 600 gpurun_out/r2_final_bench_noloop.json; tail -3 gpurun_out/r2_final_bench.err
