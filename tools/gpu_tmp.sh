timeout 1500 python -m pytest tests -m gpu -q --timeout 800 2>&1 | tail -8
python tools/kernel_zoo.py 2>&1 | grep -E "znorm|row_sum|quantile|perturb_randn" | cut -c1-175
python tools/bench_graphed.py 2>&1 | tail -4 | tee gpurun_out/r2_graphed_window_step.jsonl
