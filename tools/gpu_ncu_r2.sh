#!/usr/bin/env bash
# per-kernel evidence: CUDA-event roofline rows, ONE ncu --set full report over every kernel, launch list of the bench, sanitizer runs
mkdir -p gpurun_out
T=${TAG:-r2_v6}
python tools/kernel_zoo.py > gpurun_out/${T}_kernel_rooflines.jsonl 2> gpurun_out/zoo.err; cat gpurun_out/${T}_kernel_rooflines.jsonl | cut -c1-230
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'moments_kernel|quantile_rows_kernel|rows_kernel|batch_sum_rows_kernel|perturb_randn_kernel|fused_step_kernel|fused_pred_kernel|column_kth_kernel|row_sum_kernel|znorm_partial_kernel' -o gpurun_out/zoo_full -f python tools/kernel_zoo.py --once > gpurun_out/ncu_zoo.log 2>&1
python tools/ncu_kernels_summary.py gpurun_out/zoo_full.ncu-rep gpurun_out/${T}_kernels_ncu_full.csv > gpurun_out/ncu_zoo_summary.log 2>&1; tail -3 gpurun_out/ncu_zoo_summary.log | cut -c1-300
rm -f gpurun_out/zoo_full.ncu-rep
# the fused kernel of the bench alone: details + phase split + DRAM traffic per launch (feeds roofline.traffic)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_pred -s 4 -c 1 -o gpurun_out/fused_full -f python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-loop --no-parity --eager > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/fused_full.ncu-rep ${T}_fused gpurun_out > gpurun_out/ncu_summary.log 2>&1; head -12 gpurun_out/ncu_summary.log
ncu -i gpurun_out/fused_full.ncu-rep --page details --csv > gpurun_out/${T}_fused_ncu_details.csv 2>/dev/null
rm -f gpurun_out/fused_full.ncu-rep
# launch list of the default bench command (kernel share of the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_fused_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --no-loop --eager > gpurun_out/ncu_bench.log 2>&1
grep -c "du::" gpurun_out/${T}_fused_launches.csv
