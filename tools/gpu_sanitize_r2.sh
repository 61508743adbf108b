#!/usr/bin/env bash
# race evidence: compute-sanitizer racecheck / synccheck / memcheck.
# racecheck tracks shared-memory hazards only; the cluster kernels are run with one CTA per image (DU_FUSED_CLUSTER=1), where the
# same code paths execute with __syncthreads in place of the cluster barriers; cluster > 1 is covered by the stealing stress test.
mkdir -p gpurun_out
T=${TAG:-r2_v6}
SEL='test_fused_matches_oracle or test_fused_predictive_kernel_matches_oracle or test_fused_predictive_kernel_quantile_edges or test_fused_every_sample_count'
echo "== racecheck, non-cluster kernels (ops, widen, rng)" > gpurun_out/${T}_racecheck.txt
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 0 python -m pytest tests/test_ops_gpu.py tests/test_widen_gpu.py -m gpu -q -x --timeout 1400 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|Race reported|Error:|hazard" | tail -15 >> gpurun_out/${T}_racecheck.txt
echo "== racecheck, fused kernels, DU_FUSED_CLUSTER=1" >> gpurun_out/${T}_racecheck.txt
DU_FUSED_CLUSTER=1 DU_FUSED_THREADS=512 timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 0 python -m pytest tests/test_fused_gpu.py -m gpu -q -x --timeout 1400 -k "$SEL" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|Race reported|Error:|hazard" | tail -15 >> gpurun_out/${T}_racecheck.txt
echo "== racecheck, fused kernels, default shapes (clusters included)" >> gpurun_out/${T}_racecheck.txt
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 0 python -m pytest tests/test_fused_gpu.py -m gpu -q -x --timeout 1400 -k "test_fused_matches_oracle" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|Race reported|Error:|hazard|Segmentation|core dumped" | tail -15 >> gpurun_out/${T}_racecheck.txt
echo "rc=$?" >> gpurun_out/${T}_racecheck.txt
cat gpurun_out/${T}_racecheck.txt
echo "== synccheck, fused + ops" > gpurun_out/${T}_synccheck.txt
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 0 python -m pytest tests/test_fused_gpu.py tests/test_ops_gpu.py -m gpu -q -x --timeout 1400 -k "$SEL or test_moments or test_quantile" 2>&1 | grep -E "SYNCCHECK SUMMARY|ERROR SUMMARY|passed|failed|Error:|Barrier error" | tail -15 >> gpurun_out/${T}_synccheck.txt
cat gpurun_out/${T}_synccheck.txt
echo "== memcheck, whole GPU suite" > gpurun_out/${T}_memcheck.txt
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests -m gpu -q --timeout 2300 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|Error:" | tail -12 >> gpurun_out/${T}_memcheck.txt
cat gpurun_out/${T}_memcheck.txt
