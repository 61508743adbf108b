#!/usr/bin/env bash
# One full GPU visit: parity tests, smoke, bench (both arms), other workloads, ncu launch list and one full capture of the fused kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_ref.json
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu --eager > gpurun_out/bench_eager.json 2>> gpurun_out/bench.err
timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu --unfused 2>&1 | tail -1 > gpurun_out/bench_unfused.json
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu --dtype fp16 2>&1 | tail -1 > gpurun_out/bench_fp16.json
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu --batch-sum 0 2>&1 | tail -1 > gpurun_out/bench_nobatchsum.json
for w in cifar10_ddpm_b16_m5 imagenet64_adm_b128_m5 uvit256_latent_b128_m5 sd512_latent_b1_m16; do
  timeout 120 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_$w.json
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(f.split("/")[-1], "ms/step", round(d["ms_per_step"],5), "kernel_ms", r.get("kernel_ms"), "frac", r.get("frac"), "e2e", d.get("e2e",{}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --eager > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_ -s 4 -c 1 -o gpurun_out/fused_full -f python bench.py --steps 3 --warmup 3 --no-cpu --eager > gpurun_out/ncu_full.log 2>&1
# summarise the capture ON THE BOX (the report itself is larger than what gpurun copies back)
python tools/ncu_summary.py gpurun_out/fused_full.ncu-rep ${TAG:-r1_v5_fused} gpurun_out > gpurun_out/ncu_summary.log 2>&1
ncu -i gpurun_out/fused_full.ncu-rep --page details --csv > gpurun_out/${TAG:-r1_v5_fused}_ncu_details.csv 2>/dev/null
rm -f gpurun_out/fused_full.ncu-rep
cat gpurun_out/smoke.log | tail -5
ls -la gpurun_out | head -40
