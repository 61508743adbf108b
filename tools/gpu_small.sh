#!/usr/bin/env bash
# small workloads (launch/latency bound): kernel time per cluster size
for w in cifar10_ddpm_b16_m5 sd512_latent_b1_m16 uvit256_latent_b128_m5 imagenet64_adm_b128_m5; do
  for c in "" 1 2 4 8; do
    r=$(DU_FUSED_CLUSTER=$c timeout 120 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['roofline']['kernel_ms']*1e3,2), 'us kernel;', round(d['ms_per_step']*1e3,2), 'us step')" 2>&1 | tail -1)
    echo "$w cluster=${c:-auto} -> $r"
  done
done
