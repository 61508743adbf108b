#!/usr/bin/env bash
# tuning sweep of the fused step kernel (cluster size x keep-eps x threads); prints kernel ms per config
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 2 4 8; do for k in 0 1; do for t in 256 512; do
  r=$(DU_FUSED_CLUSTER=$c DU_FUSED_KEEP_EPS=$k DU_FUSED_THREADS=$t python bench.py --steps 30 --warmup 3 --no-cpu --batch-sum 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['roofline']['kernel_ms_back_to_back'], d['roofline']['frac'])" 2>&1 | tail -1)
  echo "cluster=$c keep=$k threads=$t -> $r"
done; done; done
python bench.py --steps 30 --warmup 3 --no-cpu --unfused 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('unfused', d['ms_per_step'], d['roofline'])"
