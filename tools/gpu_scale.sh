#!/usr/bin/env bash
# N-GPU visit (gpurun --gpus N): the uncertainty-step bench and the ImageNet-128 sampling-loop bench at N ranks.
# usage: tools/gpu_scale.sh N [skip_loop]
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$N" = "1" ]; then
  timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu | tail -1 > gpurun_out/r1_v5_scale_n1.json
  [ -z "$2" ] && timeout 600 python bench.py --gpus 1 --workload imagenet128_adm_loop | tail -1 > gpurun_out/r1_v5_loop_n1.json
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/scale_n$N.err | tail -1 > gpurun_out/r1_v5_scale_n$N.json
  [ -z "$2" ] && timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload imagenet128_adm_loop 2>> gpurun_out/scale_n$N.err | tail -1 > gpurun_out/r1_v5_loop_n$N.json
fi
python - <<PY
import json
for f in ["r1_v5_scale_n$N", "r1_v5_loop_n$N"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read())
        print(f, "n_gpus", d["n_gpus"], d["metric"], round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 2))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/scale_n$N.err 2>/dev/null | cut -c1-300
