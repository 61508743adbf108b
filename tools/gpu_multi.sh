#!/usr/bin/env bash
# two-GPU visit: NCCL exchange tests, batch-sharded bench at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu | tail -1 > gpurun_out/scale_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 2> gpurun_out/scale_n2.err | tail -1 > gpurun_out/scale_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 2>> gpurun_out/scale_n2.err | tail -1 > gpurun_out/scale_n2_ref.json
python - <<PY
import json
for f in ["scale_n1","scale_n2","scale_n2_ref"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read())
        print(f, d["n_gpus"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],5), "e2e", round(d["e2e"]["value"],1))
    except Exception as e:
        print(f, "ERR", e); print(open("gpurun_out/scale_n2.err").read()[-1500:])
PY
