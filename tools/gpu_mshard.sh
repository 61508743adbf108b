#!/usr/bin/env bash
# SD-512 latent, M = 16 sharded over N ranks (tools/gpu_mshard.sh N): NCCL test + the M-sharded step bench at 1 and N ranks
N=${1:-2}
mkdir -p gpurun_out
true
timeout 200 python bench.py --workload sd512_latent_b1_m16 --shard-m --steps 200 --warmup 10 | tail -1 > gpurun_out/r1_v5_mshard_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --workload sd512_latent_b1_m16 --shard-m --steps 200 --warmup 10 2> gpurun_out/mshard.err | tail -1 > gpurun_out/r1_v5_mshard_n$N.json
python - <<PY
import json
for n in (1, $N):
    try:
        d = json.loads(open("gpurun_out/r1_v5_mshard_n%d.json" % n).read())
        print("M-sharded n_gpus", d["n_gpus"], "us/step", round(d["us_per_step"], 2), "Mpix/s", round(d["value"], 1), "identical", d["x_prev_identical_on_all_ranks"], "launches", d["gpu_launches"])
    except Exception as e:
        print(n, "ERR", e)
PY
tail -3 gpurun_out/mshard.err | cut -c1-300
