#!/usr/bin/env bash
# N-GPU visit (gpurun --gpus N): NCCL exchange tests, the default bench line (strong split + weak sub-record + sampling loop),
# the M-sharded SD step at 1 and N ranks.   usage: tools/gpu_multi_r2.sh N
N=${1:-2}
T=${TAG:-r2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_distributed_gpu.py tests/test_bench_configs_gpu.py -m gpu -q -k "distributed or another_device" 2>&1 | tail -4
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $RUN --master-port 29541 bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/${T}_scale_n$N.err | tail -1 > gpurun_out/${T}_scale_n$N.json
timeout 300 $RUN --master-port 29542 bench.py --impl reference --gpus $N --steps 5 --warmup 3 2>> gpurun_out/${T}_scale_n$N.err | tail -1 > gpurun_out/${T}_scale_n${N}_ref.json
timeout 200 python bench.py --workload sd512_latent_b1_m16 --shard-m --steps 200 --warmup 10 | tail -1 > gpurun_out/${T}_mshard_n1.json
timeout 300 $RUN --master-port 29551 bench.py --gpus $N --workload sd512_latent_b1_m16 --shard-m --steps 200 --warmup 10 2> gpurun_out/${T}_mshard_n$N.err | tail -1 > gpurun_out/${T}_mshard_n$N.json
timeout 300 $RUN --master-port 29552 bench.py --gpus $N --workload sd512_latent_b1_m16 --shard-m --steps 200 --warmup 10 --eager 2>> gpurun_out/${T}_mshard_n$N.err | tail -1 > gpurun_out/${T}_mshard_n${N}_eager.json
python - <<PY
import json
def show(f, keys):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read())
        print(f, {k: (round(d[k], 3) if isinstance(d.get(k), float) else d.get(k)) for k in keys})
        return d
    except Exception as e:
        print(f, "ERR", e)
d = show("${T}_scale_n$N", ["n_gpus", "scaling", "value", "ms_per_step"])
if d:
    print("  e2e", round(d["e2e"]["value"], 1), "other_scaling", d.get("other_scaling"), "roofline frac", round(d["roofline"]["frac"], 3))
    lp = d.get("sampling_loop")
    if lp: print("  loop img/s", round(lp["value"], 2), "fwd ms", round(lp["model_forward_ms"], 1), lp["limiter"])
show("${T}_scale_n${N}_ref", ["value", "ms_per_step"])
for f in ["${T}_mshard_n1", "${T}_mshard_n$N", "${T}_mshard_n${N}_eager"]:
    show(f, ["n_gpus", "us_per_step", "x_prev_identical_on_all_ranks", "gpu_launches", "timed_region"])
PY
tail -3 gpurun_out/${T}_scale_n$N.err | cut -c1-300
