#!/bin/bash
# bench.py under torchrun on N GPUs of one box, launched the way the driver launches it (both arms)
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu \
    > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err
tail -c 600 gpurun_out/r2f_bench_n$N.err
python - <<P
import json
d = json.loads(open("gpurun_out/r2f_bench_n$N.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "reps", "clocks")})
print("e2e", {k: d["e2e"].get(k) for k in ("form", "value", "us_per_step")})
print("ppo", {k: d["ppo"].get(k) for k in ("value", "update_s", "rollout_s", "allreduce_calls")}, "sac", d["sac"].get("value"))
P
