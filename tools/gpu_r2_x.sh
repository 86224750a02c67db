#!/bin/bash
# peer-memory gradient all-reduce: the 2-rank check (parity with ncclAllReduce, early stop, timing), then the PPO bench line
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/ddp_fused_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -12 | tee gpurun_out/r2x_check_n$N.txt
[ -n "$SKIP_BENCH" ] || timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --no-cpu --no-vecenv --no-configs --steps 20 --warmup 3 2>gpurun_out/r2x_bench_n$N.err | tail -1 > gpurun_out/r2x_bench_n$N.json
[ -n "$SKIP_BENCH" ] && exit 0
python - <<P
import json
d = json.loads(open("gpurun_out/r2x_bench_n$N.json").read())
print({k: d["ppo"].get(k) for k in ("value", "update_s", "rollout_s", "allreduce_calls", "allreduce_impl")})
P
tail -3 gpurun_out/r2x_bench_n$N.err
