#!/bin/bash
# 2-GPU visit: fused PPO under NCCL + the 2-GPU bench line
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_fused_check.py 2>&1 | grep -v "^\s*$\|W1017\|\*\*\*\*" | tail -15 | tee gpurun_out/r2l_ddp.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err
tail -3 gpurun_out/r2l_bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2l_bench_n2.json').read().strip().splitlines()[-1])
print("N=2 value", d["value"], "us/launch", d["ms_per_step"]*1e3, "e2e", d["e2e"]["value"], d["e2e"]["us_per_step"])
p = d["ppo"]; print("ppo", p.get("value"), p.get("update_s_each"), p.get("allreduce_calls"), p.get("update_impl"), p.get("error"))
print("sac", d["sac"].get("value"), d["sac"].get("error"))
PY
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-configs --sweep 65536 > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2l_bench_n1.json').read().strip().splitlines()[-1])
print("N=1 value", d["value"], "us/launch", d["ms_per_step"]*1e3, "e2e", d["e2e"]["value"], d["e2e"]["us_per_step"], "alt", d["e2e_alternative"]["us_per_step"])
p = d["ppo"]; print("ppo", p.get("value"), p.get("update_s_each"), p.get("update_impl"), p.get("error"))
PY
