#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo_fused.py -m gpu -q 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -12 | tee gpurun_out/r2o_fused.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-vecenv --no-configs --sweep 65536 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
tail -3 gpurun_out/r2o_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
p = d["ppo"]; print("ppo", p.get("value"), p.get("update_s_each"), p.get("rollout_s_each"), p.get("update_us_per_minibatch"), p.get("error"))
PY
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-vecenv --no-configs --sweep 65536 --ppo-precision bf16 > gpurun_out/r2o_bench_bf16.json 2> gpurun_out/r2o_bench_bf16.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2o_bench_bf16.json').read().strip().splitlines()[-1])
p = d["ppo"]; print("ppo bf16", p.get("value"), p.get("update_s_each"), p.get("update_us_per_minibatch"), p.get("error"))
PY
