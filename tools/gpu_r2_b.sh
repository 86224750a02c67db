#!/bin/bash
# round-2 visit B: fused PPO update numerics + PPO bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo_fused.py -m gpu -q -x -s 2>&1 | tail -40 > gpurun_out/r2b_fused.txt; cat gpurun_out/r2b_fused.txt
timeout 600 python -m pytest tests/test_gpu_ppo.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-vecenv --no-configs --sweep 65536 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print("ppo:", json.dumps(d.get("ppo"))[:1500])
PY
