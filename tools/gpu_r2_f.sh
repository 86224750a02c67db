#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo_fused.py tests/test_gpu_mlp.py -m gpu -q 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -25 | tee gpurun_out/r2f_fused.txt
timeout 1500 python -m pytest tests/test_ref_golden.py tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_ppo.py -m gpu -q -x 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -25 | tee gpurun_out/r2f_parity.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16x3 2>&1 | grep -v Warn | grep -A14 "^update" | tee gpurun_out/r2f_prof_x3.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16 2>&1 | grep -v Warn | grep -A12 "^update" | tee gpurun_out/r2f_prof_bf16.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-vecenv --no-configs --sweep 65536 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print("ppo:", json.dumps(d.get("ppo"))[:1800])
print("sac:", json.dumps(d.get("sac"))[:600])
PY
