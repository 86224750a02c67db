#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo_fused.py tests/test_gpu_mlp.py -m gpu -q 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -8 | tee gpurun_out/r2i_fused.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "step_host or vec_env" 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -8 | tee gpurun_out/r2i_host.txt
timeout 900 python -m pytest tests/test_host_api.py tests/test_c_abi_from_c.py -m gpu -q -x 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -8
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16x3 2>&1 | grep -v Warn | grep -A40 "^update" | grep -v "at::\|at_cuda\|Mem" | tee gpurun_out/r2i_prof_x3.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16 2>&1 | grep -v Warn | grep -A12 "^update" | grep -v "at::\|at_cuda\|Mem" | tee gpurun_out/r2i_prof_bf16.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --no-ppo --sweep 65536 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
for k in ("e2e", "e2e_alternative", "e2e_staged", "e2e_vecenv", "e2e_vecenv_config1"):
    print(k, json.dumps(d.get(k))[:300])
PY
