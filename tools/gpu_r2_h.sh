#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo_fused.py tests/test_gpu_mlp.py -m gpu -q 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -12 | tee gpurun_out/r2h_fused.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_host_api.py tests/test_c_abi_from_c.py -m gpu -q -x 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -25 | tee gpurun_out/r2h_parity.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16x3 2>&1 | grep -v Warn | grep -A40 "^update" | grep -v "at::\|at_cuda\|Mem" | tee gpurun_out/r2h_prof_x3.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --sweep 65536 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -3 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
for k in ("e2e", "e2e_zero_copy", "e2e_staged", "e2e_vecenv", "e2e_vecenv_config1"):
    print(k, json.dumps(d.get(k))[:400])
p = d["ppo"]; print("ppo", p["value"], p["update_s_each"], p["rollout_s_each"], p.get("update_us_per_minibatch"))
PY
