#!/bin/bash
# programmatic dependent launch between the PPO contraction kernels: parity tests, then A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_ppo_fused.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2s_tests.txt
for p in bf16x3 bf16; do
  echo "== PDL on, $p" | tee -a gpurun_out/r2s_ab.txt
  timeout 300 python tools/profile_ppo_fused.py 1048576 32768 $p 2>&1 | grep -E "^update|sum of kernel" | tee -a gpurun_out/r2s_ab.txt
  echo "== PDL off, $p" | tee -a gpurun_out/r2s_ab.txt
  DN_MLP_NO_PDL=1 timeout 300 python tools/profile_ppo_fused.py 1048576 32768 $p 2>&1 | grep -E "^update|sum of kernel" | tee -a gpurun_out/r2s_ab.txt
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-vecenv --no-configs 2>&1 | tail -1 > gpurun_out/r2s_bench.json
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2s_bench.json").read())
print({k: d["ppo"].get(k) for k in ("value", "update_s", "rollout_s", "iter_s")})
P
