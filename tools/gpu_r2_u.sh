#!/bin/bash
# grouped policy + value launches: parity tests, A/B timing, smoke (incl. the fused PPO half), PPO bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_ppo_fused.py tests/test_gpu_ppo.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2u_tests.txt
DN_MLP_NO_GROUP=1 timeout 600 python -m pytest tests/test_gpu_ppo_fused.py -x -q -m gpu 2>&1 | tail -2 | tee -a gpurun_out/r2u_tests.txt
for p in bf16x3 bf16; do
  echo "== grouped, $p" | tee -a gpurun_out/r2u_ab.txt
  timeout 300 python tools/profile_ppo_fused.py 1048576 32768 $p 2>&1 | grep -E "^update|umma_gemm|head_kernel|reduce_kernel" | cut -c1-150 | tee -a gpurun_out/r2u_ab.txt
  echo "== one launch per net, $p" | tee -a gpurun_out/r2u_ab.txt
  DN_MLP_NO_GROUP=1 timeout 300 python tools/profile_ppo_fused.py 1048576 32768 $p 2>&1 | grep -E "^update" | tee -a gpurun_out/r2u_ab.txt
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2u_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-vecenv --no-configs 2>/dev/null | tail -1 > gpurun_out/r2u_bench.json
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2u_bench.json").read())
print({k: d["ppo"].get(k) for k in ("value", "update_s", "rollout_s")})
P
