#!/bin/bash
# racecheck / synccheck of the host-call paths (resident server included) and of one fused PPO minibatch
mkdir -p gpurun_out
for tool in racecheck synccheck; do
  echo "== $tool, host-call tests" | tee -a gpurun_out/r2y_sanitizer.txt
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "resident_server_equals and 777 or zero_copy" 2>&1 | tail -6 | tee -a gpurun_out/r2y_sanitizer.txt
done
echo "== synccheck, fused PPO forward" | tee -a gpurun_out/r2y_sanitizer.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_ppo_fused.py -m gpu -q -x -k "forward" 2>&1 | tail -6 | tee -a gpurun_out/r2y_sanitizer.txt
