#!/bin/bash
# Final single-GPU visit of the round: the whole GPU suite (with every parity report), smoke, both bench arms, sanitizer.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rP 2>&1 | grep -v "^$" > gpurun_out/r2f_pytest_gpu_full.txt
grep -E "^\[|passed|failed|error|PASSED|^_+ test" gpurun_out/r2f_pytest_gpu_full.txt | cut -c1-600 > gpurun_out/r2f_pytest_gpu.txt
tail -3 gpurun_out/r2f_pytest_gpu_full.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2f_smoke.txt
( time timeout 1500 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err ) 2>&1 | tail -3 | tee gpurun_out/r2f_bench_time.txt
( time timeout 900 python bench.py --impl reference > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err ) 2>&1 | tail -3 | tee -a gpurun_out/r2f_bench_time.txt
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "reps", "gpu_launches", "clocks")})
print("roofline", {k: d["roofline"].get(k) for k in ("achieved", "frac", "traffic")})
print("e2e", {k: d["e2e"].get(k) for k in ("form", "value", "us_per_step")}, d.get("e2e_forms"))
print("ppo", {k: d["ppo"].get(k) for k in ("value", "update_s", "rollout_s")}, "sac", d["sac"].get("value"))
print("vecenv12", d.get("e2e_vecenv_config1"))
for c in d.get("sweep", []): print("sweep", c.get("envs"), c.get("us_per_launch"), c.get("roofline_frac"))
r = json.loads(open("gpurun_out/r2f_bench_reference.json").read().strip().splitlines()[-1])
print("reference", r.get("value"), r.get("cpu_baseline"))
P
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "step_host or ragged or zero_copy" 2>&1 | tail -8 > gpurun_out/r2f_sanitizer.txt
echo "exit $?" >> gpurun_out/r2f_sanitizer.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ppo_fused.py -m gpu -q -x -k "forward" 2>&1 | tail -8 >> gpurun_out/r2f_sanitizer.txt
echo "exit $?" >> gpurun_out/r2f_sanitizer.txt
tail -20 gpurun_out/r2f_sanitizer.txt
