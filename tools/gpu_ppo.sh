#!/bin/bash
# PPO learner visit: graph-vs-eager test, profile, PPO bench lines.  Run under gpurun.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ppo.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_ppo.txt; cat gpurun_out/pytest_ppo.txt
python tools/profile_ppo.py > gpurun_out/ppo_profile.txt 2>&1; head -3 gpurun_out/ppo_profile.txt; tail -8 gpurun_out/ppo_profile.txt
python bench.py --steps 200 --warmup 20 --no-cpu --no-vecenv --sweep > gpurun_out/bench_ppo.json 2> gpurun_out/bench_ppo.err; tail -c 1500 gpurun_out/bench_ppo.json; tail -3 gpurun_out/bench_ppo.err
