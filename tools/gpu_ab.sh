#!/bin/bash
mkdir -p gpurun_out
for mode in 128 64; do
  if [ $mode = 64 ]; then cp tools/micro/libdronenav64.so drl-dronenavigation_b200/libdronenav.so; fi
  if [ $mode = 64 ]; then timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py -m gpu -x -q 2>&1 | tail -2; fi
  timeout 300 python bench.py --steps 1000 --warmup 100 --no-cpu --no-vecenv --no-ppo --sweep 1048576 4194304 > gpurun_out/bench_b$mode.json 2>gpurun_out/bench_b$mode.err
  tail -1 gpurun_out/bench_b$mode.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_b$mode.json').read().strip().splitlines()[-1])
print("block=$mode headline", round(d["ms_per_step"]*1e3,3), "resident", round(d["l2_resident"]["us_per_launch"],3), "sweep", [(c["envs"], round(c["us_per_launch"],2)) for c in d["sweep"]], "S1", round(d["roofline_hbm_s1"].get("us_per_launch"),2), "e2e", round(d["e2e"]["us_per_step"],2))
print("   configs", [(c["config"], c["envs"], round(c["us_per_launch"],2)) for c in d["baseline_configs"][:5]])
PY
done
