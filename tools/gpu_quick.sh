#!/bin/bash
# quick kernel iteration visit: parity tests, short bench, one ncu capture.  Run under gpurun.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 500 --warmup 50 --no-cpu --no-vecenv --no-ppo --sweep 4194304 > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print("4096: us/launch", d["ms_per_step"]*1e3, "graph", d["l2_resident"]["us_per_launch"], "floor", d.get("launch_floor"))
print("4Mi S8:", d["roofline_hbm"]["us_per_launch"], d["roofline_hbm"]["frac"], " S1:", d["roofline_hbm_s1"]["us_per_launch"], d["roofline_hbm_s1"]["frac"])
print("e2e", d["e2e"]["us_per_step"])
PY
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 1 -f -o gpurun_out/prof_n4194304_s8 python tools/profile_step.py 4194304 8 34 > gpurun_out/ncu_q.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 1 -f -o gpurun_out/prof_n4096_s8 python tools/profile_step.py 4096 8 34 >> gpurun_out/ncu_q.log 2>&1
ls gpurun_out | head -30
