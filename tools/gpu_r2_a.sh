#!/bin/bash
# round-2 visit A: tcgen05 GEMM numerics + bench at the driver's flags (checks the small --steps fix)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2a_mlp.txt; cat gpurun_out/r2a_mlp.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-vecenv --no-ppo --no-configs --sweep 4194304 > gpurun_out/r2a_bench20.json 2> gpurun_out/r2a_bench20.err
tail -3 gpurun_out/r2a_bench20.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2a_bench20.json').read().strip().splitlines()[-1])
print("K=20: us/launch", d["ms_per_step"]*1e3, "value", d["value"], "reps", d.get("reps"), "timed_s", d.get("timed_region_s"), "| resident", d["l2_resident"]["us_per_launch"], "| e2e", d["e2e"]["us_per_step"])
PY
