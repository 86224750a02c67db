#!/bin/bash
# resident step server: parity tests, then the e2e forms of the bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "step_host" 2>&1 | tail -15 | tee gpurun_out/r2t_tests.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-ppo --no-configs --no-vecenv 2>gpurun_out/r2t_bench.err | tail -1 > gpurun_out/r2t_bench.json
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2t_bench.json").read())
print(json.dumps(d["e2e_forms"], indent=1))
print({k: d["e2e"].get(k) for k in ("form", "value", "us_per_step", "gpu_launches", "server_steps", "server_residencies")})
print("value", d["value"], "ms_per_step", d["ms_per_step"])
P
tail -5 gpurun_out/r2t_bench.err
