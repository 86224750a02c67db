#!/bin/bash
# quick kernel iteration visit: parity tests, short bench.  Run under gpurun.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 1000 --warmup 100 --no-cpu --no-vecenv --no-ppo --sweep 4194304 > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err
tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print("4096 headline us/launch", d["ms_per_step"]*1e3, "value", d["value"], "| resident graph", d["l2_resident"]["us_per_launch"], "| flushed", d["flushed_event_bracket"]["us_per_launch"], "| floor", d.get("launch_floor"))
print("4Mi S8:", d["roofline_hbm"]["us_per_launch"], d["roofline_hbm"]["frac"], " S1:", d["roofline_hbm_s1"]["us_per_launch"], d["roofline_hbm_s1"]["frac"])
print("e2e", d["e2e"]["us_per_step"])
for c in d.get("baseline_configs", []): print(c.get("config"), c.get("envs"), c.get("us_per_launch"), c.get("value"), c.get("error"))
PY
