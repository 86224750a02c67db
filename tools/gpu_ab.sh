#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -4
for mode in pdl nopdl; do
  if [ $mode = nopdl ]; then export DN_NO_PDL=1; else unset DN_NO_PDL; fi
  timeout 300 python bench.py --steps 1000 --warmup 100 --no-cpu --no-vecenv --no-ppo --no-configs --sweep 65536 4194304 > gpurun_out/bench_$mode.json 2>gpurun_out/bench_$mode.err
  tail -2 gpurun_out/bench_$mode.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$mode.json').read().strip().splitlines()[-1])
print("$mode", "headline us", d["ms_per_step"]*1e3, "| resident graph", d["l2_resident"]["us_per_launch"], "plain", d["l2_resident"]["plain_us_per_launch"], "| flushed", d["flushed_event_bracket"]["us_per_launch"])
print("   sweep", [(c["envs"], round(c["us_per_launch"],2)) for c in d["sweep"]], "S1", d["roofline_hbm_s1"].get("us_per_launch"), "e2e", d["e2e"]["us_per_step"])
PY
done
