#!/bin/bash
mkdir -p gpurun_out
DN_PIPE=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined" 2>&1 | tail -2
for mode in 0 2 3 4 8; do
  if [ $mode = 0 ]; then unset DN_PIPE; else export DN_PIPE=$mode; fi
  timeout 300 python bench.py --steps 300 --warmup 30 --no-cpu --no-vecenv --no-ppo --no-configs --sweep 1048576 4194304 > gpurun_out/bench_p$mode.json 2>gpurun_out/bench_p$mode.err
  tail -1 gpurun_out/bench_p$mode.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_p$mode.json').read().strip().splitlines()[-1])
print("DN_PIPE=$mode", "sweep", [(c["envs"], round(c["us_per_launch"],2), round(c["roofline_frac"],3)) for c in d["sweep"]], "S1", round(d["roofline_hbm_s1"].get("us_per_launch"),2), round(d["roofline_hbm_s1"].get("frac"),3))
PY
done
