#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or full_size or lockstep" 2>&1 | tail -4
for mode in pipe nopipe; do
  if [ $mode = nopipe ]; then export DN_NO_PIPE=1; else unset DN_NO_PIPE; fi
  timeout 300 python bench.py --steps 300 --warmup 30 --no-cpu --no-vecenv --no-ppo --no-configs --sweep 1048576 4194304 > gpurun_out/bench_$mode.json 2>gpurun_out/bench_$mode.err
  tail -2 gpurun_out/bench_$mode.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$mode.json').read().strip().splitlines()[-1])
print("$mode", "sweep", [(c["envs"], round(c["us_per_launch"],2), round(c["roofline_frac"],3)) for c in d["sweep"]], "S1", d["roofline_hbm_s1"].get("us_per_launch"), d["roofline_hbm_s1"].get("frac"))
PY
done
unset DN_NO_PIPE
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 1 -f -o gpurun_out/prof_pipe_n4194304_s8 python tools/profile_step.py 4194304 8 34 > gpurun_out/ncu_pipe.log 2>&1
