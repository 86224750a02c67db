#!/bin/bash
# One GPU visit: parity tests, bench, ncu launch list, ncu full captures.  Run under gpurun.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -rA 2>&1 | grep -v "^$" | tail -400 > gpurun_out/pytest_gpu.txt; tail -5 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3 | tee gpurun_out/bench_default_time.txt
python bench.py --steps 1000 --warmup 100 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 64 --warmup 3 --no-cpu --no-vecenv --no-ppo --no-configs --rotating-handles 8 --sweep 4194304 > gpurun_out/ncu_bench.log 2>&1
for cfg in "4194304 8" "4194304 1" "4096 8"; do
  set -- $cfg
  ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 1 -f -o gpurun_out/prof_n$1_s$2 \
      python tools/profile_step.py $1 $2 34 > gpurun_out/ncu_n$1_s$2.log 2>&1
done
ls -la gpurun_out
