#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/e2e_breakdown.py 2>&1 | tee gpurun_out/r2r_e2e_breakdown.txt
