#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/micro/epilogue_probe.py 2>&1 | tail -12 | tee gpurun_out/r2j_probe.txt
