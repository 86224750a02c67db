#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_fused_check.py > gpurun_out/r2l_ddp_full.txt 2>&1
grep -n "Error\|error\|assert\|Traceback\|File \|ok:" gpurun_out/r2l_ddp_full.txt | head -30
