#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_fused_check.py > gpurun_out/r2m_ddp_full.txt 2>&1
grep -n "Error\|assert\|ok:" gpurun_out/r2m_ddp_full.txt | head -10
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -30 | tee gpurun_out/r2m_pytest_gpu.txt
